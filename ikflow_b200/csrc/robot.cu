// Kinematics half of libikflow_b200: forward kinematics, geometric Jacobian, the Levenberg-Marquardt step, pose
// errors and the device-side refine/select loop of IKFlowSolver._generate_exact_ik_solutions.
//
// Replaces (reference = jstmn/ikflow @ 2f4636e, third-party jrl @ 2ba7c39 for the arithmetic):
//   robot.forward_kinematics                      ikflow/ikflow_solver.py:114
//   geodesic_distance_between_quaternions         ikflow/ikflow_solver.py:116
//   robot.inverse_kinematics_step_levenburg_marquardt   ikflow/ikflow_solver.py:205,208
//   robot.clamp_to_joint_limits                   ikflow/ikflow_solver.py:102
//   the LM / select / compact loop                ikflow/ikflow_solver.py:197-233
//
// Execution model: one sample per group of 8 lanes (4 samples per warp).  Lane j of a group owns actuated joint j:
// it builds its local transform A_j = F_j * Motion_j(q_j) (F_j = the fixed URDF origins since the previous actuated
// joint, folded on the host), the chain is closed with a 3-round shuffle scan T_j = A_0 ... A_j instead of a serial
// walk, lane j then holds Jacobian column j, row j of J^T J + lambda I and solves the ndof x ndof system with a
// row-distributed LU with partial pivoting (the algorithm behind torch.linalg.solve).  Everything is fp32.
// None of this is HBM- or tensor-bound (about 5 kFLOP and 84 B per sample-step); it is sized for latency.

#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "common.h"

namespace ikf {

constexpr int kGroup = 8;           // lanes per sample
constexpr int kMaxDof = IKF_MAX_DOF;  // == kGroup
constexpr int kThreads = 128;       // 16 samples per CTA
static_assert(kMaxDof == kGroup, "one lane per actuated joint");

// 3x4 rigid transform [R | p], row-major.
struct Xf {
  float r[9];
  float p[3];
};

struct RobotDev {
  int ndof;
  int kind[kMaxDof];       // 1 revolute, 2 prismatic
  float pre[kMaxDof][12];  // F_j
  float axis[kMaxDof][3];  // unit axis in the joint frame
  float post[12];          // fixed links after the last actuated joint
  float lo[kMaxDof], hi[kMaxDof];
};

}  // namespace ikf

struct IkfRobot {
  ikf::RobotDev dev;
  int device;
  double lo64[ikf::kMaxDof], hi64[ikf::kMaxDof];  // the limits as given (the sampler maps its uniforms in fp64)
};

namespace ikf {

__device__ __forceinline__ Xf xf_identity() {
  Xf t;
#pragma unroll
  for (int i = 0; i < 9; ++i) t.r[i] = (i % 4 == 0) ? 1.f : 0.f;
  t.p[0] = t.p[1] = t.p[2] = 0.f;
  return t;
}

__device__ __forceinline__ Xf xf_load(const float* m) {
  Xf t;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) t.r[3 * i + j] = m[4 * i + j];
    t.p[i] = m[4 * i + 3];
  }
  return t;
}

// a * b
__device__ __forceinline__ Xf xf_mul(const Xf& a, const Xf& b) {
  Xf c;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.r[3 * i + j] = a.r[3 * i] * b.r[j] + a.r[3 * i + 1] * b.r[3 + j] + a.r[3 * i + 2] * b.r[6 + j];
    c.p[i] = a.r[3 * i] * b.p[0] + a.r[3 * i + 1] * b.p[1] + a.r[3 * i + 2] * b.p[2] + a.p[i];
  }
  return c;
}

__device__ __forceinline__ Xf xf_shfl(const Xf& a, int src_lane_in_group) {
  Xf c;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.r[i] = __shfl_sync(0xffffffffu, a.r[i], src_lane_in_group, kGroup);
#pragma unroll
  for (int i = 0; i < 3; ++i) c.p[i] = __shfl_sync(0xffffffffu, a.p[i], src_lane_in_group, kGroup);
  return c;
}

// Motion of joint j by q: Rodrigues rotation about the unit axis, or a translation along it.
__device__ __forceinline__ Xf joint_motion(int kind, const float* ax, float q) {
  Xf t = xf_identity();
  if (kind == 1) {
    float s, c;
    sincosf(q, &s, &c);
    const float v = 1.f - c, kx = ax[0], ky = ax[1], kz = ax[2];
    t.r[0] = kx * kx * v + c;
    t.r[1] = kx * ky * v - kz * s;
    t.r[2] = kx * kz * v + ky * s;
    t.r[3] = ky * kx * v + kz * s;
    t.r[4] = ky * ky * v + c;
    t.r[5] = ky * kz * v - kx * s;
    t.r[6] = kz * kx * v - ky * s;
    t.r[7] = kz * ky * v + kx * s;
    t.r[8] = kz * kz * v + c;
  } else if (kind == 2) {
    t.p[0] = ax[0] * q;
    t.p[1] = ax[1] * q;
    t.p[2] = ax[2] * q;
  }
  return t;
}

// Result of the chain scan for the calling lane.
struct Chain {
  Xf own;  // T_j = A_0 ... A_j  (identity-extended for lanes >= ndof)
  Xf ee;   // T_{ndof-1} * post: the end-effector frame (same in all lanes of the group)
};

__device__ __forceinline__ Chain chain_scan(const RobotDev& rb, int j, float qj) {
  Xf a = xf_identity();
  if (j < rb.ndof) a = xf_mul(xf_load(rb.pre[j]), joint_motion(rb.kind[j], rb.axis[j], qj));
  // inclusive scan over the group: T_j = T_{j-d} * T_j
#pragma unroll
  for (int d = 1; d < kGroup; d <<= 1) {
    Xf left = xf_shfl(a, (j >= d) ? j - d : j);
    if (j >= d) a = xf_mul(left, a);
  }
  Chain c;
  c.own = a;
  c.ee = xf_mul(xf_shfl(a, kGroup - 1), xf_load(rb.post));
  return c;
}

// Rotation matrix -> unit quaternion wxyz, choosing the best conditioned of the four candidates (the pytorch3d-style
// routine jrl uses); ties resolve to the first maximum like torch.argmax.
__device__ __forceinline__ void rot_to_quat(const float* m, float* q) {
  const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7],
              m22 = m[8];
  float qa[4];
  qa[0] = sqrtf(fmaxf(1.f + m00 + m11 + m22, 0.f));
  qa[1] = sqrtf(fmaxf(1.f + m00 - m11 - m22, 0.f));
  qa[2] = sqrtf(fmaxf(1.f - m00 + m11 - m22, 0.f));
  qa[3] = sqrtf(fmaxf(1.f - m00 - m11 + m22, 0.f));
  int best = 0;
  float bv = qa[0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (qa[i] > bv) {
      bv = qa[i];
      best = i;
    }
  float c[4];
  if (best == 0) {
    c[0] = qa[0] * qa[0]; c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01;
  } else if (best == 1) {
    c[0] = m21 - m12; c[1] = qa[1] * qa[1]; c[2] = m10 + m01; c[3] = m02 + m20;
  } else if (best == 2) {
    c[0] = m02 - m20; c[1] = m10 + m01; c[2] = qa[2] * qa[2]; c[3] = m12 + m21;
  } else {
    c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = qa[3] * qa[3];
  }
  const float den = 2.f * fmaxf(bv, 0.1f);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = c[i] / den;
}

__device__ __forceinline__ void ee_pose(const Xf& ee, float* pose) {
  pose[0] = ee.p[0];
  pose[1] = ee.p[1];
  pose[2] = ee.p[2];
  rot_to_quat(ee.r, pose + 3);
}

// IKFlowSolver._calculate_pose_error (ikflow_solver.py:112-117): L2 position error, quaternion geodesic in [0, pi].
__device__ __forceinline__ void pose_error(const float* cur, const float* tgt, float* pos_err, float* rot_err) {
  const float dx = cur[0] - tgt[0], dy = cur[1] - tgt[1], dz = cur[2] - tgt[2];
  *pos_err = sqrtf(dx * dx + dy * dy + dz * dz);
  float dot = tgt[3] * cur[3] + tgt[4] * cur[4] + tgt[5] * cur[5] + tgt[6] * cur[6];
  const float eps = 1e-7f;
  dot = fminf(fmaxf(dot, -1.f + eps), 1.f - eps);
  const float pi = 3.14159265358979323846f, two_pi = 6.28318530717958647692f;
  const float d = 2.f * acosf(dot);
  float r = fmodf(d + pi, two_pi);  // torch.remainder: result takes the sign of the divisor
  if (r < 0.f) r += two_pi;
  *rot_err = fabsf(r - pi);
}

// torch.clamp semantics: NaN stays NaN (fminf / fmaxf alone would turn it into a limit)
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x != x ? x : fminf(fmaxf(x, lo), hi); }

// One LM step for the sample owned by this group.  Returns the new value of joint j (lanes >= ndof return 0).
// e = [rpy(q_t (x) q_cur^-1); p_t - p_cur], J rows 0-2 angular / 3-5 linear, dq = solve(J^T J + lambd I, J^T e).
__device__ __forceinline__ float lm_step_group(const RobotDev& rb, int j, float qj, const float* tgt, float lambd,
                                               int clamp) {
  const int n = rb.ndof;
  const Chain ch = chain_scan(rb, j, qj);
  float cur[7];
  ee_pose(ch.ee, cur);

  // Jacobian column j
  float col[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (j < n) {
    const float* ax = rb.axis[j];
    const float wx = ch.own.r[0] * ax[0] + ch.own.r[1] * ax[1] + ch.own.r[2] * ax[2];
    const float wy = ch.own.r[3] * ax[0] + ch.own.r[4] * ax[1] + ch.own.r[5] * ax[2];
    const float wz = ch.own.r[6] * ax[0] + ch.own.r[7] * ax[1] + ch.own.r[8] * ax[2];
    if (rb.kind[j] == 1) {
      const float dx = cur[0] - ch.own.p[0], dy = cur[1] - ch.own.p[1], dz = cur[2] - ch.own.p[2];
      col[0] = wx; col[1] = wy; col[2] = wz;
      col[3] = wy * dz - wz * dy;
      col[4] = wz * dx - wx * dz;
      col[5] = wx * dy - wy * dx;
    } else {
      col[3] = wx; col[4] = wy; col[5] = wz;
    }
  }

  // error vector
  float e[6];
  {
    // quaternion_inverse(cur) = conj / |q|^2, then Hamilton product target (x) inv
    const float nn = cur[3] * cur[3] + cur[4] * cur[4] + cur[5] * cur[5] + cur[6] * cur[6];
    const float w2 = cur[3] / nn, x2 = -cur[4] / nn, y2 = -cur[5] / nn, z2 = -cur[6] / nn;
    const float w1 = tgt[3], x1 = tgt[4], y1 = tgt[5], z1 = tgt[6];
    const float q0 = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2;
    const float q1 = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
    const float q2 = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
    const float q3 = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;
    e[0] = atan2f(2.f * (q0 * q1 + q2 * q3), 1.f - 2.f * (q1 * q1 + q2 * q2));
    e[1] = asinf(clampf(2.f * (q0 * q2 - q3 * q1), -1.f, 1.f));
    e[2] = atan2f(2.f * (q0 * q3 + q1 * q2), 1.f - 2.f * (q2 * q2 + q3 * q3));
    e[3] = tgt[0] - cur[0];
    e[4] = tgt[1] - cur[1];
    e[5] = tgt[2] - cur[2];
  }

  // row j of [A | b], A = J^T J + lambd I (identity rows for the padding lanes keep the LU well defined)
  float row[kMaxDof + 1];
#pragma unroll
  for (int i = 0; i < kMaxDof; ++i) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += col[k] * __shfl_sync(0xffffffffu, col[k], i, kGroup);
    if (i == j) acc = (j >= n) ? 1.f : acc + lambd;
    row[i] = acc;
  }
  {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += col[k] * e[k];
    row[kMaxDof] = acc;
  }

  // LU with partial pivoting, one row per lane
#pragma unroll
  for (int k = 0; k < kMaxDof - 1; ++k) {
    // pivot search over lanes >= k (first maximum wins, like LAPACK isamax)
    float best = (j >= k) ? fabsf(row[k]) : -1.f;
    int arg = j;
#pragma unroll
    for (int d = kGroup / 2; d >= 1; d >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, d, kGroup);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, d, kGroup);
      if (ob > best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
      }
    }
    // swap rows k <-> arg
    const int src = (j == k) ? arg : ((j == arg) ? k : j);
#pragma unroll
    for (int c = k; c <= kMaxDof; ++c) row[c] = __shfl_sync(0xffffffffu, row[c], src, kGroup);
    // eliminate below the pivot
    const float piv = __shfl_sync(0xffffffffu, row[k], k, kGroup);
    const float f = (j > k) ? row[k] * (1.f / piv) : 0.f;  // LAPACK getf2 scales the column by the reciprocal
#pragma unroll
    for (int c = k + 1; c <= kMaxDof; ++c) {
      const float pr = __shfl_sync(0xffffffffu, row[c], k, kGroup);
      row[c] -= f * pr;
    }
  }
  // back substitution
  float x = 0.f;
#pragma unroll
  for (int k = kMaxDof - 1; k >= 0; --k) {
    const float xk = __shfl_sync(0xffffffffu, row[kMaxDof] / row[k], k, kGroup);
    if (j == k) x = xk;
    if (j < k) row[kMaxDof] -= row[k] * xk;
  }

  float out = qj + x;
  if (j >= n) return 0.f;
  if (clamp) out = clampf(out, rb.lo[j], rb.hi[j]);
  return out;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernels

__global__ void __launch_bounds__(kThreads) fk_kernel(RobotDev rb, const float* __restrict__ q,
                                                      float* __restrict__ poses, int m) {
  const int j = threadIdx.x % kGroup;
  const int s = (blockIdx.x * kThreads + threadIdx.x) / kGroup;
  const bool live = s < m;
  const float qj = (live && j < rb.ndof) ? q[(size_t)s * rb.ndof + j] : 0.f;
  const Chain ch = chain_scan(rb, j, qj);
  float pose[7];
  ee_pose(ch.ee, pose);
  if (live && j < 7) {
    float v = pose[0];
#pragma unroll
    for (int i = 1; i < 7; ++i)
      if (j == i) v = pose[i];
    poses[(size_t)s * 7 + j] = v;
  }
}

__global__ void __launch_bounds__(256) clamp_kernel(RobotDev rb, float* __restrict__ q, int m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * rb.ndof) return;
  const int j = i % rb.ndof;
  q[i] = clampf(q[i], rb.lo[j], rb.hi[j]);
}

__global__ void __launch_bounds__(kThreads) lm_step_kernel(RobotDev rb, const float* __restrict__ poses,
                                                           int pose_rows, const float* q_in, float* q_out, int m,
                                                           float lambd, int clamp) {
  const int j = threadIdx.x % kGroup;
  const int s = (blockIdx.x * kThreads + threadIdx.x) / kGroup;
  const bool live = s < m;
  const float qj = (live && j < rb.ndof) ? q_in[(size_t)s * rb.ndof + j] : 0.f;
  float tgt[7] = {0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f};
  if (live) {
    const float* p = poses + (size_t)(s % pose_rows) * 7;
#pragma unroll
    for (int i = 0; i < 7; ++i) tgt[i] = __ldg(p + i);
  }
  const float out = lm_step_group(rb, j, qj, tgt, lambd, clamp);
  if (live && j < rb.ndof) q_out[(size_t)s * rb.ndof + j] = out;
}

// mode 0: pose errors only; mode 1: + joint-limit flag (evaluation_utils.py:100-112: any(q > hi or q < lo))
__global__ void __launch_bounds__(kThreads) pose_error_kernel(RobotDev rb, const float* __restrict__ q,
                                                              const float* __restrict__ poses, int pose_rows,
                                                              float* __restrict__ pos_err, float* __restrict__ rot_err,
                                                              uint8_t* __restrict__ limits_exceeded, int m) {
  const int j = threadIdx.x % kGroup;
  const int s = (blockIdx.x * kThreads + threadIdx.x) / kGroup;
  const bool live = s < m;
  const float qj = (live && j < rb.ndof) ? q[(size_t)s * rb.ndof + j] : 0.f;
  const Chain ch = chain_scan(rb, j, qj);
  float cur[7], tgt[7] = {0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f};
  ee_pose(ch.ee, cur);
  if (live) {
    const float* p = poses + (size_t)(s % pose_rows) * 7;
#pragma unroll
    for (int i = 0; i < 7; ++i) tgt[i] = __ldg(p + i);
  }
  float pe, re;
  pose_error(cur, tgt, &pe, &re);
  const bool exceeded = live && j < rb.ndof && (qj > rb.hi[j] || qj < rb.lo[j]);
  const unsigned ball = __ballot_sync(0xffffffffu, exceeded);
  if (live && j == 0) {
    pos_err[s] = pe;
    rot_err[s] = re;
    if (limits_exceeded) {
      const int g = (threadIdx.x % 32) / kGroup;
      limits_exceeded[s] = ((ball >> (g * kGroup)) & 0xffu) ? 1 : 0;
    }
  }
}

// The LM / select / compact loop of IKFlowSolver._generate_exact_ik_solutions (ikflow_solver.py:197-233) without host
// round trips.  Poses are independent in the reference loop: at every step all repeats of a still-unsolved pose take
// one LM step; if any repeat then meets both thresholds, the valid repeat with the LARGEST repeat index supplies the
// solution (the Python loop at :217-222 overwrites in increasing row order, rows are repeat-major) and every repeat of
// that pose is dropped (:231-232).  One group owns one pose and walks its repeats in increasing order, so "last valid
// wins" needs no atomics.
__global__ void __launch_bounds__(kThreads) lm_refine_kernel(RobotDev rb, const float* __restrict__ poses,
                                                             float* __restrict__ q, int n, int repeat_count,
                                                             int n_steps, float pos_thr, float rot_thr, float lambd,
                                                             float* __restrict__ final_q,
                                                             uint8_t* __restrict__ final_valid,
                                                             int32_t* __restrict__ n_valid) {
  const int j = threadIdx.x % kGroup;
  const int p = (blockIdx.x * kThreads + threadIdx.x) / kGroup;
  const bool live = p < n;
  float tgt[7] = {0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f};
  if (live) {
#pragma unroll
    for (int i = 0; i < 7; ++i) tgt[i] = __ldg(poses + (size_t)p * 7 + i);
  }
  bool solved = !live;  // dead groups idle through the shuffles
  float sol = 0.f;
  for (int step = 0; step < n_steps; ++step) {
    // whole warp leaves together once its four poses are done
    if (__all_sync(0xffffffffu, solved)) break;
    bool any_valid = false;
    for (int k = 0; k < repeat_count; ++k) {
      const size_t row = (size_t)k * n + (live ? p : 0);
      float qj = (!solved && j < rb.ndof) ? q[row * rb.ndof + j] : 0.f;
      qj = lm_step_group(rb, j, qj, tgt, lambd, 1);
      const Chain ch = chain_scan(rb, j, qj);
      float cur[7], pe, re;
      ee_pose(ch.ee, cur);
      pose_error(cur, tgt, &pe, &re);
      if (!solved) {
        if (j < rb.ndof) q[row * rb.ndof + j] = qj;
        if (pe < pos_thr && re < rot_thr) {
          any_valid = true;
          sol = qj;
        }
      }
    }
    if (any_valid) solved = true;
  }
  const bool ok = live && solved;
  if (live) {
    if (j < rb.ndof) final_q[(size_t)p * rb.ndof + j] = ok ? sol : 0.f;
    if (j == 0) final_valid[p] = ok ? 1 : 0;
  }
  if (n_valid) {
    const unsigned ball = __ballot_sync(0xffffffffu, ok && j == 0);
    if (threadIdx.x % 32 == 0 && ball) atomicAdd(n_valid, __popc(ball));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Target-pose generator: q ~ U(lo + eps, hi - eps) per joint, poses = FK(q), one launch (the step before the hot path:
// jrl Robot.sample_joint_angles_and_poses, called at scripts/benchmark_runtime.py:83-86, scripts/evaluate.py:137-139,
// tests/ikflow_solver_test.py:70-72; the klampt self-collision rejection stays out).
//
// Counter-based RNG, Philox4x32-10 (Salmon et al., SC'11): key = 64-bit seed, counter = (sample index lo, hi, block, 0).
// Sample s, joint j uses word j % 4 of block j / 4; u = (word + 0.5) * 2^-32 in (0, 1), mapped in fp64 and rounded to
// fp32 once -- the arithmetic of the oracle's sampler.  A sample depends only on (seed, s): any partition of the index
// range over launches, streams or GPUs gives the same samples.

struct SampleLimits {
  double lo[kMaxDof], span[kMaxDof];
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

__global__ void __launch_bounds__(kThreads) sample_kernel(RobotDev rb, SampleLimits lim, uint64_t seed, uint64_t first,
                                                          float* __restrict__ q_out, float* __restrict__ poses, int m) {
  const int j = threadIdx.x % kGroup;
  const int s = (blockIdx.x * kThreads + threadIdx.x) / kGroup;
  const bool live = s < m;
  const uint64_t idx = first + (uint64_t)(live ? s : 0);
  uint32_t w[4];
  philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)(j >> 2), 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  const uint32_t word = (j & 3) == 0 ? w[0] : (j & 3) == 1 ? w[1] : (j & 3) == 2 ? w[2] : w[3];
  const double u = ((double)word + 0.5) * (1.0 / 4294967296.0);
  // no FMA contraction: the oracle evaluates lo + u * (hi - lo) with two roundings in fp64
  const float qj = (live && j < rb.ndof) ? (float)__dadd_rn(lim.lo[j], __dmul_rn(u, lim.span[j])) : 0.f;
  if (live && j < rb.ndof) q_out[(size_t)s * rb.ndof + j] = qj;
  const Chain ch = chain_scan(rb, j, qj);
  float pose[7];
  ee_pose(ch.ee, pose);
  if (live && j < 7) {
    float v = pose[0];
#pragma unroll
    for (int i = 1; i < 7; ++i)
      if (j == i) v = pose[i];
    poses[(size_t)s * 7 + j] = v;
  }
}

static inline int grid_for(int samples) { return (samples * kGroup + kThreads - 1) / kThreads; }

static void mat_mul_3x4(const double* a, const double* b, double* c) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) c[4 * i + j] = a[4 * i] * b[j] + a[4 * i + 1] * b[4 + j] + a[4 * i + 2] * b[8 + j];
    c[4 * i + 3] = a[4 * i] * b[3] + a[4 * i + 1] * b[7] + a[4 * i + 2] * b[11] + a[4 * i + 3];
  }
}

}  // namespace ikf

using namespace ikf;

extern "C" {

int ikf_robot_create(int n_links, const int32_t* kind, const double* fixed_T, const double* axis, const double* lo,
                     const double* hi, int device, IkfRobot** out) {
  if (!out) return fail(IKF_EINVAL, "ikf_robot_create: out is NULL");
  *out = nullptr;
  if (n_links < 1 || n_links > IKF_MAX_LINKS || !kind || !fixed_T || !axis || !lo || !hi)
    return fail(IKF_EINVAL, "ikf_robot_create: bad arguments (n_links=%d, max %d)", n_links, IKF_MAX_LINKS);
  int ndev = 0;
  IKF_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(IKF_EDEVICE, "ikf_robot_create: no CUDA device %d", device);
  IkfRobot* r = new (std::nothrow) IkfRobot();
  if (!r) return fail(IKF_ENOMEM, "ikf_robot_create: host allocation failed");
  std::memset(&r->dev, 0, sizeof(r->dev));
  r->device = device;
  const double ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  double acc[12];
  std::memcpy(acc, ident, sizeof(acc));
  int nd = 0;
  for (int l = 0; l < n_links; ++l) {
    double tmp[12];
    mat_mul_3x4(acc, fixed_T + 12 * l, tmp);
    std::memcpy(acc, tmp, sizeof(acc));
    if (kind[l] == 0) continue;
    if (kind[l] != 1 && kind[l] != 2) {
      delete r;
      return fail(IKF_EINVAL, "ikf_robot_create: link %d has unknown kind %d", l, kind[l]);
    }
    if (nd >= kMaxDof) {
      delete r;
      return fail(IKF_EINVAL, "ikf_robot_create: more than %d actuated joints", kMaxDof);
    }
    const double* a = axis + 3 * l;
    const double nrm = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (!(nrm > 0.0)) {
      delete r;
      return fail(IKF_EINVAL, "ikf_robot_create: link %d has a zero axis", l);
    }
    r->dev.kind[nd] = kind[l];
    for (int i = 0; i < 12; ++i) r->dev.pre[nd][i] = (float)acc[i];
    for (int i = 0; i < 3; ++i) r->dev.axis[nd][i] = (float)(a[i] / nrm);
    r->dev.lo[nd] = (float)lo[nd];
    r->dev.hi[nd] = (float)hi[nd];
    r->lo64[nd] = lo[nd];
    r->hi64[nd] = hi[nd];
    std::memcpy(acc, ident, sizeof(acc));
    ++nd;
  }
  if (nd == 0) {
    delete r;
    return fail(IKF_EINVAL, "ikf_robot_create: chain has no actuated joint");
  }
  r->dev.ndof = nd;
  for (int i = 0; i < 12; ++i) r->dev.post[i] = (float)acc[i];
  for (int j = nd; j < kMaxDof; ++j) {
    for (int i = 0; i < 12; ++i) r->dev.pre[j][i] = (float)ident[i];
    r->dev.axis[j][2] = 1.f;
  }
  *out = r;
  return IKF_OK;
}

void ikf_robot_destroy(IkfRobot* robot) { delete robot; }

int ikf_robot_ndof(const IkfRobot* robot) { return robot ? robot->dev.ndof : IKF_EINVAL; }

#define IKF_ROBOT_PROLOGUE(name)                                                \
  if (!robot) return fail(IKF_EINVAL, name ": robot is NULL");                  \
  if (m < 0) return fail(IKF_EINVAL, name ": negative sample count %d", m);     \
  if (m == 0) return IKF_OK;                                                    \
  DeviceGuard guard(robot->device);                                             \
  if (!guard.ok) return fail(IKF_ECUDA, name ": cudaSetDevice(%d) failed", robot->device); \
  cudaStream_t st = (cudaStream_t)stream;

#define IKF_LAUNCH_CHECK(name)                                                                      \
  do {                                                                                              \
    g_launch_count.fetch_add(1, std::memory_order_relaxed);                                         \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess) return fail(IKF_ECUDA, name ": launch failed: %s", cudaGetErrorString(_e)); \
  } while (0)

int ikf_forward_kinematics(IkfRobot* robot, const float* q, float* poses_out, int m, void* stream) {
  IKF_ROBOT_PROLOGUE("ikf_forward_kinematics");
  if (!q || !poses_out) return fail(IKF_EINVAL, "ikf_forward_kinematics: NULL tensor");
  fk_kernel<<<grid_for(m), kThreads, 0, st>>>(robot->dev, q, poses_out, m);
  IKF_LAUNCH_CHECK("ikf_forward_kinematics");
  return IKF_OK;
}

int ikf_sample_joint_angles_and_poses(IkfRobot* robot, uint64_t seed, uint64_t first_index, double joint_limit_eps,
                                      float* q_out, float* poses_out, int m, void* stream) {
  IKF_ROBOT_PROLOGUE("ikf_sample_joint_angles_and_poses");
  if (!q_out || !poses_out) return fail(IKF_EINVAL, "ikf_sample_joint_angles_and_poses: NULL tensor");
  if (!(joint_limit_eps >= 0.0)) return fail(IKF_EINVAL, "ikf_sample_joint_angles_and_poses: negative joint_limit_eps");
  SampleLimits lim;
  for (int j = 0; j < kMaxDof; ++j) {
    const double lo = j < robot->dev.ndof ? robot->lo64[j] + joint_limit_eps : 0.0;
    const double hi = j < robot->dev.ndof ? robot->hi64[j] - joint_limit_eps : 0.0;
    if (hi < lo) return fail(IKF_EINVAL, "ikf_sample_joint_angles_and_poses: joint %d has an empty range", j);
    lim.lo[j] = lo;
    lim.span[j] = hi - lo;
  }
  sample_kernel<<<grid_for(m), kThreads, 0, st>>>(robot->dev, lim, seed, first_index, q_out, poses_out, m);
  IKF_LAUNCH_CHECK("ikf_sample_joint_angles_and_poses");
  return IKF_OK;
}

int ikf_clamp_to_joint_limits(IkfRobot* robot, float* q, int m, void* stream) {
  IKF_ROBOT_PROLOGUE("ikf_clamp_to_joint_limits");
  if (!q) return fail(IKF_EINVAL, "ikf_clamp_to_joint_limits: NULL tensor");
  const int total = m * robot->dev.ndof;
  clamp_kernel<<<(total + 255) / 256, 256, 0, st>>>(robot->dev, q, m);
  IKF_LAUNCH_CHECK("ikf_clamp_to_joint_limits");
  return IKF_OK;
}

int ikf_lm_step(IkfRobot* robot, const float* poses, int pose_rows, const float* q_in, float* q_out, int m,
                float lambd, int clamp, void* stream) {
  IKF_ROBOT_PROLOGUE("ikf_lm_step");
  if (!poses || !q_in || !q_out || pose_rows < 1) return fail(IKF_EINVAL, "ikf_lm_step: bad arguments");
  lm_step_kernel<<<grid_for(m), kThreads, 0, st>>>(robot->dev, poses, pose_rows, q_in, q_out, m, lambd, clamp);
  IKF_LAUNCH_CHECK("ikf_lm_step");
  return IKF_OK;
}

int ikf_pose_error(IkfRobot* robot, const float* q, const float* poses, int pose_rows, float* pos_err, float* rot_err,
                   int m, void* stream) {
  IKF_ROBOT_PROLOGUE("ikf_pose_error");
  if (!poses || !q || !pos_err || !rot_err || pose_rows < 1) return fail(IKF_EINVAL, "ikf_pose_error: bad arguments");
  pose_error_kernel<<<grid_for(m), kThreads, 0, st>>>(robot->dev, q, poses, pose_rows, pos_err, rot_err, nullptr, m);
  IKF_LAUNCH_CHECK("ikf_pose_error");
  return IKF_OK;
}

int ikf_evaluate_solutions(IkfRobot* robot, const float* q, const float* poses, int pose_rows, float* pos_err,
                           float* rot_err, uint8_t* limits_exceeded, int m, void* stream) {
  IKF_ROBOT_PROLOGUE("ikf_evaluate_solutions");
  if (!poses || !q || !pos_err || !rot_err || !limits_exceeded || pose_rows < 1)
    return fail(IKF_EINVAL, "ikf_evaluate_solutions: bad arguments");
  pose_error_kernel<<<grid_for(m), kThreads, 0, st>>>(robot->dev, q, poses, pose_rows, pos_err, rot_err,
                                                      limits_exceeded, m);
  IKF_LAUNCH_CHECK("ikf_evaluate_solutions");
  return IKF_OK;
}

int ikf_lm_refine(IkfRobot* robot, const float* poses, float* q_seeds, int n, int repeat_count, int n_steps,
                  float pos_thr, float rot_thr, float lambd, float* final_q, uint8_t* final_valid,
                  int32_t* n_valid_dev, void* stream) {
  const int m = n;
  IKF_ROBOT_PROLOGUE("ikf_lm_refine");
  if (!poses || !q_seeds || !final_q || !final_valid || repeat_count < 1 || n_steps < 0)
    return fail(IKF_EINVAL, "ikf_lm_refine: bad arguments");
  if (n_valid_dev) IKF_CUDA(cudaMemsetAsync(n_valid_dev, 0, sizeof(int32_t), st));
  lm_refine_kernel<<<grid_for(n), kThreads, 0, st>>>(robot->dev, poses, q_seeds, n, repeat_count, n_steps, pos_thr,
                                                     rot_thr, lambd, final_q, final_valid, n_valid_dev);
  IKF_LAUNCH_CHECK("ikf_lm_refine");
  return IKF_OK;
}

}  // extern "C"
