// Flow half of libikflow_b200: the inverse (latent -> joint space) pass of the IKFlow conditional normalising flow.
//
// Replaces  output_rev, _ = self.nn_model(latent, c=conditional, rev=True)  + slice + clamp
//           (jstmn/ikflow @ 2f4636e, ikflow/ikflow_solver.py:98-102; graph built by ikflow/model.py:291-356 out of
//           FrEIA 0.2 GLOWCouplingBlock / PermuteRandom / FixedLinearTransform).
//
// One persistent kernel runs a contiguous range of coupling blocks (normally all of them) without any intermediate
// tensor in HBM:
//   * a ROW GROUP is 64 batch rows; rows are independent, so row groups never synchronise with each other;
//   * a TEAM of NT = hidden/64 CTAs owns a row group; CTA t of the team owns hidden features [64t, 64t+64) of every
//     layer.  The tiny flow state u[64][W], the condition and all coupling arithmetic are replicated in every CTA;
//   * first layer of a subnet (K = split+cond <= 16): fp32 SIMT straight from the replicated state;
//   * hidden x hidden layers: warp-level mma.sync.m16n8k16 bf16 tiles with fp32 accumulation, every fp32 operand
//     split into a bf16 head and tail (head*head + head*tail + tail*head, "bf16x3").  The 64x64 weight tiles are
//     pre-split, pre-swizzled and tile-ordered on the host so that one 16 KB bulk-TMA copy (cp.async.bulk) feeds a
//     pipeline stage; the activation operand is the 64-row x 64-k chunk another CTA of the team produced for the
//     previous layer, exchanged through an L2-resident scratch ring with release/acquire flags (chunk c is usable as
//     soon as CTA c has finished -- no team-wide barrier);
//   * last layer of a subnet (N = 2*split <= 16): every CTA reduces its own 64 features in fp32 straight from the
//     accumulators, the NT partial sums are exchanged and summed in a fixed order by every CTA (bitwise identical
//     replicas), then s = clamp*0.636*atan(.), y = (x - t)*exp(-s), the PermuteRandom gather, and at the very end the
//     FixedLinearTransform inverse, the [:, :ndof] slice and the joint-limit clamp -- all in the same kernel.
// Log-determinants are not computed (the solver discards them, ikflow_solver.py:98).

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "flow_common.cuh"
#include "flow_mma.cuh"
#include "flow_umma.cuh"

namespace ikf {

// ---------------------------------------------------------------------------------------------------------------------
// host side

static inline uint16_t bf16_bits_rn(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40);  // NaN
  const uint32_t lsb = (x >> 16) & 1u;
  x += 0x7fffu + lsb;
  return (uint16_t)(x >> 16);
}
// fp16 (round to nearest even; overflow -> inf) through the host half of cuda_fp16.h
static inline uint16_t fp16_bits_rn(float f) { return __half_as_ushort(__float2half_rn(f)); }
static inline float fp16_bits_to_float(uint16_t b) { return __half2float(__ushort_as_half(b)); }
static inline float bf16_bits_to_float(uint16_t b) {
  uint32_t x = (uint32_t)b << 16;
  float f;
  std::memcpy(&f, &x, 4);
  return f;
}

static bool desc_ok(const IkfFlowDesc* d) {
  if (!d) return false;
  if (d->ndim_tot < 2 || d->ndim_tot > IKF_MAX_WIDTH) return false;
  if (d->dim_cond < 1 || d->dim_cond > 8) return false;
  if (d->nb_nodes < 1 || d->nb_nodes > 256) return false;
  if (d->coeff_fn_config < 1 || d->coeff_fn_config > 4) return false;
  if (d->hidden < 64 || d->hidden % 64 != 0 || d->hidden > 2048) return false;
  if (d->ndof < 1 || d->ndof > d->ndim_tot) return false;
  if (d->precision < IKF_PRECISION_BF16X3 || d->precision > IKF_PRECISION_AUTO) return false;
  const int s1 = d->ndim_tot / 2, s2 = d->ndim_tot - s1;
  if (s2 + d->dim_cond > kPad || 2 * s2 > kPad) return false;
  return true;
}

// fp32 values per subnet in state-dict order: for each Linear weight [out,in] then bias [out]
static size_t subnet_weight_count(const IkfFlowDesc* d, int in_dim, int out_dim) {
  const size_t H = d->hidden;
  size_t n = H * in_dim + H;
  n += (size_t)(d->coeff_fn_config - 1) * (H * H + H);
  n += (size_t)out_dim * H + out_dim;
  return n;
}

}  // namespace ikf

// Default cluster size of the weight multicast: measured on B200 (scripts/time_flow.py, same box) clusters of 2 gain
// 3-6 % wherever the weight stream paces the hidden layers (64- / 128-row groups, exchanged first layer) and cost 1.5 %
// in the just-in-time kernel of the 512-pose headline (its hidden layers are paced by the SIMT generation of the
// operand, and the lock step of the two teams only adds jitter); clusters of 4 are never better than 2.
constexpr int kDefaultCluster = 2;
constexpr int kDefaultClusterJit = 1;
constexpr bool kDefaultKSplit = true;

struct IkfFlow {
  IkfFlowDesc desc;
  int device = 0;
  int num_sms = 0;
  int NT = 0, n_big = 0, slots_max = 0;
  void* blob = nullptr;  // one device allocation holding everything below
  size_t blob_bytes = 0, big_w_bytes = 0;
  ikf::FlowParams base;  // pointers + model constants; per-call fields filled at launch
  size_t smem32 = 0, smem64 = 0, smem128 = 0, smem32j = 0;
  float logdet_m = 0.f;  // FixedLinearTransform.logDetM (forward pass)
  float* m_fwd_dev = nullptr;  // device [kPad][kPad], inside blob
  bool jit = false;  // umma engine, 32-row groups: first layer computed just in time by every CTA (flow_umma.cuh)
  int engine = 0;  // 0 = mma.sync tiles (flow_mma.cuh), 1 = tcgen05 / TMEM (flow_umma.cuh)
  uint32_t epoch = 1;  // doubles as launch id; 0 is "no launch aborted"
  int last_grid = 0, last_cluster = 1;
  unsigned long long* trace = nullptr;
  int trace_layers = 0;
  std::vector<int> host_perm;  // [nb][kPad] perm_inv, then [nb][kPad] perm: composed per launch into FlowParams::phys
  size_t smem_bytes = 0;
  // One handle owns ONE exchange workspace (activation ring, partial sums, flags) and one sequence-number space, and a
  // launch occupies every SM (cooperative), so launches of a handle cannot overlap anyway: `mu` serialises the host
  // side (epoch, parameters), `done` / `last_stream` order a launch after the previous one when it comes from another
  // stream.  Any number of host threads and streams may therefore share a handle.
  std::mutex mu;
  cudaEvent_t done = nullptr;
  cudaStream_t last_stream = nullptr;
  bool has_last = false;
  // developer switches, read ONCE when the handle is created (IKFLOW_B200_RT / _DEBUG)
  int forced_rt = 0;
  int debug = 0;
  bool profiling_launch = false;  // IKFLOW_B200_PROFILING_LAUNCH: see flow_launch_locked
  bool tail_split = true;          // IKFLOW_B200_TAIL_SPLIT=0|1: see flow_launch
  int row_base = 0;                // FlowParams::row_base of the next launch (set under the lock)
  // Mirror of the device status word in mapped host memory, written by the kernel itself when it gives up on a wait:
  // [0] status bits, [1] id of the aborted launch.  Read without any synchronisation by the next call on the handle.
  volatile uint32_t* status_host = nullptr;
  const char* last_kernel = "";
  // the kernels of this handle (engine, operand format): single row group per CTA: [0] 32 rows, [1] 64, [2] 128 (tcgen05
  // only), [3] 32 with the just-in-time first layer (tcgen05 only); what a launch picks (flow_launch_locked): [4] k-split pairs
  // up to 576 rows, [7] / [6] / [5] ping-pong with two 32- / 64- / 128-row groups per CTA beyond; [0]-[3] when those are
  // switched off, refused by the driver or not applicable (mma.sync engine, IKFLOW_B200_RT)
  struct Kernel {
    const void* fn = nullptr;
    int threads = 0;
    size_t smem = 0;
    const char* name = "";
    int max_slots_cs[5] = {0, 0, 0, 0, 0};  // [cs]: team slots that are co-resident when launched in clusters of cs CTAs (cs = 2, 4)
  } kern[8];  // [7]: ping-pong with two 32-row groups; [4]: k-split pairs (64-row groups shared by two CTAs per feature tile), [5] / [6]: ping-pong (two 128- / 64-row groups per CTA); tcgen05 only
  // tcgen05 engine: CTAs per cluster for the weight multicast across teams (1 = off); IKFLOW_B200_CLUSTER overrides
  // fused gather (ikf_flow_set_peers): peer-mapped gathered buffers / flag arrays of every rank of the node
  int n_ranks = 0, rank = 0;
  float* peer_out[ikf::kMaxPeers] = {};
  uint32_t* peer_flag[ikf::kMaxPeers] = {};
  uint32_t peer_seq = 0, peer_count_total = 0;
  int cluster_pref = kDefaultCluster, cluster_pref_jit = kDefaultClusterJit, cluster_pref_pp = 1;  // (IKFLOW_B200_CLUSTER_PP: ping-pong kernels)
  // k-split pairs for batches of up to ks_slots_max x 64 rows (IKFLOW_B200_KSPLIT=0|1 overrides the default)
  bool ksplit = kDefaultKSplit;
  int ks_slots_max = 0;
  int ks_private = 2;  // FlowParams::ks_private for hidden = 1024 (IKFLOW_B200_KS_PRIVATE=0|2|4|8)
  // ping-pong kernel for batches of more than one wave of 128-row groups (IKFLOW_B200_PP=0|1)
  bool pingpong = true;
  bool cluster_ok = true;  // cleared if the driver refuses a cooperative launch with clusters
};

using namespace ikf;

extern "C" {

size_t ikf_flow_weight_count(const IkfFlowDesc* desc) {
  if (!desc_ok(desc)) return 0;
  const int s1 = desc->ndim_tot / 2, s2 = desc->ndim_tot - s1;
  return (size_t)desc->nb_nodes * (subnet_weight_count(desc, s1 + desc->dim_cond, 2 * s2) +
                                   subnet_weight_count(desc, s2 + desc->dim_cond, 2 * s1));
}

void ikf_flow_destroy(IkfFlow* flow) {
  if (!flow) return;
  {
    DeviceGuard guard(flow->device);
    if (flow->done) {
      if (flow->has_last) cudaEventSynchronize(flow->done);  // the workspace must outlive the last launch
      cudaEventDestroy(flow->done);
    }
    if (flow->blob) cudaFree(flow->blob);
    if (flow->status_host) cudaFreeHost((void*)flow->status_host);
  }
  delete flow;
}

int ikf_flow_create(const IkfFlowDesc* desc, const float* weights, size_t n_weights, const int64_t* perm_inv,
                    const float* m_inv, const float* flt_b, const float* joint_lo, const float* joint_hi, int device,
                    IkfFlow** out) {
  if (!out) return fail(IKF_EINVAL, "ikf_flow_create: out is NULL");
  *out = nullptr;
  if (!desc_ok(desc))
    return fail(IKF_EINVAL,
                "ikf_flow_create: unsupported description (need 2<=ndim_tot<=16, dim_cond<=8, coeff_fn_config 1..4, "
                "hidden a multiple of 64 <=2048, split+cond<=16)");
  if (!weights || !perm_inv || !m_inv || !flt_b || !joint_lo || !joint_hi)
    return fail(IKF_EINVAL, "ikf_flow_create: NULL parameter array");
  if (n_weights != ikf_flow_weight_count(desc))
    return fail(IKF_EINVAL, "ikf_flow_create: got %zu weights, the description needs %zu", n_weights,
                ikf_flow_weight_count(desc));
  int ndev = 0;
  IKF_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(IKF_EDEVICE, "ikf_flow_create: no CUDA device %d", device);
  cudaDeviceProp prop;
  IKF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(IKF_EDEVICE, "ikf_flow_create: device %d is sm_%d%d, this library is sm_100a only",
                                    device, prop.major, prop.minor);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(IKF_ECUDA, "ikf_flow_create: cudaSetDevice(%d) failed", device);

  IkfFlow* f = new (std::nothrow) IkfFlow();
  if (!f) return fail(IKF_ENOMEM, "ikf_flow_create: host allocation failed");
  f->desc = *desc;
  f->device = device;
  f->num_sms = prop.multiProcessorCount;
  const int W = desc->ndim_tot, s1 = W / 2, s2 = W - s1, H = desc->hidden;
  const int n_big = desc->coeff_fn_config - 1, nb = desc->nb_nodes, n_sub = 2 * nb;
  // engine: tcgen05/TMEM tiles of 128 features when the hidden size allows it, mma.sync tiles of 64 otherwise
  // (IKFLOW_B200_ENGINE=mma|umma overrides, for A/B comparisons)
  int engine = (H % umma::kFTU == 0 && H <= 1024 && n_big > 0) ? 1 : 0;
  if (const char* env = std::getenv("IKFLOW_B200_ENGINE")) {
    if (std::strcmp(env, "mma") == 0) engine = 0;
    else if (std::strcmp(env, "umma") == 0) {
      if (!engine) {
        delete f;
        return fail(IKF_EINVAL, "ikf_flow_create: IKFLOW_B200_ENGINE=umma needs hidden %% 128 == 0, hidden <= 1024 and coeff_fn_config >= 2");
      }
    } else {
      delete f;
      return fail(IKF_EINVAL, "ikf_flow_create: IKFLOW_B200_ENGINE must be 'mma' or 'umma' (got '%s')", env);
    }
  }
  f->engine = engine;
  if (f->desc.precision == IKF_PRECISION_AUTO) f->desc.precision = engine ? IKF_PRECISION_FP16X3 : IKF_PRECISION_BF16X3;
  const bool f16 = f->desc.precision == IKF_PRECISION_FP16X3;
  if (f16 && !engine) {
    delete f;
    return fail(IKF_EINVAL, "ikf_flow_create: IKF_PRECISION_FP16X3 is implemented by the tcgen05 engine only (hidden %% 128 == 0, hidden <= 1024, coeff_fn_config >= 2)");
  }
  if (const char* env = std::getenv("IKFLOW_B200_RT")) {  // debugging / A-B comparisons
    const int v = std::atoi(env);
    if (v == 32 || v == 64 || (v == 128 && engine)) f->forced_rt = v;
  }
  if (const char* env = std::getenv("IKFLOW_B200_DEBUG")) f->debug = std::atoi(env);
  if (const char* env = std::getenv("IKFLOW_B200_PROFILING_LAUNCH")) f->profiling_launch = std::atoi(env) != 0;
  if (const char* env = std::getenv("IKFLOW_B200_TAIL_SPLIT")) f->tail_split = std::atoi(env) != 0;
  const int FT = engine ? umma::kFTU : kFT;  // hidden features per CTA
  const int NT = H / FT;                     // CTAs per team
  const int KCH = H / kKC;                   // 64-wide k-chunks per hidden layer
  const int ctas_per_sm = engine ? 1 : kCtasPerSm;
  f->NT = NT;
  f->n_big = n_big;
  f->slots_max = std::max(1, ctas_per_sm * f->num_sms / NT);
  if (NT > f->num_sms) {
    delete f;
    return fail(IKF_EDEVICE, "ikf_flow_create: hidden=%d needs %d co-resident CTAs, device has %d SMs", H, NT, f->num_sms);
  }
  // offsets inside the per-(subnet, feature tile) block of small fp32 parameters
  const int o_first_w = 0, o_first_b = kPad * FT, o_big_b = o_first_b + FT, o_last_w = o_big_b + kMaxBig * FT;
  const int o_last_b = o_last_w + kPad * FT, small_block = o_last_b + kPad;

  // ---- host-side repack ----
  const size_t plane = (size_t)FT * kKC;  // bf16 elements of one [FT features][64 k] plane
  const size_t big_elems = (size_t)n_sub * n_big * NT * KCH * 2 * plane;
  const size_t small_floats = (size_t)n_sub * NT * small_block;
  std::vector<uint16_t> big(big_elems);
  std::vector<float> small(small_floats, 0.f);
  const float* wp = weights;
  for (int i = 0; i < nb; ++i) {
    for (int sidx = 0; sidx < 2; ++sidx) {
      const int n = 2 * i + sidx;
      const int in_dim = (sidx == 0 ? s1 : s2) + desc->dim_cond;
      const int out_dim = 2 * (sidx == 0 ? s2 : s1);
      // first Linear [H, in_dim], bias [H]  ->  transposed [k][feature]
      const float* w0 = wp;
      const float* b0 = wp + (size_t)H * in_dim;
      wp = b0 + H;
      for (int f_ = 0; f_ < H; ++f_) {
        float* blk = small.data() + ((size_t)n * NT + f_ / FT) * small_block;
        for (int k = 0; k < in_dim; ++k) blk[o_first_w + k * FT + f_ % FT] = w0[(size_t)f_ * in_dim + k];
        blk[o_first_b + f_ % FT] = b0[f_];
      }
      // hidden Linears [H, H], bias [H]  ->  [feature tile][k-chunk][head|tail] swizzled planes
      for (int l = 0; l < n_big; ++l) {
        const float* w = wp;
        const float* b = wp + (size_t)H * H;
        wp = b + H;
        for (int tt = 0; tt < NT; ++tt) {
          for (int c = 0; c < KCH; ++c) {
            uint16_t* hi = big.data() + ((((size_t)n * n_big + l) * NT + tt) * KCH + c) * 2 * plane;
            uint16_t* lo = hi + plane;
            for (int r = 0; r < FT; ++r) {
              const float* src = w + (size_t)(tt * FT + r) * H + c * kKC;
              for (int e = 0; e < kKC; ++e) {
                const uint32_t off = tile_off_bytes(r, e) / 2;
                if (f16) {  // scaled tails, see umma::kTailScaleF16
                  const uint16_t hb = fp16_bits_rn(src[e]);
                  hi[off] = hb;
                  lo[off] = fp16_bits_rn((src[e] - fp16_bits_to_float(hb)) * umma::kTailScaleF16);
                } else {
                  const uint16_t hb = bf16_bits_rn(src[e]);
                  hi[off] = hb;
                  lo[off] = bf16_bits_rn(src[e] - bf16_bits_to_float(hb));
                }
              }
            }
          }
        }
        for (int f_ = 0; f_ < H; ++f_)
          small[((size_t)n * NT + f_ / FT) * small_block + o_big_b + l * FT + f_ % FT] = b[f_];
      }
      // last Linear [out_dim, H], bias [out_dim]  ->  rows per feature tile (umma: float4 slots XOR (o >> 2))
      const float* wl = wp;
      const float* bl = wp + (size_t)out_dim * H;
      wp = bl + out_dim;
      for (int tt = 0; tt < NT; ++tt) {
        float* blk = small.data() + ((size_t)n * NT + tt) * small_block;
        for (int o = 0; o < out_dim; ++o) {
          for (int e = 0; e < FT; ++e) {
            // umma engine: float4 slot j4 = e / 4 of output row o is stored at slot (j4 & ~7) | ((j4 ^ (j4 >> 3) ^ (o >> 2)) & 7)
            // so that the four k-quarters and the output groups read by one warp fall into different banks
            const int j4 = e >> 2;
            const int pos = engine ? ((((j4 & ~7) | ((j4 ^ (j4 >> 3) ^ (o >> 2)) & 7)) << 2) | (e & 3)) : e;
            blk[o_last_w + o * FT + pos] = wl[(size_t)o * H + tt * FT + e];
          }
          blk[o_last_b + o] = bl[o];
        }
      }
    }
  }

  // umma engine, just-in-time first layer: per (subnet, 64-feature k-chunk) [16 k][64 f] weights + [64] bias
  std::vector<float> first_jit;
  if (engine) {
    first_jit.assign((size_t)n_sub * KCH * umma::kJitChunkFloats, 0.f);
    const float* wq = weights;
    for (int i = 0; i < nb; ++i)
      for (int sidx = 0; sidx < 2; ++sidx) {
        const int n = 2 * i + sidx;
        const int in_dim = (sidx == 0 ? s1 : s2) + desc->dim_cond;
        const int out_dim = 2 * (sidx == 0 ? s2 : s1);
        const float* w0 = wq;
        const float* b0 = wq + (size_t)H * in_dim;
        for (int f_ = 0; f_ < H; ++f_) {
          float* blk = first_jit.data() + ((size_t)n * KCH + f_ / kKC) * umma::kJitChunkFloats;
          for (int k = 0; k < in_dim; ++k) blk[k * kKC + f_ % kKC] = w0[(size_t)f_ * in_dim + k];
          blk[kPad * kKC + f_ % kKC] = b0[f_];
        }
        wq += subnet_weight_count(desc, in_dim, out_dim);
      }
  }

  std::vector<int> perm(2 * nb * kPad, 0);  // [nb][kPad] perm_inv, then [nb][kPad] perm (its inverse; forward pass)
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < W; ++j) {
      const int64_t v = perm_inv[(size_t)i * W + j];
      if (v < 0 || v >= W) {
        delete f;
        return fail(IKF_EINVAL, "ikf_flow_create: perm_inv[%d][%d]=%lld out of range", i, j, (long long)v);
      }
      perm[i * kPad + j] = (int)v;
      perm[(nb + i) * kPad + (int)v] = j;  // perm[perm_inv[j]] = j
    }
  f->host_perm = perm;
  std::vector<float> consts(2 * kPad * kPad + 3 * kPad, 0.f);  // M_inv | b | lo | hi | M
  {
    // FixedLinearTransform forward needs M = M_inv^-1 and logDetM = log|det M|: Gauss-Jordan in double (W <= 16);
    // ikf_flow_set_forward_tables replaces both by the stored parameters
    std::vector<double> a((size_t)W * 2 * W, 0.0);
    for (int i = 0; i < W; ++i) {
      for (int j = 0; j < W; ++j) a[(size_t)i * 2 * W + j] = m_inv[(size_t)i * W + j];
      a[(size_t)i * 2 * W + W + i] = 1.0;
    }
    double logdet_inv = 0.0;
    bool singular = false;
    for (int c = 0; c < W && !singular; ++c) {
      int piv = c;
      for (int r = c + 1; r < W; ++r)
        if (std::fabs(a[(size_t)r * 2 * W + c]) > std::fabs(a[(size_t)piv * 2 * W + c])) piv = r;
      if (a[(size_t)piv * 2 * W + c] == 0.0) { singular = true; break; }
      if (piv != c)
        for (int j = 0; j < 2 * W; ++j) std::swap(a[(size_t)piv * 2 * W + j], a[(size_t)c * 2 * W + j]);
      const double d = a[(size_t)c * 2 * W + c];
      logdet_inv += std::log(std::fabs(d));
      for (int j = 0; j < 2 * W; ++j) a[(size_t)c * 2 * W + j] /= d;
      for (int r = 0; r < W; ++r) {
        if (r == c) continue;
        const double m = a[(size_t)r * 2 * W + c];
        if (m != 0.0)
          for (int j = 0; j < 2 * W; ++j) a[(size_t)r * 2 * W + j] -= m * a[(size_t)c * 2 * W + j];
      }
    }
    if (singular) {
      delete f;
      return fail(IKF_EINVAL, "ikf_flow_create: M_inv is singular");
    }
    for (int i = 0; i < W; ++i)
      for (int j = 0; j < W; ++j) consts[kPad * kPad + 3 * kPad + i * kPad + j] = (float)a[(size_t)i * 2 * W + W + j];
    f->logdet_m = (float)(-logdet_inv);
  }
  for (int i = 0; i < W; ++i)
    for (int j = 0; j < W; ++j) consts[i * kPad + j] = m_inv[(size_t)i * W + j];
  for (int j = 0; j < W; ++j) consts[kPad * kPad + j] = flt_b[j];
  for (int j = 0; j < desc->ndof; ++j) {
    consts[kPad * kPad + kPad + j] = joint_lo[j];
    consts[kPad * kPad + 2 * kPad + j] = joint_hi[j];
  }

  // ---- device blob ----
  auto align_up = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  // upper bound, refined below from the occupancy (tcgen05 engine: the ping-pong kernel runs two exchange slots per team)
  const int slots = engine ? 2 * f->slots_max : f->slots_max;
  const size_t off_big = 0;
  const size_t off_small = align_up(off_big + big_elems * 2);
  const size_t off_jit = align_up(off_small + small_floats * 4);
  const size_t off_perm = align_up(off_jit + first_jit.size() * 4);
  const size_t off_consts = align_up(off_perm + perm.size() * 4);
  const size_t off_act = align_up(off_consts + consts.size() * 4);
  const size_t act_bytes = (size_t)slots * 2 * NT * (engine ? umma::kAStrideU : kAChunkStride);
  const size_t off_partial = align_up(off_act + act_bytes);
  const size_t partial_bytes = (size_t)slots * 2 * NT * (engine ? umma::kRTMaxU * umma::kPartRowBytes : kRTMax * kPad * 4);
  const size_t off_flags = align_up(off_partial + partial_bytes);
  const size_t flag_bytes = ((size_t)slots * 2 * NT * 2 + 2 + 2) * 4;  // + status [2] + fused-gather counter [2]
  f->blob_bytes = align_up(off_flags + flag_bytes);
  f->big_w_bytes = big_elems * 2;
  cudaError_t e = cudaMalloc(&f->blob, f->blob_bytes);
  if (e != cudaSuccess) {
    delete f;
    return fail(IKF_ENOMEM, "ikf_flow_create: cudaMalloc(%zu) failed: %s", f->blob_bytes, cudaGetErrorString(e));
  }
  {
    void* hs = nullptr;
    e = cudaHostAlloc(&hs, 4 * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) {
      std::memset(hs, 0, 4 * sizeof(uint32_t));
      f->status_host = (volatile uint32_t*)hs;
      e = cudaEventCreateWithFlags(&f->done, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
      ikf_flow_destroy(f);
      return fail(IKF_ECUDA, "ikf_flow_create: host status / event setup failed: %s", cudaGetErrorString(e));
    }
  }
  uint8_t* base = (uint8_t*)f->blob;
  e = cudaMemset(base + off_act, 0, f->blob_bytes - off_act);
  if (e == cudaSuccess && big_elems) e = cudaMemcpy(base + off_big, big.data(), big_elems * 2, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(base + off_small, small.data(), small_floats * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !first_jit.empty())
    e = cudaMemcpy(base + off_jit, first_jit.data(), first_jit.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(base + off_perm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(base + off_consts, consts.data(), consts.size() * 4, cudaMemcpyHostToDevice);
  if (engine) {
    f->smem32 = sizeof(umma::Smem<32>) + 1024;
    f->smem64 = sizeof(umma::Smem<64>) + 1024;
    f->smem128 = sizeof(umma::Smem<128>) + 1024;
    f->smem32j = sizeof(umma::Smem<32, true>) + 1024;
    // the just-in-time first layer needs every hidden layer to start at ring stage 0
    f->jit = KCH % umma::Cfg<32, true>::kStages == 0 && f->smem32j <= (size_t)prop.sharedMemPerBlockOptin &&
             s2 + desc->dim_cond <= umma::kJitMaxK;
    if (const char* env = std::getenv("IKFLOW_B200_JIT")) f->jit = f->jit && std::atoi(env) != 0;  // A/B comparisons
  } else {
    f->smem32 = sizeof(FlowSmem<32>) + 1024;
    f->smem64 = sizeof(FlowSmem<64>) + 1024;
  }
  f->smem_bytes = f->smem64;
  if (engine && f16) {
    f->kern[0] = {(const void*)umma::flow_inverse_umma_kernel<32, false, true>, umma::Cfg<32>::kThreads, f->smem32, "ikf::umma::flow_inverse_umma_kernel<32,false,true>"};
    f->kern[1] = {(const void*)umma::flow_inverse_umma_kernel<64, false, true>, umma::Cfg<64>::kThreads, f->smem64, "ikf::umma::flow_inverse_umma_kernel<64,false,true>"};
    f->kern[2] = {(const void*)umma::flow_inverse_umma_kernel<128, false, true>, umma::Cfg<128>::kThreads, f->smem128, "ikf::umma::flow_inverse_umma_kernel<128,false,true>"};
    f->kern[3] = {(const void*)umma::flow_inverse_umma_kernel<32, true, true>, umma::Cfg<32, true>::kThreads, f->smem32j, "ikf::umma::flow_inverse_umma_kernel<32,true,true>"};
    f->kern[4] = {(const void*)umma::flow_inverse_umma_kernel<32, false, true, true>, umma::Cfg<32, false, true>::kThreads, sizeof(umma::Smem<32, false, true>) + 1024, "ikf::umma::flow_inverse_umma_kernel<32,false,true,ksplit>"};
    f->kern[5] = {(const void*)umma::flow_inverse_umma_kernel<128, false, true, false, true>, umma::Cfg<128, false, false, true>::kThreads, sizeof(umma::Smem<128, false, false, true>) + 1024, "ikf::umma::flow_inverse_umma_kernel<128,false,true,pingpong>"};
    f->kern[6] = {(const void*)umma::flow_inverse_umma_kernel<64, false, true, false, true>, umma::Cfg<64, false, false, true>::kThreads, sizeof(umma::Smem<64, false, false, true>) + 1024, "ikf::umma::flow_inverse_umma_kernel<64,false,true,pingpong>"};
    f->kern[7] = {(const void*)umma::flow_inverse_umma_kernel<32, false, true, false, true>, umma::Cfg<32, false, false, true>::kThreads, sizeof(umma::Smem<32, false, false, true>) + 1024, "ikf::umma::flow_inverse_umma_kernel<32,false,true,pingpong>"};
  } else if (engine) {
    f->kern[0] = {(const void*)umma::flow_inverse_umma_kernel<32>, umma::Cfg<32>::kThreads, f->smem32, "ikf::umma::flow_inverse_umma_kernel<32,false,false>"};
    f->kern[1] = {(const void*)umma::flow_inverse_umma_kernel<64>, umma::Cfg<64>::kThreads, f->smem64, "ikf::umma::flow_inverse_umma_kernel<64,false,false>"};
    f->kern[2] = {(const void*)umma::flow_inverse_umma_kernel<128>, umma::Cfg<128>::kThreads, f->smem128, "ikf::umma::flow_inverse_umma_kernel<128,false,false>"};
    f->kern[3] = {(const void*)umma::flow_inverse_umma_kernel<32, true>, umma::Cfg<32, true>::kThreads, f->smem32j, "ikf::umma::flow_inverse_umma_kernel<32,true,false>"};
    f->kern[4] = {(const void*)umma::flow_inverse_umma_kernel<32, false, false, true>, umma::Cfg<32, false, true>::kThreads, sizeof(umma::Smem<32, false, true>) + 1024, "ikf::umma::flow_inverse_umma_kernel<32,false,false,ksplit>"};
    f->kern[5] = {(const void*)umma::flow_inverse_umma_kernel<128, false, false, false, true>, umma::Cfg<128, false, false, true>::kThreads, sizeof(umma::Smem<128, false, false, true>) + 1024, "ikf::umma::flow_inverse_umma_kernel<128,false,false,pingpong>"};
    f->kern[6] = {(const void*)umma::flow_inverse_umma_kernel<64, false, false, false, true>, umma::Cfg<64, false, false, true>::kThreads, sizeof(umma::Smem<64, false, false, true>) + 1024, "ikf::umma::flow_inverse_umma_kernel<64,false,false,pingpong>"};
    f->kern[7] = {(const void*)umma::flow_inverse_umma_kernel<32, false, false, false, true>, umma::Cfg<32, false, false, true>::kThreads, sizeof(umma::Smem<32, false, false, true>) + 1024, "ikf::umma::flow_inverse_umma_kernel<32,false,false,pingpong>"};
  } else {
    f->kern[0] = {(const void*)flow_inverse_kernel<32>, kThreads, f->smem32, "ikf::flow_inverse_kernel<32>"};
    f->kern[1] = {(const void*)flow_inverse_kernel<64>, kThreads, f->smem64, "ikf::flow_inverse_kernel<64>"};
  }
  if (!f->jit) f->kern[3] = IkfFlow::Kernel();
  if (const char* env = std::getenv("IKFLOW_B200_KSPLIT")) f->ksplit = std::atoi(env) != 0;
  if (const char* env = std::getenv("IKFLOW_B200_KS_PRIVATE")) {
    const int v = std::atoi(env);
    if (v >= 0 && v <= 8 && v % 2 == 0) f->ks_private = v;
  }
  if (!f->ksplit || KCH % 4 != 0 || NT > 16) f->kern[4] = IkfFlow::Kernel();  // (flag bits: 2 NT <= 32; whole chunk pairs per half)
  if (const char* env = std::getenv("IKFLOW_B200_PP")) f->pingpong = std::atoi(env) != 0;
  if (!f->pingpong || f->kern[5].smem > (size_t)prop.sharedMemPerBlockOptin) f->kern[5] = IkfFlow::Kernel();
  if (!f->pingpong || f->kern[6].smem > (size_t)prop.sharedMemPerBlockOptin) f->kern[6] = IkfFlow::Kernel();
  if (const char* env = std::getenv("IKFLOW_B200_PP64")) { if (std::atoi(env) == 0) f->kern[6] = IkfFlow::Kernel(); }
  if (!f->pingpong || f->kern[7].smem > (size_t)prop.sharedMemPerBlockOptin) f->kern[7] = IkfFlow::Kernel();
  if (const char* env = std::getenv("IKFLOW_B200_PP32")) { if (std::atoi(env) == 0) f->kern[7] = IkfFlow::Kernel(); }
  {
    // the teams spin on each other's flags, so every CTA of a launch must be resident: size the slot count from
    // what the device really fits
    int occ = ctas_per_sm;
    for (const IkfFlow::Kernel& k : f->kern) {
      if (!k.fn || e != cudaSuccess) continue;
      e = cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem);
      int o = 0;
      if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k.fn, k.threads, k.smem);
      occ = std::min(occ, o);
    }
    if (e == cudaSuccess && occ < 1) {
      ikf_flow_destroy(f);
      return fail(IKF_EDEVICE, "ikf_flow_create: the flow kernel does not fit on an SM of device %d", device);
    }
    f->slots_max = std::max(1, occ * f->num_sms / NT);
  }
  if (const char* env = std::getenv("IKFLOW_B200_CLUSTER")) {
    const int v = std::atoi(env);
    if (v == 1 || v == 2 || v == 4) f->cluster_pref = f->cluster_pref_jit = v;
  }
  if (const char* env = std::getenv("IKFLOW_B200_CLUSTER_PP")) {
    const int v = std::atoi(env);
    if (v == 1 || v == 2 || v == 4) f->cluster_pref_pp = v;
  }
  if (e == cudaSuccess && engine && (std::max(std::max(f->cluster_pref, f->cluster_pref_jit), f->cluster_pref_pp) > 1 || f->kern[4].fn)) {
    // how many clusters of cs CTAs the device holds at once (clusters are placed inside a GPC, so this is not num_sms / cs)
    for (IkfFlow::Kernel& k : f->kern) {
      if (!k.fn) continue;
      for (int cs = 2; cs <= 4; cs *= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * NT);
        cfg.blockDim = dim3(k.threads);
        cfg.dynamicSmemBytes = k.smem;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = cs, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, k.fn, &cfg) != cudaSuccess) {
          cudaGetLastError();
          n_clusters = 0;
        }
        k.max_slots_cs[cs] = (n_clusters * cs / NT) / cs * cs;  // whole groups of cs teams
        if (&k == &f->kern[4] && cs == 2) f->ks_slots_max = n_clusters / NT;  // a k-split team = NT clusters of 2
      }
    }
  }
  if (e != cudaSuccess) {
    ikf_flow_destroy(f);
    return fail(IKF_ECUDA, "ikf_flow_create: device setup failed: %s", cudaGetErrorString(e));
  }

  FlowParams& p = f->base;
  std::memset(&p, 0, sizeof(p));
  p.W = W; p.s1 = s1; p.s2 = s2; p.dim_cond = desc->dim_cond; p.nb_nodes = nb; p.n_big = n_big; p.H = H; p.NT = NT;
  p.ndof = desc->ndof; p.precision = f->desc.precision;
  p.clamp_scale = (float)((double)desc->rnvp_clamp * 0.636);
  p.big_w = (const __nv_bfloat16*)(base + off_big);
  p.small = (const float*)(base + off_small);
  p.first_jit = engine ? (const float*)(base + off_jit) : nullptr;
  p.perm_inv = (const int*)(base + off_perm);
  p.perm_fwd = p.perm_inv + (size_t)nb * kPad;
  p.m_inv = (const float*)(base + off_consts);
  p.m_fwd = p.m_inv + kPad * kPad + 3 * kPad;
  f->m_fwd_dev = (float*)(base + off_consts) + kPad * kPad + 3 * kPad;
  p.flt_b = p.m_inv + kPad * kPad;
  p.lo = p.flt_b + kPad;
  p.hi = p.lo + kPad;
  p.act = base + off_act;
  p.partial = (float*)(base + off_partial);
  p.act_flag = (uint32_t*)(base + off_flags);
  p.part_flag = p.act_flag + (size_t)slots * 2 * NT;
  p.status = p.part_flag + (size_t)slots * 2 * NT;
  p.peer_counter = p.status + 2;
  {
    void* dev_view = nullptr;
    e = cudaHostGetDevicePointer(&dev_view, (void*)f->status_host, 0);
    if (e != cudaSuccess) {
      ikf_flow_destroy(f);
      return fail(IKF_ECUDA, "ikf_flow_create: cudaHostGetDevicePointer failed: %s", cudaGetErrorString(e));
    }
    p.status_host = (uint32_t*)dev_view;
  }
  *out = f;
  return IKF_OK;
}

// fused gather of one launch: offset (floats) of the gathered buffer inside every rank's symmetric allocation, its row
// stride, and the first row of this rank's shard
struct GatherArgs {
  size_t offset;
  int ld, row0;
};

static int flow_launch_locked(IkfFlow* flow, const float* in, int in_ld, const float* cond, int cond_ld, int cond_rows,
                              int cond_cols, float* out, int out_ld, int out_cols, int batch, int block_first, int block_last,
                              int finalize, int clamp, void* stream, const char* name, int forward, float* logdet_out,
                              const GatherArgs* gather);

static int flow_launch(IkfFlow* flow, const float* in, int in_ld, const float* cond, int cond_ld, int cond_rows,
                       int cond_cols, float* out, int out_ld, int out_cols, int batch, int block_first, int block_last,
                       int finalize, int clamp, void* stream, const char* name, int forward = 0, float* logdet_out = nullptr,
                       const GatherArgs* gather = nullptr) {
  if (!flow) return fail(IKF_EINVAL, "%s: flow is NULL", name);
  std::lock_guard<std::mutex> lock(flow->mu);  // see IkfFlow::mu
  // Tail split: the ping-pong kernel with 128-row groups works in rounds of slots_max x 256 rows, and a partly filled last
  // round costs as much as a full one.  A remainder that a smaller kernel finishes faster (up to one wave of 128-row
  // groups) goes into a second launch of its own.  (Not with the fused gather: the ranks must agree on the launches.)
  if (flow->engine && flow->kern[5].fn && flow->tail_split && !flow->forced_rt && gather == nullptr && in && out) {
    const int round_rows = flow->slots_max * 2 * 128;
    const int rem = batch % round_rows;
    if (batch > round_rows && rem > 0 && rem <= flow->slots_max * 128) {
      const int first = batch - rem;
      int rc = flow_launch_locked(flow, in, in_ld, cond, cond_ld, cond_rows, cond_cols, out, out_ld, out_cols, first, block_first, block_last,
                                  finalize, clamp, stream, name, forward, logdet_out, nullptr);
      if (rc != IKF_OK) return rc;
      flow->row_base = first;
      rc = flow_launch_locked(flow, in + (size_t)first * in_ld, in_ld, cond, cond_ld, cond_rows, cond_cols, out + (size_t)first * out_ld, out_ld,
                              out_cols, rem, block_first, block_last, finalize, clamp, stream, name, forward,
                              logdet_out ? logdet_out + first : nullptr, nullptr);
      flow->row_base = 0;
      return rc;
    }
  }
  return flow_launch_locked(flow, in, in_ld, cond, cond_ld, cond_rows, cond_cols, out, out_ld, out_cols, batch, block_first, block_last,
                            finalize, clamp, stream, name, forward, logdet_out, gather);
}

static int flow_launch_locked(IkfFlow* flow, const float* in, int in_ld, const float* cond, int cond_ld, int cond_rows,
                              int cond_cols, float* out, int out_ld, int out_cols, int batch, int block_first, int block_last,
                              int finalize, int clamp, void* stream, const char* name, int forward, float* logdet_out,
                              const GatherArgs* gather) {
  if (batch < 0) return fail(IKF_EINVAL, "%s: negative batch %d", name, batch);
  if (batch == 0) return IKF_OK;
  const IkfFlowDesc& d = flow->desc;
  if (!in || !cond || !out) return fail(IKF_EINVAL, "%s: NULL tensor", name);
  if (in_ld < d.ndim_tot || out_cols < 1 || out_cols > d.ndim_tot || out_ld < out_cols)
    return fail(IKF_EINVAL, "%s: bad leading dimension / column count", name);
  if (cond_rows < 1 || cond_cols < 1 || cond_cols > d.dim_cond || cond_ld < cond_cols)
    return fail(IKF_EINVAL, "%s: bad condition shape (%d rows, %d cols, ld %d; dim_cond %d)", name, cond_rows,
                cond_cols, cond_ld, d.dim_cond);
  if (block_first >= d.nb_nodes || block_last < 0 || block_first < block_last)
    return fail(IKF_EINVAL, "%s: bad block range [%d..%d] for %d blocks", name, block_first, block_last, d.nb_nodes);
  DeviceGuard guard(flow->device);
  if (!guard.ok) return fail(IKF_ECUDA, "%s: cudaSetDevice(%d) failed", name, flow->device);
  // A previous launch that gave up on an inter-CTA wait has written into the mapped status word by now (or will have
  // by the time its stream is synchronised): refuse to build on its garbage.  No synchronisation: a plain host read.
  if (flow->status_host[0] & (IKF_STATUS_SYNC_TIMEOUT | IKF_STATUS_RANGE)) {
    const uint32_t bits = flow->status_host[0], id = flow->status_host[1];
    flow->status_host[0] &= ~(IKF_STATUS_SYNC_TIMEOUT | IKF_STATUS_RANGE);
    if (bits & IKF_STATUS_RANGE)
      return fail(IKF_ESTATUS, "%s: an earlier launch on this handle left the fp16 range (IKF_STATUS_RANGE: a hidden activation "
                  "exceeded 65504 in magnitude or was NaN): its output is invalid.  Create the flow with IKF_PRECISION_BF16X3 "
                  "(IKFLOW_B200_PRECISION=bf16x3), which has the exponent range of fp32", name);
    return fail(IKF_ESTATUS, "%s: an earlier launch on this handle (id %u) timed out waiting for another CTA "
                "(IKF_STATUS_SYNC_TIMEOUT): its output is invalid.  The usual cause is a kernel of another process or "
                "stream holding SMs, so that the cooperative launch could not make progress", name, id);
  }

  FlowParams p = flow->base;
  p.in = in; p.in_ld = in_ld; p.cond = cond; p.cond_ld = cond_ld; p.cond_rows = cond_rows; p.cond_cols = cond_cols;
  p.out = out; p.out_ld = out_ld; p.out_cols = out_cols; p.batch = batch;
  p.block_first = block_first; p.block_last = block_last; p.finalize = finalize; p.clamp_out = clamp;
  p.row_base = flow->row_base;
  p.forward = forward; p.logdet_out = logdet_out; p.logdet_m = flow->logdet_m;
  {
    // PermuteRandom folded into the indexing of the flow state (FlowParams::phys)
    const int nblk = block_first - block_last + 1;
    p.fold = (flow->engine && nblk <= kMaxFold) ? 1 : 0;
    if (p.fold) {
      const int nb = flow->base.nb_nodes;
      uint8_t cur[kPad];
      for (int j = 0; j < kPad; ++j) cur[j] = (uint8_t)j;
      for (int bi = 0; bi < nblk; ++bi) {
        const int blk = forward ? block_last + bi : block_first - bi;
        uint8_t nxt[kPad];
        if (forward) {  // x[:, perm] before the block: the table of block bi includes its own gather
          for (int j = 0; j < kPad; ++j) nxt[j] = j < p.W ? cur[flow->host_perm[(size_t)(nb + blk) * kPad + j]] : (uint8_t)j;
          std::memcpy(cur, nxt, kPad);
          std::memcpy(p.phys[bi], cur, kPad);
        } else {        // x[:, perm_inv] after the block
          std::memcpy(p.phys[bi], cur, kPad);
          for (int j = 0; j < kPad; ++j) nxt[j] = j < p.W ? cur[flow->host_perm[(size_t)blk * kPad + j]] : (uint8_t)j;
          std::memcpy(cur, nxt, kPad);
        }
      }
      std::memcpy(p.phys[nblk], cur, kPad);
    }
  }
  if (forward && !flow->engine) return fail(IKF_EINVAL, "%s: the forward pass is implemented by the tcgen05 engine only (hidden %% 128 == 0, hidden <= 1024, coeff_fn_config >= 2)", name);
  // Row groups of 32 while that still fits in one wave of teams (more CTAs in flight, and a partner CTA on every SM
  // to compute while a team waits on an exchange), 64 beyond.
  int rt = ((batch + 31) / 32 <= flow->slots_max) ? 32 : 64;
  // umma engine: 128-row groups once 64-row groups would need more than one wave (twice the rows per weight byte)
  if (flow->engine && (batch + 63) / 64 > flow->slots_max) rt = 128;
  if (flow->forced_rt) rt = flow->forced_rt;
  p.n_rowgroups = (batch + rt - 1) / rt;
  p.slots = std::min(p.n_rowgroups, flow->slots_max);
  // k-split pairs (Cfg::KS): 64-row groups, two CTAs per feature tile; for every batch that fits one wave of such teams
  const bool ks = flow->engine && flow->kern[4].fn && flow->cluster_ok && !flow->forced_rt && (batch + 63) / 64 <= flow->ks_slots_max;
  if (ks) {
    rt = 64;
    p.n_rowgroups = (batch + 63) / 64;
    p.slots = p.n_rowgroups;
  }
  // ping-pong (Cfg::PP): two 128-row groups per CTA, as soon as the 128-row groups would not fit one wave of teams
  const bool pp = flow->engine && flow->kern[5].fn && !flow->forced_rt && !ks && (batch + 127) / 128 > flow->slots_max;
  if (pp) {
    rt = 128;
    p.n_rowgroups = (batch + 127) / 128;
    p.slots = std::min((p.n_rowgroups + 1) / 2, flow->slots_max);  // teams; each runs the exchange slots 2 s, 2 s + 1
  }
  // ... and two 64-row groups per CTA where the single-group kernel would run one wave of 128-row groups (1152 < B <= 2304)
  const bool pp64 = flow->engine && flow->kern[6].fn && !flow->forced_rt && !ks && !pp && (batch + 63) / 64 > flow->slots_max &&
                    (batch + 127) / 128 <= flow->slots_max;
  if (pp64) {
    rt = 64;
    p.n_rowgroups = (batch + 63) / 64;
    p.slots = std::min((p.n_rowgroups + 1) / 2, flow->slots_max);
  }
  // ... and two 32-row groups per CTA where the single-group kernel would run one wave of 64-row groups (576 < B <= 1152)
  const bool pp32 = flow->engine && flow->kern[7].fn && !flow->forced_rt && !ks && !pp && !pp64 && (batch + 31) / 32 > flow->slots_max &&
                    (batch + 63) / 64 <= flow->slots_max;
  if (pp32) {
    rt = 32;
    p.n_rowgroups = (batch + 31) / 32;
    p.slots = std::min((p.n_rowgroups + 1) / 2, flow->slots_max);
  }
  const IkfFlow::Kernel& k = flow->kern[pp ? 5 : pp64 ? 6 : pp32 ? 7 : ks ? 4 : (rt == 32 && flow->jit) ? 3 : rt == 32 ? 0 : rt == 64 ? 1 : 2];
  // Clusters of cs CTAs = the CTAs with the same feature tile of cs neighbouring teams share every weight chunk by
  // multicast (FlowParams::cluster).  Needs cs teams at least; the slot count becomes a multiple of cs (surplus teams
  // walk empty row groups).
  int cs = 1;
  if (ks) cs = 2;  // (the pair is the cluster; no weight multicast: its CTAs multiply different k-chunks)
  else if (flow->engine && flow->cluster_ok) {
    const bool anypp = pp || pp64 || pp32;
    const int teams_wanted = anypp ? (p.n_rowgroups + 1) / 2 : p.n_rowgroups;  // (ping-pong: two row groups per team)
    for (int c = (&k == &flow->kern[3]) ? flow->cluster_pref_jit : anypp ? flow->cluster_pref_pp : flow->cluster_pref; c > 1; c >>= 1)
      if (teams_wanted >= c && k.max_slots_cs[c] >= c) {
        cs = c;
        break;
      }
    if (cs > 1) p.slots = std::min((teams_wanted + cs - 1) / cs * cs, k.max_slots_cs[cs]);
  }
  p.cluster = cs;
  p.ksplit = ks ? 1 : 0;
  p.ks_private = (ks && flow->base.H / 64 == 16) ? flow->ks_private : 0;  // (hidden = 1024: 12 split + 4 private chunks)
  p.n_peers = 0;
  if (gather) {
    p.n_peers = flow->n_ranks;
    p.peer_rank = flow->rank;
    p.peer_row0 = gather->row0;
    p.peer_ld = gather->ld;
    for (int r = 0; r < flow->n_ranks; ++r) {
      p.peer_out[r] = flow->peer_out[r] + gather->offset;
      p.peer_flag[r] = flow->peer_flag[r];
    }
    p.peer_seq = ++flow->peer_seq;
    flow->peer_count_total += (uint32_t)p.slots * (ks ? 2u : 1u);  // the writer CTAs: t = 0 of every team slot (k-split: both halves)
    p.peer_count_target = flow->peer_count_total;
  }
  p.epoch = flow->epoch;
  p.trace = flow->trace;
  p.trace_layers = flow->trace_layers;
  p.debug = flow->debug;
  // every flag of this launch stays below epoch + 1 + (row groups per slot) * (subnets) * (exchanges per subnet)
  const uint32_t rg_per_slot = (uint32_t)((p.n_rowgroups + p.slots * ((pp || pp64 || pp32) ? 2 : 1) - 1) / (p.slots * ((pp || pp64 || pp32) ? 2 : 1)));
  flow->epoch += rg_per_slot * 2u * (uint32_t)(block_first - block_last + 1) * (uint32_t)(flow->n_big + 1) + 2u;

  const int grid = p.slots * flow->NT * (ks ? 2 : 1);
  flow->last_grid = grid;
  flow->last_cluster = cs;
  void* args[] = {(void*)&p};
  // cooperative launch = the driver guarantees that all CTAs of all teams are co-resident (the teams spin on each
  // other's flags)
  const void* fn = k.fn;
  const int threads = k.threads;
  const size_t smem = k.smem;
  flow->last_kernel = k.name;
  flow->smem_bytes = smem;
  // stream-ordered reuse of the workspace: a launch from another stream waits for the previous launch of this handle
  cudaStream_t st = (cudaStream_t)stream;
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &capturing);
  if (flow->has_last && flow->last_stream != st && capturing == cudaStreamCaptureStatusNone) {
    cudaError_t we = cudaStreamWaitEvent(st, flow->done, 0);
    if (we != cudaSuccess) return fail(IKF_ECUDA, "%s: cudaStreamWaitEvent failed: %s", name, cudaGetErrorString(we));
  }
  cudaError_t e;
  if (cs > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeCooperative;
    attrs[0].val.cooperative = 1;
    attrs[1].id = cudaLaunchAttributeClusterDimension;
    attrs[1].val.clusterDim.x = cs, attrs[1].val.clusterDim.y = 1, attrs[1].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 2;
    if (flow->profiling_launch) {
      // IKFLOW_B200_PROFILING_LAUNCH=1 (ncu only): the cluster launch WITHOUT the cooperative attribute.  ncu cannot launch
      // cooperative grids in clusters; under the profiler kernels run alone, and a grid of at most one CTA per SM is then
      // co-resident anyway.  Never for production: nothing guarantees co-residency here.
      cfg.attrs = attrs + 1;
      cfg.numAttrs = 1;
    }
    e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e != cudaSuccess) {
      // the driver will not place this grid in clusters: remember it and fall back to the plain launch (nothing ran)
      cudaGetLastError();
      flow->cluster_ok = false;
      last_error_ref() = std::string("cluster launch refused (") + cudaGetErrorString(e) + "), clusters disabled for this handle";
      // start over without clusters (no k-split pairs either): the selection above sees cluster_ok == false
      if (ks) flow->kern[4] = IkfFlow::Kernel();
      if (gather) {
        flow->peer_count_total -= (uint32_t)p.slots * (ks ? 2u : 1u);
        --flow->peer_seq;
      }
      flow->epoch = p.epoch;
      return flow_launch_locked(flow, in, in_ld, cond, cond_ld, cond_rows, cond_cols, out, out_ld, out_cols, batch, block_first,
                                block_last, finalize, clamp, stream, name, forward, logdet_out, gather);
    }
  } else {
    e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, st);
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail(IKF_ECUDA, "%s: launch failed: %s", name, cudaGetErrorString(e));
  if (capturing == cudaStreamCaptureStatusNone) {
    e = cudaEventRecord(flow->done, st);
    if (e != cudaSuccess) return fail(IKF_ECUDA, "%s: cudaEventRecord failed: %s", name, cudaGetErrorString(e));
    flow->last_stream = st;
    flow->has_last = true;
  }  // (inside a stream capture the graph owns the ordering: use one capturing stream per handle)
  return IKF_OK;
}

int ikf_flow_inverse(IkfFlow* flow, const float* latent, int latent_ld, const float* cond, int cond_ld, int cond_rows,
                     int cond_cols, float* out, int out_ld, int out_cols, int batch, int clamp, void* stream) {
  if (!flow) return fail(IKF_EINVAL, "ikf_flow_inverse: flow is NULL");
  return flow_launch(flow, latent, latent_ld, cond, cond_ld, cond_rows, cond_cols, out, out_ld, out_cols, batch,
                     flow->desc.nb_nodes - 1, 0, 1, clamp, stream, "ikf_flow_inverse");
}

int ikf_flow_inverse_blocks(IkfFlow* flow, const float* state_in, int in_ld, const float* cond, int cond_ld,
                            int cond_rows, int cond_cols, float* state_out, int out_ld, int batch, int block_first,
                            int block_last, void* stream) {
  if (!flow) return fail(IKF_EINVAL, "ikf_flow_inverse_blocks: flow is NULL");
  return flow_launch(flow, state_in, in_ld, cond, cond_ld, cond_rows, cond_cols, state_out, out_ld,
                     flow->desc.ndim_tot, batch, block_first, block_last, 0, 0, stream, "ikf_flow_inverse_blocks");
}

int ikf_flow_forward(IkfFlow* flow, const float* x, int x_ld, const float* cond, int cond_ld, int cond_rows, int cond_cols,
                     float* z_out, int out_ld, float* logdet_out, int batch, void* stream) {
  if (!flow) return fail(IKF_EINVAL, "ikf_flow_forward: flow is NULL");
  return flow_launch(flow, x, x_ld, cond, cond_ld, cond_rows, cond_cols, z_out, out_ld, flow->desc.ndim_tot, batch,
                     flow->desc.nb_nodes - 1, 0, 0, 0, stream, "ikf_flow_forward", 1, logdet_out);
}

int ikf_flow_set_peers(IkfFlow* flow, int n_ranks, int rank, void* const* gather_bufs, void* const* flag_bufs) {
  if (!flow) return fail(IKF_EINVAL, "ikf_flow_set_peers: flow is NULL");
  std::lock_guard<std::mutex> lock(flow->mu);
  if (n_ranks == 0) {  // switch the fused gather off
    flow->n_ranks = 0;
    return IKF_OK;
  }
  if (!flow->engine) return fail(IKF_EINVAL, "ikf_flow_set_peers: the fused gather is implemented by the tcgen05 engine only");
  if (n_ranks < 1 || n_ranks > kMaxPeers || rank < 0 || rank >= n_ranks || !gather_bufs || !flag_bufs)
    return fail(IKF_EINVAL, "ikf_flow_set_peers: need 1 <= n_ranks <= %d, 0 <= rank < n_ranks and both pointer tables", kMaxPeers);
  for (int r = 0; r < n_ranks; ++r) {
    if (!gather_bufs[r] || !flag_bufs[r]) return fail(IKF_EINVAL, "ikf_flow_set_peers: NULL pointer for rank %d", r);
    flow->peer_out[r] = (float*)gather_bufs[r];
    flow->peer_flag[r] = (uint32_t*)flag_bufs[r];
  }
  flow->n_ranks = n_ranks;
  flow->rank = rank;
  flow->peer_seq = 0;  // the flag arrays start zeroed on every rank (the caller's job, before its barrier)
  return IKF_OK;
}

int ikf_flow_inverse_gather(IkfFlow* flow, const float* latent, int latent_ld, const float* cond, int cond_ld, int cond_rows,
                            int cond_cols, int out_cols, int batch, int clamp, size_t gather_offset, int gather_ld, int row0,
                            void* stream) {
  if (!flow) return fail(IKF_EINVAL, "ikf_flow_inverse_gather: flow is NULL");
  if (flow->n_ranks < 1) return fail(IKF_EINVAL, "ikf_flow_inverse_gather: call ikf_flow_set_peers first");
  if (batch < 1) return fail(IKF_EINVAL, "ikf_flow_inverse_gather: every rank needs at least one row (got %d)", batch);
  if (gather_ld < out_cols || row0 < 0) return fail(IKF_EINVAL, "ikf_flow_inverse_gather: bad gathered layout (ld %d, row0 %d)", gather_ld, row0);
  GatherArgs g{gather_offset, gather_ld, row0};
  float* own = flow->peer_out[flow->rank] + gather_offset + (size_t)row0 * gather_ld;
  return flow_launch(flow, latent, latent_ld, cond, cond_ld, cond_rows, cond_cols, own, gather_ld, out_cols, batch,
                     flow->desc.nb_nodes - 1, 0, 1, clamp, stream, "ikf_flow_inverse_gather", 0, nullptr, &g);
}

int ikf_flow_set_forward_tables(IkfFlow* flow, const float* m, float log_det_m) {
  if (!flow || !m) return fail(IKF_EINVAL, "ikf_flow_set_forward_tables: bad arguments");
  DeviceGuard guard(flow->device);
  if (!guard.ok) return fail(IKF_ECUDA, "ikf_flow_set_forward_tables: cudaSetDevice(%d) failed", flow->device);
  const int W = flow->desc.ndim_tot;
  std::vector<float> padded(kPad * kPad, 0.f);
  for (int i = 0; i < W; ++i)
    for (int j = 0; j < W; ++j) padded[i * kPad + j] = m[(size_t)i * W + j];
  IKF_CUDA(cudaMemcpy(flow->m_fwd_dev, padded.data(), padded.size() * 4, cudaMemcpyHostToDevice));
  flow->logdet_m = log_det_m;
  return IKF_OK;
}

int ikf_flow_status(IkfFlow* flow, void* stream, uint32_t* status_out) {
  if (!flow || !status_out) return fail(IKF_EINVAL, "ikf_flow_status: bad arguments");
  DeviceGuard guard(flow->device);
  if (!guard.ok) return fail(IKF_ECUDA, "ikf_flow_status: cudaSetDevice(%d) failed", flow->device);
  std::lock_guard<std::mutex> lock(flow->mu);
  uint32_t host[2] = {0, 0};
  IKF_CUDA(cudaMemcpyAsync(host, flow->base.status, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  IKF_CUDA(cudaMemsetAsync(flow->base.status, 0, sizeof(uint32_t), (cudaStream_t)stream));
  IKF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  *status_out = host[0] | flow->status_host[0];
  flow->status_host[0] = 0;
  return IKF_OK;
}

int ikf_flow_poll_status(IkfFlow* flow, uint32_t* status_out) {
  if (!flow || !status_out) return fail(IKF_EINVAL, "ikf_flow_poll_status: bad arguments");
  *status_out = flow->status_host[0];  // mapped host memory, written by the kernels: no synchronisation
  return IKF_OK;
}

int ikf_flow_debug_trace(IkfFlow* flow, unsigned long long* dev_stamps, int n_layers) {
  if (!flow || n_layers < 0) return fail(IKF_EINVAL, "ikf_flow_debug_trace: bad arguments");
  flow->trace = n_layers > 0 ? dev_stamps : nullptr;
  flow->trace_layers = n_layers;
  return IKF_OK;
}

int ikf_flow_precision(IkfFlow* flow) { return flow ? flow->desc.precision : IKF_EINVAL; }

const char* ikf_flow_last_kernel(IkfFlow* flow) { return flow ? flow->last_kernel : ""; }

int ikf_flow_last_cluster(IkfFlow* flow) { return flow ? flow->last_cluster : IKF_EINVAL; }

int ikf_flow_info(IkfFlow* flow, size_t* packed_weight_bytes, int* grid_ctas_last, int* smem_bytes) {
  if (!flow) return fail(IKF_EINVAL, "ikf_flow_info: flow is NULL");
  if (packed_weight_bytes) *packed_weight_bytes = flow->big_w_bytes + (size_t)2 * flow->desc.nb_nodes * flow->NT * (flow->engine ? umma::kSmallBytesU : kSmallBytes);
  if (grid_ctas_last) *grid_ctas_last = flow->last_grid;
  if (smem_bytes) *smem_bytes = (int)flow->smem_bytes;
  return IKF_OK;
}

}  // extern "C"
