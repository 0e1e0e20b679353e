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

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "common.h"

namespace ikf {

constexpr int kFT = 64;                // hidden features per CTA
constexpr int kKC = 64;                // k elements per pipeline stage
constexpr int kRTMax = 64;             // largest row group
constexpr int kWTileBytes = kFT * kKC * 2;      // one bf16 plane of a weight chunk: 8 KB
constexpr int kWChunkBytes = 2 * kWTileBytes;   // head + tail: 16 KB
constexpr int kAChunkStride = 2 * kRTMax * kKC * 2;  // bytes reserved per activation chunk in the scratch ring: 16 KB
constexpr int kComputeWarps = 4;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kLoaderWarp = kComputeWarps;
constexpr int kStorerWarp = kComputeWarps + 1;
constexpr int kThreads = (kComputeWarps + 2) * 32;
constexpr int kCtasPerSm = 2;  // two teams share every SM: one computes while the other waits on an exchange
constexpr int kPad = 16;       // padded width of state / small-layer dimensions
constexpr int kMaxBig = 3;     // hidden x hidden layers per subnet (coeff_fn_config - 1)
// per (subnet, feature tile) block of small fp32 parameters, one bulk copy:
//   first_wT [16 k][64 f] | first_b [64] | big_b [kMaxBig][64] | last_w [16 o][64 f] | last_b [16]
constexpr int kSmallFirstW = 0;
constexpr int kSmallFirstB = kSmallFirstW + kPad * kFT;
constexpr int kSmallBigB = kSmallFirstB + kFT;
constexpr int kSmallLastW = kSmallBigB + kMaxBig * kFT;
constexpr int kSmallLastB = kSmallLastW + kPad * kFT;
constexpr int kSmallFloats = kSmallLastB + kPad;  // 2320
constexpr int kSmallBytes = kSmallFloats * 4;     // 9280, multiple of 16
static_assert(kSmallBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

constexpr float kLeakySlope = 0.01f;  // nn.LeakyReLU() default, ikflow/model.py:74-83

struct FlowParams {
  int W, s1, s2, dim_cond, nb_nodes, n_big, H, NT, ndof, precision;
  float clamp_scale;  // rnvp_clamp * 0.636, rounded to fp32 the way torch rounds the Python scalar
  const __nv_bfloat16* big_w;  // [subnet][n_big][NT t][NT c][head|tail][64 f][64 k] swizzled
  const float* small;          // [subnet][NT t][kSmallFloats]
  const int* perm_inv;         // [nb_nodes][kPad]
  const float* m_inv;          // [kPad][kPad]  out_j = sum_i (u_i - b_i) m_inv[i][j]
  const float* flt_b;          // [kPad]
  const float* lo;             // [kPad] joint limits
  const float* hi;
  uint8_t* act;         // [slot][2][NT c][kAChunkStride]: head [RT][64 k] then tail, swizzled
  float* partial;       // [slot][2][NT t][kRTMax r][16 o]
  uint32_t* act_flag;   // [slot][2][NT]
  uint32_t* part_flag;  // [slot][2][NT]
  uint32_t* status;     // [0] status bits, [1] id of the launch that aborted
  uint32_t epoch;       // sequence numbers of this launch start at epoch + 1
  unsigned long long* trace;  // debug: [cta < NT][layer][16] globaltimer stamps of team 0 (NULL = off)
  int trace_layers;
  const float* in;
  const float* cond;
  float* out;
  int in_ld, cond_ld, cond_rows, cond_cols, out_ld, out_cols;
  int batch, block_first, block_last, finalize, clamp_out, n_rowgroups, slots;
};

// ---------------------------------------------------------------------------------------------------------------------
// PTX helpers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (try_wait may suspend the thread for a hardware-defined time when the phase is still open).
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a kernel bug must surface as an error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory"); }
// compute threads + storer warp: "the outgoing chunk is staged"
__device__ __forceinline__ void bar_staged_arrive() {
  asm volatile("bar.arrive 2, %0;" ::"n"(kComputeThreads + 32) : "memory");
}
__device__ __forceinline__ void bar_staged_sync() {
  asm volatile("bar.sync 2, %0;" ::"n"(kComputeThreads + 32) : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Wait until *flag has reached `expected` (wrap-safe).  Gives up (and makes every later wait of this launch give up)
// after about a second: the results are then garbage and IKF_STATUS_SYNC_TIMEOUT is reported, but the GPU is not hung.
// The caller issues the acquire fence.
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t expected, uint32_t* status, uint32_t launch_id) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (true) {
    if ((int32_t)(ld_relaxed(flag) - expected) >= 0) return;
    ++spins;
    if (spins == 64) t0 = clock64();
    if (spins > 64) {
      __nanosleep(20);
      if ((spins & 255u) == 0) {
        if (ld_relaxed(status + 1) == launch_id) return;
        if (clock64() - t0 > 2500000000LL) {
          atomicOr(status, IKF_STATUS_SYNC_TIMEOUT);
          atomicExch(status + 1, launch_id);
          return;
        }
      }
    }
  }
}

__device__ __forceinline__ void trace_ev(const FlowParams& p, int layer, int ev) {
  if (p.trace != nullptr && blockIdx.x < p.NT && layer < p.trace_layers) {
    unsigned long long tns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
    p.trace[((size_t)blockIdx.x * p.trace_layers + layer) * 16 + ev] = tns;
  }
}

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : v * kLeakySlope; }

// byte offset of element (row, k) inside a [rows][64] bf16 tile: 128-byte rows, 16-byte chunks XOR-swizzled by
// row % 8 (conflict-free ldmatrix; also the canonical K-major SWIZZLE_128B operand layout of the tensor cores)
__device__ __host__ __forceinline__ uint32_t tile_off_bytes(int row, int k) {
  return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

template <int RT>
struct FlowCfg {
  static constexpr int kMI = RT / 32;                 // m16 tiles per warp (warp tile = 16*kMI rows x 32 features)
  static constexpr int kATileBytes = RT * kKC * 2;    // one bf16 plane of an activation chunk
  static constexpr int kAChunkBytes = 2 * kATileBytes;
  static constexpr int kStageBytes = kAChunkBytes + kWChunkBytes;
  static constexpr int kStages = RT == 64 ? 2 : 3;
};

template <int RT>
struct __align__(1024) FlowSmem {
  using C = FlowCfg<RT>;
  uint8_t ring[C::kStages][C::kStageBytes];  // [activation head|tail][weight head|tail]
  uint8_t staging[C::kAChunkBytes];          // outgoing activation chunk; reused as the last layer's warp partials
  float small[2][kSmallFloats];
  float u[RT][kPad];   // flow state
  float cnd[RT][8];    // condition
  float a[RT][kPad];   // output of the last layer of the current subnet
  uint64_t full[C::kStages], empty[C::kStages];
  uint64_t small_full[2], small_empty[2];
  uint64_t staging_free;
};

// ---------------------------------------------------------------------------------------------------------------------
// the kernel

template <int RT>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) flow_inverse_kernel(const FlowParams p) {
  using C = FlowCfg<RT>;
  constexpr int MI = C::kMI;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  FlowSmem<RT>& sm =
      *reinterpret_cast<FlowSmem<RT>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int NT = p.NT;
  const int slot = blockIdx.x / NT;
  const int t = blockIdx.x % NT;
  const uint32_t launch_id = p.epoch;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], kComputeWarps);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.small_full[b], 1);
      mbar_init(&sm.small_empty[b], kComputeWarps);
    }
    mbar_init(&sm.staging_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  uint8_t* act_slot = p.act + (size_t)slot * 2 * NT * kAChunkStride;
  float* part_slot = p.partial + (size_t)slot * 2 * NT * kRTMax * kPad;
  uint32_t* aflag = p.act_flag + (size_t)slot * 2 * NT;
  uint32_t* pflag = p.part_flag + (size_t)slot * 2 * NT;

  const int n_blocks = p.block_first - p.block_last + 1;
  const int steps_per_rg = 2 * n_blocks;  // subnets per row group
  const int my_rgs = (p.n_rowgroups - slot + p.slots - 1) / p.slots;
  const int total_steps = my_rgs * steps_per_rg;

  // The three roles walk the same schedule: step g = (row group, block, subnet); within a step the layers in order.
  // Activation exchange number x uses scratch buffer x % 2; its flags carry epoch + 1 + (writes so far to that buffer).

  if (warp == kLoaderWarp) {
    // ===== loader: bulk-TMA producer for the small-parameter blocks and the weight/activation ring =====
    // The whole warp polls (lane c watches the flag of chunk c, relaxed loads); lane 0 issues the copies.  Weight
    // chunks are issued as soon as their stage is free, the activation chunk of a stage follows when its producer has
    // published it -- in the fixed order c = t, t+1, ... so that the fp32 accumulation order (and therefore the
    // result) never depends on timing.
    uint32_t ring_pos = 0;
    uint32_t act_w[2] = {0, 0};  // writes so far into each activation scratch buffer
    uint32_t xchg = 0;           // activation exchanges so far
    auto prefetch_small = [&](int g) {
      if (g >= total_steps) return;
      const int b = g & 1;
      if (g >= 2) mbar_wait(&sm.small_empty[b], ((g >> 1) - 1) & 1);
      if (lane == 0) {
        const int in_rg = g % steps_per_rg;
        const int blk = p.block_first - in_rg / 2;
        const int n = 2 * blk + (in_rg & 1);
        mbar_arrive_expect_tx(&sm.small_full[b], kSmallBytes);
        bulk_g2s(sm.small[b], p.small + ((size_t)n * NT + t) * kSmallFloats, kSmallBytes, &sm.small_full[b]);
      }
      __syncwarp();
    };
    prefetch_small(0);
    for (int g = 0; g < total_steps; ++g) {
      const int in_rg = g % steps_per_rg;
      const int blk = p.block_first - in_rg / 2;
      const int n = 2 * blk + (in_rg & 1);
      if (p.n_big == 0) prefetch_small(g + 1);
      for (int l = 0; l < p.n_big; ++l) {
        // input of hidden layer l = exchange number xchg (written by the previous layer of every team member)
        const int buf = xchg & 1;
        const uint32_t expected = p.epoch + 1 + act_w[buf];
        const uint8_t* wbase =
            reinterpret_cast<const uint8_t*>(p.big_w) + (((size_t)n * p.n_big + l) * NT + t) * NT * kWChunkBytes;
        const uint8_t* abase = act_slot + (size_t)buf * NT * kAChunkStride;
        const uint32_t* my_flag = aflag + buf * NT + (lane < NT ? lane : 0);
        {
          // The model (203 MB for Panda) does not stay in L2 between calls, so every weight byte comes from HBM once
          // per launch; pull the NEXT hidden layer's slice of this CTA into L2 now, one whole layer ahead of its use,
          // so that the ring refills at L2 latency instead of DRAM latency.
          int n2 = n, l2 = l + 1;
          if (l2 == p.n_big) {
            l2 = 0;
            n2 = -1;
            if (g + 1 < total_steps) {
              const int in_rg2 = (g + 1) % steps_per_rg;
              n2 = 2 * (p.block_first - in_rg2 / 2) + (in_rg2 & 1);
            }
          }
          if (n2 >= 0) {
            const uint8_t* wnext =
                reinterpret_cast<const uint8_t*>(p.big_w) + (((size_t)n2 * p.n_big + l2) * NT + t) * NT * kWChunkBytes;
            for (int k = lane; k < NT; k += 32) bulk_prefetch_l2(wnext + (size_t)((t + k) % NT) * kWChunkBytes, kWChunkBytes);
          }
        }
        uint32_t ready = 0;  // bit c: chunk c has been published (warp-uniform)
        int issued_w = 0, issued_a = 0;
        bool gave_up = false;
        uint32_t spins = 0;
        long long t0 = 0;
        while (issued_a < NT) {
          // 1) which stages are free for the next weight chunks?  lane k looks at chunk issued_w + k
          int n_w = 0;
          {
            bool free_ = false;
            if (lane < kStages && issued_w + lane < NT) {
              const uint32_t pos = ring_pos + issued_w + lane;
              const uint32_t use = pos / kStages;
              free_ = use == 0 || mbar_test_wait(&sm.empty[pos % kStages], (use - 1) & 1);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, free_);
            n_w = __ffs(~m) - 1;  // consecutive free stages starting at chunk issued_w
          }
          // 2) poll the flags that are still outstanding
          if (!gave_up && ready != 0xffffffffu) {
            bool ok = false;
            if (lane < NT && !((ready >> lane) & 1u)) ok = (int32_t)(ld_relaxed(my_flag) - expected) >= 0;
            ready |= __ballot_sync(0xffffffffu, ok);
            if (__popc(ready) == NT) ready = 0xffffffffu;
          }
          // 3) activation chunks, in order, for stages whose weight copy is (being) issued
          int n_go = 0;
          while (n_go < kStages && issued_a + n_go < issued_w + n_w && ((ready >> ((t + issued_a + n_go) % NT)) & 1u))
            ++n_go;
          if (n_w > 0 || n_go > 0) {
            // One bulk copy keeps its issuing thread busy for ~0.4 us whatever its size, but copies issued by
            // different lanes run concurrently (measured: scripts/ubench/ingest2.cu) -- so every pending copy gets
            // its own lane and all of them go out in one instruction: lanes 0.. the weight chunks, lanes 8.. the
            // activation chunks.
            const bool do_w = lane < n_w;
            const bool do_a = lane >= 8 && lane < 8 + n_go;
            void* dst = nullptr;
            const void* src = nullptr;
            uint32_t bytes = 0;
            uint64_t* bar = nullptr;
            if (do_w) {
              const int iw = issued_w + lane;
              const int st = (ring_pos + iw) % kStages;
              mbar_arrive_expect_tx(&sm.full[st], C::kStageBytes);
              dst = sm.ring[st] + C::kAChunkBytes;
              src = wbase + (size_t)((t + iw) % NT) * kWChunkBytes;
              bytes = kWChunkBytes;
              bar = &sm.full[st];
            }
            if (do_a) {
              const int ia = issued_a + (lane - 8);
              const int st = (ring_pos + ia) % kStages;
              dst = sm.ring[st];
              src = abase + (size_t)((t + ia) % NT) * kAChunkStride;
              bytes = C::kAChunkBytes;
              bar = &sm.full[st];
            }
            __syncwarp();
            // The activation chunks were written by bulk stores (async proxy) that completed before their flag was
            // released and are read here by bulk copies from L2 (no L1 in the path): a proxy fence orders the copies
            // after the flag reads.
            if (n_go > 0) fence_proxy_async();
            if (do_w || do_a) bulk_g2s(dst, src, bytes, bar);
            __syncwarp();
            if (lane == 0) {
              if (issued_w == 0 && n_w > 0) trace_ev(p, g * 4 + l, 0);
              if (issued_a == 0 && n_go > 0) trace_ev(p, g * 4 + l, 1);
              if (issued_a + n_go == NT) trace_ev(p, g * 4 + l, 2);
            }
            const bool first = issued_a == 0 && n_go > 0;
            issued_w += n_w;
            issued_a += n_go;
            spins = 0;
            if (l == 0 && first) prefetch_small(g + 1);
          } else {
            // nothing to do yet: back off a little; give up after about a second (see wait_flag)
            ++spins;
            if (spins == 64) t0 = clock64();
            if (spins > 64) {
              __nanosleep(20);
              if ((spins & 255u) == 0 && !gave_up) {
                int bail = 0;
                if (lane == 0) {
                  if (ld_relaxed(p.status + 1) == launch_id) bail = 1;
                  else if (clock64() - t0 > 2500000000LL) {
                    atomicOr(p.status, IKF_STATUS_SYNC_TIMEOUT);
                    atomicExch(p.status + 1, launch_id);
                    bail = 1;
                  }
                }
                if (__shfl_sync(0xffffffffu, bail, 0)) {
                  gave_up = true;
                  ready = 0xffffffffu;
                }
              }
            }
          }
        }
        ring_pos += NT;
        ++act_w[buf];
        ++xchg;
      }
    }
  } else if (warp == kStorerWarp) {
    // ===== storer: publishes this CTA's activation chunk to the team =====
    uint32_t act_w[2] = {0, 0};
    uint32_t xchg = 0;
    for (int g = 0; g < total_steps; ++g) {
      for (int l = 0; l < p.n_big; ++l) {
        const int buf = xchg & 1;
        bar_staged_sync();  // compute warps have written + proxy-fenced the staging buffer
        if (lane == 0) {
          uint8_t* dst = act_slot + ((size_t)buf * NT + t) * kAChunkStride;
          trace_ev(p, g * 4 + l, 3);
          bulk_s2g(dst, sm.staging, C::kAChunkBytes);
          bulk_commit();
          bulk_wait_all();
          trace_ev(p, g * 4 + l, 4);
          st_release(aflag + buf * NT + t, p.epoch + 1 + act_w[buf]);
          trace_ev(p, g * 4 + l, 5);
          mbar_arrive(&sm.staging_free);
        }
        __syncwarp();
        ++act_w[buf];
        ++xchg;
      }
    }
  } else {
    // ===== compute warps: 2 x 2 over the RT x 64 tile, warp tile (RT/2) rows x 32 features =====
    const int warp_m = warp >> 1;
    const int warp_n = warp & 1;
    const int gq = lane >> 2, tq = lane & 3;
    const int row_base = warp_m * (16 * MI);  // + mi * 16 + gq (+ 8)
    const int col_base = warp_n * 32;         // + ni * 8 + 2 * tq (+ 1)
    uint32_t ring_pos = 0;
    uint32_t part_w[2] = {0, 0};
    uint32_t pxchg = 0;
    uint32_t staged = 0;  // chunks handed to the storer so far

    // per-lane ldmatrix offsets inside a tile (k16 step added later)
    const int a_row_in = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int a_kc = lane >> 4;  // which 16-byte k chunk of the k16 step
    const int b_row_in = (lane & 7) + (lane >> 4) * 8;
    const int b_kc = (lane >> 3) & 1;

    int g = 0;
    for (int rg = slot; rg < p.n_rowgroups; rg += p.slots) {
      // ---- load the flow state and the condition of this row group ----
      for (int i = tid; i < RT * kPad; i += kComputeThreads) {
        const int r = i / kPad, j = i % kPad;
        const int row = rg * RT + r;
        float uv = 0.f, cv = 0.f;
        if (row < p.batch) {
          if (j < p.W) uv = p.in[(size_t)row * p.in_ld + j];
          if (j < p.cond_cols) cv = p.cond[(size_t)(row % p.cond_rows) * p.cond_ld + j];
        }
        sm.u[r][j] = uv;
        if (j < 8) sm.cnd[r][j] = cv;
      }
      bar_compute();

      for (int blk = p.block_first; blk >= p.block_last; --blk) {
        for (int sidx = 0; sidx < 2; ++sidx, ++g) {
          const int sb = g & 1;
          const float* sp = sm.small[sb];
          mbar_wait(&sm.small_full[sb], (g >> 1) & 1);
          // subnet1 reads the first half and transforms the second; subnet2 the other way round
          const int in_off = sidx == 0 ? 0 : p.s1;
          const int in_len = sidx == 0 ? p.s1 : p.s2;
          const int tg_off = sidx == 0 ? p.s1 : 0;
          const int tg_len = sidx == 0 ? p.s2 : p.s1;

          // v[mi][ni][e]: activation (after bias + LeakyReLU) of row row_base + mi*16 + gq + 8*(e>>1),
          //               feature col_base + ni*8 + 2*tq + (e&1)  -- the mma accumulator fragment layout
          float v[MI][4][4];

          // ---- first layer: fp32 SIMT from the replicated state ----
          {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
              const float2 b2 = *reinterpret_cast<const float2*>(sp + kSmallFirstB + col_base + ni * 8 + 2 * tq);
#pragma unroll
              for (int mi = 0; mi < MI; ++mi) {
                v[mi][ni][0] = b2.x;
                v[mi][ni][1] = b2.y;
                v[mi][ni][2] = b2.x;
                v[mi][ni][3] = b2.y;
              }
            }
            const int kin = in_len + p.dim_cond;
            for (int k = 0; k < kin; ++k) {
              float x[MI][2];
#pragma unroll
              for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int r = row_base + mi * 16 + gq + 8 * h;
                  x[mi][h] = k < in_len ? sm.u[r][in_off + k] : sm.cnd[r][k - in_len];
                }
              const float* wr = sp + kSmallFirstW + k * kFT + col_base + 2 * tq;
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) {
                const float2 w2 = *reinterpret_cast<const float2*>(wr + ni * 8);
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                  v[mi][ni][0] = fmaf(x[mi][0], w2.x, v[mi][ni][0]);
                  v[mi][ni][1] = fmaf(x[mi][0], w2.y, v[mi][ni][1]);
                  v[mi][ni][2] = fmaf(x[mi][1], w2.x, v[mi][ni][2]);
                  v[mi][ni][3] = fmaf(x[mi][1], w2.y, v[mi][ni][3]);
                }
              }
            }
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
              for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[mi][ni][e] = leaky(v[mi][ni][e]);
          }

          for (int l = 0; l <= p.n_big; ++l) {
            if (l > 0) {
              // ---- hidden layer l-1: bf16x3 tensor-core tiles over the NT k-chunks ----
#pragma unroll
              for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                  for (int e = 0; e < 4; ++e) v[mi][ni][e] = 0.f;
              if (tid == 0) trace_ev(p, g * 4 + l - 1, 6);
              for (int i = 0; i < NT; ++i) {
                const int s = ring_pos % kStages;
                mbar_wait(&sm.full[s], (ring_pos / kStages) & 1);
                if (tid == 0 && i == 0) trace_ev(p, g * 4 + l - 1, 7);
                const uint32_t a_hi = smem_u32(sm.ring[s]);
                const uint32_t a_lo = a_hi + C::kATileBytes;
                const uint32_t w_hi = a_hi + C::kAChunkBytes;
                const uint32_t w_lo = w_hi + kWTileBytes;
#pragma unroll
                for (int kk = 0; kk < kKC / 16; ++kk) {
                  uint32_t ah[MI][4], al[MI][4], bh[2][4], bl[2][4];
#pragma unroll
                  for (int mi = 0; mi < MI; ++mi) {
                    const int row = row_base + mi * 16 + a_row_in;
                    const uint32_t off = row * 128 + ((((kk * 2 + a_kc) ^ (row & 7)) & 7) << 4);
                    ldsm_x4(a_hi + off, ah[mi]);
                    ldsm_x4(a_lo + off, al[mi]);
                  }
#pragma unroll
                  for (int nj = 0; nj < 2; ++nj) {
                    const int row = col_base + nj * 16 + b_row_in;
                    const uint32_t off = row * 128 + ((((kk * 2 + b_kc) ^ (row & 7)) & 7) << 4);
                    ldsm_x4(w_hi + off, bh[nj]);
                    ldsm_x4(w_lo + off, bl[nj]);
                  }
#pragma unroll
                  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
                      const int nj = ni >> 1, o = (ni & 1) * 2;
                      if (p.precision == IKF_PRECISION_BF16X3) {
                        mma_bf16(v[mi][ni], al[mi], bh[nj][o], bh[nj][o + 1]);
                        mma_bf16(v[mi][ni], ah[mi], bl[nj][o], bl[nj][o + 1]);
                      }
                      mma_bf16(v[mi][ni], ah[mi], bh[nj][o], bh[nj][o + 1]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[s]);
                ++ring_pos;
              }
              if (tid == 0) trace_ev(p, g * 4 + l - 1, 8);
              const float* bb = sp + kSmallBigB + (l - 1) * kFT + col_base + 2 * tq;
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) {
                const float2 b2 = *reinterpret_cast<const float2*>(bb + ni * 8);
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                  v[mi][ni][0] = leaky(v[mi][ni][0] + b2.x);
                  v[mi][ni][1] = leaky(v[mi][ni][1] + b2.y);
                  v[mi][ni][2] = leaky(v[mi][ni][2] + b2.x);
                  v[mi][ni][3] = leaky(v[mi][ni][3] + b2.y);
                }
              }
            }

            if (l < p.n_big) {
              // ---- publish: split into bf16 head/tail, stage the swizzled chunk, hand it to the storer ----
              if (staged > 0) mbar_wait(&sm.staging_free, (staged - 1) & 1);
#pragma unroll
              for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    const float f0 = v[mi][ni][2 * h], f1 = v[mi][ni][2 * h + 1];
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(f0), h1 = __float2bfloat16_rn(f1);
                    const __nv_bfloat16 l0 = __float2bfloat16_rn(f0 - __bfloat162float(h0));
                    const __nv_bfloat16 l1 = __float2bfloat16_rn(f1 - __bfloat162float(h1));
                    const uint32_t off = tile_off_bytes(row_base + mi * 16 + gq + 8 * h, col_base + ni * 8 + 2 * tq);
                    *reinterpret_cast<uint32_t*>(sm.staging + off) = pack_bf16(h0, h1);
                    *reinterpret_cast<uint32_t*>(sm.staging + C::kATileBytes + off) = pack_bf16(l0, l1);
                  }
              fence_proxy_async();
              bar_staged_arrive();
              if (tid == 0) trace_ev(p, g * 4 + l, 10);
              ++staged;
            }
          }

          // ---- last layer: this CTA's 64 features of every output, in fp32 straight from the activations ----
          if (tid == 0) trace_ev(p, g * 4 + 3, 11);
          const int pb = pxchg & 1;
          if (staged > 0) mbar_wait(&sm.staging_free, (staged - 1) & 1);  // the storer is done with the buffer
          float* wpart = reinterpret_cast<float*>(sm.staging);           // [2 warp_n][RT][16]
          {
#pragma unroll
            for (int o4 = 0; o4 < 4; ++o4) {
              float po[MI][2][4];
#pragma unroll
              for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                  for (int oo = 0; oo < 4; ++oo) po[mi][h][oo] = 0.f;
#pragma unroll
              for (int oo = 0; oo < 4; ++oo) {
                const float* wr = sp + kSmallLastW + (o4 * 4 + oo) * kFT + col_base + 2 * tq;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                  const float2 w2 = *reinterpret_cast<const float2*>(wr + ni * 8);
#pragma unroll
                  for (int mi = 0; mi < MI; ++mi) {
                    po[mi][0][oo] = fmaf(v[mi][ni][0], w2.x, po[mi][0][oo]);
                    po[mi][0][oo] = fmaf(v[mi][ni][1], w2.y, po[mi][0][oo]);
                    po[mi][1][oo] = fmaf(v[mi][ni][2], w2.x, po[mi][1][oo]);
                    po[mi][1][oo] = fmaf(v[mi][ni][3], w2.y, po[mi][1][oo]);
                  }
                }
              }
#pragma unroll
              for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                  for (int oo = 0; oo < 4; ++oo) {
                    float s = po[mi][h][oo];
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    po[mi][h][oo] = s;
                  }
                  if (tq == o4) {
                    const int r = row_base + mi * 16 + gq + 8 * h;
                    *reinterpret_cast<float4*>(&wpart[(warp_n * RT + r) * kPad + 4 * o4]) =
                        make_float4(po[mi][h][0], po[mi][h][1], po[mi][h][2], po[mi][h][3]);
                  }
                }
            }
          }
          bar_compute();
          for (int i = tid; i < RT * 4; i += kComputeThreads) {
            const int r = i >> 2, o4 = i & 3;
            const float4 x0 = *reinterpret_cast<const float4*>(&wpart[r * kPad + 4 * o4]);
            const float4 x1 = *reinterpret_cast<const float4*>(&wpart[(RT + r) * kPad + 4 * o4]);
            float* dst = part_slot + (((size_t)pb * NT + t) * kRTMax + r) * kPad + 4 * o4;
            __stcg(reinterpret_cast<float4*>(dst), make_float4(x0.x + x1.x, x0.y + x1.y, x0.z + x1.z, x0.w + x1.w));
          }
          bar_compute();
          const uint32_t pexp = p.epoch + 1 + part_w[pb];
          if (tid == 0) st_release(pflag + pb * NT + t, pexp);  // cumulative over the barrier: one fence per CTA
          if (warp == 0) {
            for (int c = lane; c < NT; c += 32) wait_flag(pflag + pb * NT + c, pexp, p.status, launch_id);
            __threadfence();  // acquire
          }
          bar_compute();
          if (tid == 0) trace_ev(p, g * 4 + 3, 12);
          for (int i = tid; i < RT * 4; i += kComputeThreads) {
            // fixed summation order over the team: every CTA obtains bitwise identical coefficients
            const int r = i >> 2, o4 = i & 3;
            float4 s = *reinterpret_cast<const float4*>(sp + kSmallLastB + 4 * o4);
            const float* src = part_slot + ((size_t)pb * NT * kRTMax + r) * kPad + 4 * o4;
            for (int c = 0; c < NT; ++c) {
              const float4 x = __ldcg(reinterpret_cast<const float4*>(src + (size_t)c * kRTMax * kPad));
              s.x += x.x;
              s.y += x.y;
              s.z += x.z;
              s.w += x.w;
            }
            *reinterpret_cast<float4*>(&sm.a[r][4 * o4]) = s;
          }
          // this subnet's small parameters are no longer needed
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.small_empty[sb]);
          ++part_w[pb];
          ++pxchg;
          bar_compute();
          // ---- affine coupling, reverse direction: y = (x - t) * exp(-clamp * 0.636 * atan(s)) ----
          for (int i = tid; i < RT * tg_len; i += kComputeThreads) {
            const int r = i / tg_len, j = i % tg_len;
            const float sc = p.clamp_scale * atanf(sm.a[r][j]);
            const float tr = sm.a[r][tg_len + j];
            sm.u[r][tg_off + j] = (sm.u[r][tg_off + j] - tr) * expf(-sc);
          }
          bar_compute();
          if (tid == 0) trace_ev(p, g * 4 + 3, 13);
        }
        // ---- PermuteRandom reverse: u = u[:, perm_inv] ----
        {
          constexpr int kPer = (RT * kPad + kComputeThreads - 1) / kComputeThreads;
          float tmp[kPer];
#pragma unroll
          for (int c = 0; c < kPer; ++c) {
            const int i = tid + c * kComputeThreads;
            tmp[c] = i < RT * p.W ? sm.u[i / p.W][p.perm_inv[blk * kPad + i % p.W]] : 0.f;
          }
          bar_compute();
#pragma unroll
          for (int c = 0; c < kPer; ++c) {
            const int i = tid + c * kComputeThreads;
            if (i < RT * p.W) sm.u[i / p.W][i % p.W] = tmp[c];
          }
          bar_compute();
        }
      }

      // ---- write this row group (team member 0 only; all replicas are identical) ----
      if (t == 0) {
        for (int i = tid; i < RT * p.out_cols; i += kComputeThreads) {
          const int r = i / p.out_cols, j = i % p.out_cols;
          const int row = rg * RT + r;
          if (row >= p.batch) continue;
          float o;
          if (p.finalize) {
            // FixedLinearTransform reverse (x - b) @ M_inv, slice, joint-limit clamp (ikflow_solver.py:98-102)
            o = 0.f;
            for (int k = 0; k < p.W; ++k) o = fmaf(sm.u[r][k] - p.flt_b[k], p.m_inv[k * kPad + j], o);
            if (p.clamp_out && j < p.ndof) o = fminf(fmaxf(o, p.lo[j]), p.hi[j]);
          } else {
            o = sm.u[r][j];
          }
          if (!isfinite(o)) atomicOr(p.status, IKF_STATUS_NONFINITE);
          p.out[(size_t)row * p.out_ld + j] = o;
        }
      }
      bar_compute();
    }
  }
}

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
  if (d->precision != IKF_PRECISION_BF16X3 && d->precision != IKF_PRECISION_BF16X1) return false;
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

struct IkfFlow {
  IkfFlowDesc desc;
  int device = 0;
  int num_sms = 0;
  int NT = 0, n_big = 0, slots_max = 0;
  void* blob = nullptr;  // one device allocation holding everything below
  size_t blob_bytes = 0, big_w_bytes = 0;
  ikf::FlowParams base;  // pointers + model constants; per-call fields filled at launch
  size_t smem32 = 0, smem64 = 0;
  uint32_t epoch = 1;  // doubles as launch id; 0 is "no launch aborted"
  int last_grid = 0;
  unsigned long long* trace = nullptr;
  int trace_layers = 0;
  size_t smem_bytes = 0;
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
  if (flow->blob) {
    DeviceGuard guard(flow->device);
    cudaFree(flow->blob);
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
  const int W = desc->ndim_tot, s1 = W / 2, s2 = W - s1, H = desc->hidden, NT = H / kFT;
  const int n_big = desc->coeff_fn_config - 1, nb = desc->nb_nodes, n_sub = 2 * nb;
  f->NT = NT;
  f->n_big = n_big;
  f->slots_max = std::max(1, kCtasPerSm * f->num_sms / NT);
  if (NT > f->num_sms) {
    delete f;
    return fail(IKF_EDEVICE, "ikf_flow_create: hidden=%d needs %d co-resident CTAs, device has %d SMs", H, NT, f->num_sms);
  }

  // ---- host-side repack ----
  const size_t big_elems = (size_t)n_sub * n_big * NT * NT * 2 * (kFT * kKC);
  const size_t small_floats = (size_t)n_sub * NT * kSmallFloats;
  std::vector<uint16_t> big(big_elems);
  std::vector<float> small(small_floats, 0.f);
  const float* wp = weights;
  for (int i = 0; i < nb; ++i) {
    for (int sidx = 0; sidx < 2; ++sidx) {
      const int n = 2 * i + sidx;
      const int in_dim = (sidx == 0 ? s1 : s2) + desc->dim_cond;
      const int out_dim = 2 * (sidx == 0 ? s2 : s1);
      // first Linear [H, in_dim], bias [H]
      const float* w0 = wp;
      const float* b0 = wp + (size_t)H * in_dim;
      wp = b0 + H;
      for (int f_ = 0; f_ < H; ++f_) {
        float* blk = small.data() + ((size_t)n * NT + f_ / kFT) * kSmallFloats;
        for (int k = 0; k < in_dim; ++k) blk[kSmallFirstW + k * kFT + f_ % kFT] = w0[(size_t)f_ * in_dim + k];
        blk[kSmallFirstB + f_ % kFT] = b0[f_];
      }
      // hidden Linears [H, H], bias [H]
      for (int l = 0; l < n_big; ++l) {
        const float* w = wp;
        const float* b = wp + (size_t)H * H;
        wp = b + H;
        for (int tt = 0; tt < NT; ++tt) {
          for (int c = 0; c < NT; ++c) {
            uint16_t* hi = big.data() + ((((size_t)n * n_big + l) * NT + tt) * NT + c) * 2 * (kFT * kKC);
            uint16_t* lo = hi + (kFT * kKC);
            for (int r = 0; r < kFT; ++r) {
              const float* src = w + (size_t)(tt * kFT + r) * H + c * kKC;
              for (int e = 0; e < kKC; ++e) {
                const uint32_t off = tile_off_bytes(r, e) / 2;
                const uint16_t hb = bf16_bits_rn(src[e]);
                hi[off] = hb;
                lo[off] = bf16_bits_rn(src[e] - bf16_bits_to_float(hb));
              }
            }
          }
        }
        for (int f_ = 0; f_ < H; ++f_)
          small[((size_t)n * NT + f_ / kFT) * kSmallFloats + kSmallBigB + l * kFT + f_ % kFT] = b[f_];
      }
      // last Linear [out_dim, H], bias [out_dim]
      const float* wl = wp;
      const float* bl = wp + (size_t)out_dim * H;
      wp = bl + out_dim;
      for (int tt = 0; tt < NT; ++tt) {
        float* blk = small.data() + ((size_t)n * NT + tt) * kSmallFloats;
        for (int o = 0; o < out_dim; ++o) {
          for (int e = 0; e < kFT; ++e) blk[kSmallLastW + o * kFT + e] = wl[(size_t)o * H + tt * kFT + e];
          blk[kSmallLastB + o] = bl[o];
        }
      }
    }
  }

  std::vector<int> perm(nb * kPad, 0);
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < W; ++j) {
      const int64_t v = perm_inv[(size_t)i * W + j];
      if (v < 0 || v >= W) {
        delete f;
        return fail(IKF_EINVAL, "ikf_flow_create: perm_inv[%d][%d]=%lld out of range", i, j, (long long)v);
      }
      perm[i * kPad + j] = (int)v;
    }
  std::vector<float> consts(kPad * kPad + 3 * kPad, 0.f);
  for (int i = 0; i < W; ++i)
    for (int j = 0; j < W; ++j) consts[i * kPad + j] = m_inv[(size_t)i * W + j];
  for (int j = 0; j < W; ++j) consts[kPad * kPad + j] = flt_b[j];
  for (int j = 0; j < desc->ndof; ++j) {
    consts[kPad * kPad + kPad + j] = joint_lo[j];
    consts[kPad * kPad + 2 * kPad + j] = joint_hi[j];
  }

  // ---- device blob ----
  auto align_up = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  const int slots = std::max(1, kCtasPerSm * f->num_sms / NT);  // upper bound; refined below from the occupancy
  const size_t off_big = 0;
  const size_t off_small = align_up(off_big + big_elems * 2);
  const size_t off_perm = align_up(off_small + small_floats * 4);
  const size_t off_consts = align_up(off_perm + perm.size() * 4);
  const size_t off_act = align_up(off_consts + consts.size() * 4);
  const size_t act_bytes = (size_t)slots * 2 * NT * kAChunkStride;
  const size_t off_partial = align_up(off_act + act_bytes);
  const size_t partial_bytes = (size_t)slots * 2 * NT * kRTMax * kPad * 4;
  const size_t off_flags = align_up(off_partial + partial_bytes);
  const size_t flag_bytes = ((size_t)slots * 2 * NT * 2 + 2) * 4;
  f->blob_bytes = align_up(off_flags + flag_bytes);
  f->big_w_bytes = big_elems * 2;
  cudaError_t e = cudaMalloc(&f->blob, f->blob_bytes);
  if (e != cudaSuccess) {
    delete f;
    return fail(IKF_ENOMEM, "ikf_flow_create: cudaMalloc(%zu) failed: %s", f->blob_bytes, cudaGetErrorString(e));
  }
  uint8_t* base = (uint8_t*)f->blob;
  e = cudaMemset(base + off_act, 0, f->blob_bytes - off_act);
  if (e == cudaSuccess && big_elems) e = cudaMemcpy(base + off_big, big.data(), big_elems * 2, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(base + off_small, small.data(), small_floats * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(base + off_perm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(base + off_consts, consts.data(), consts.size() * 4, cudaMemcpyHostToDevice);
  f->smem32 = sizeof(FlowSmem<32>) + 1024;
  f->smem64 = sizeof(FlowSmem<64>) + 1024;
  f->smem_bytes = f->smem64;
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(flow_inverse_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem32);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(flow_inverse_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem64);
  if (e == cudaSuccess) {
    // the teams spin on each other's flags, so every CTA of a launch must be resident: size the slot count from
    // what the device really fits (two CTAs per SM by design)
    int occ32 = 0, occ64 = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ32, flow_inverse_kernel<32>, kThreads, f->smem32);
    if (e == cudaSuccess)
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ64, flow_inverse_kernel<64>, kThreads, f->smem64);
    const int occ = std::min(occ32, occ64);
    if (e == cudaSuccess && occ < 1) {
      ikf_flow_destroy(f);
      return fail(IKF_EDEVICE, "ikf_flow_create: the flow kernel does not fit on an SM of device %d", device);
    }
    f->slots_max = std::max(1, std::min(occ, kCtasPerSm) * f->num_sms / NT);
  }
  if (e != cudaSuccess) {
    ikf_flow_destroy(f);
    return fail(IKF_ECUDA, "ikf_flow_create: device setup failed: %s", cudaGetErrorString(e));
  }

  FlowParams& p = f->base;
  std::memset(&p, 0, sizeof(p));
  p.W = W; p.s1 = s1; p.s2 = s2; p.dim_cond = desc->dim_cond; p.nb_nodes = nb; p.n_big = n_big; p.H = H; p.NT = NT;
  p.ndof = desc->ndof; p.precision = desc->precision;
  p.clamp_scale = (float)((double)desc->rnvp_clamp * 0.636);
  p.big_w = (const __nv_bfloat16*)(base + off_big);
  p.small = (const float*)(base + off_small);
  p.perm_inv = (const int*)(base + off_perm);
  p.m_inv = (const float*)(base + off_consts);
  p.flt_b = p.m_inv + kPad * kPad;
  p.lo = p.flt_b + kPad;
  p.hi = p.lo + kPad;
  p.act = base + off_act;
  p.partial = (float*)(base + off_partial);
  p.act_flag = (uint32_t*)(base + off_flags);
  p.part_flag = p.act_flag + (size_t)slots * 2 * NT;
  p.status = p.part_flag + (size_t)slots * 2 * NT;
  *out = f;
  return IKF_OK;
}

int ikf_flow_reserve(IkfFlow* flow, int max_batch) {
  if (!flow || max_batch < 0) return fail(IKF_EINVAL, "ikf_flow_reserve: bad arguments");
  return IKF_OK;  // the workspace is per team, not per row: nothing grows with the batch
}

static int flow_launch(IkfFlow* flow, const float* in, int in_ld, const float* cond, int cond_ld, int cond_rows,
                       int cond_cols, float* out, int out_ld, int out_cols, int batch, int block_first, int block_last,
                       int finalize, int clamp, void* stream, const char* name) {
  if (!flow) return fail(IKF_EINVAL, "%s: flow is NULL", name);
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

  FlowParams p = flow->base;
  p.in = in; p.in_ld = in_ld; p.cond = cond; p.cond_ld = cond_ld; p.cond_rows = cond_rows; p.cond_cols = cond_cols;
  p.out = out; p.out_ld = out_ld; p.out_cols = out_cols; p.batch = batch;
  p.block_first = block_first; p.block_last = block_last; p.finalize = finalize; p.clamp_out = clamp;
  // Row groups of 32 while that still fits in one wave of teams (more CTAs in flight, and a partner CTA on every SM
  // to compute while a team waits on an exchange), 64 beyond.
  const int rt = ((batch + 31) / 32 <= flow->slots_max) ? 32 : 64;
  p.n_rowgroups = (batch + rt - 1) / rt;
  p.slots = std::min(p.n_rowgroups, flow->slots_max);
  p.epoch = flow->epoch;
  p.trace = flow->trace;
  p.trace_layers = flow->trace_layers;
  // every flag of this launch stays below epoch + 1 + (row groups per slot) * (subnets) * (exchanges per subnet)
  const uint32_t rg_per_slot = (uint32_t)((p.n_rowgroups + p.slots - 1) / p.slots);
  flow->epoch += rg_per_slot * 2u * (uint32_t)(block_first - block_last + 1) * (uint32_t)(flow->n_big + 1) + 2u;

  const int grid = p.slots * flow->NT;
  flow->last_grid = grid;
  void* args[] = {(void*)&p};
  // cooperative launch = the driver guarantees that all CTAs of all teams are co-resident (the teams spin on each
  // other's flags)
  const void* fn = rt == 32 ? (const void*)flow_inverse_kernel<32> : (const void*)flow_inverse_kernel<64>;
  cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), args,
                                              rt == 32 ? flow->smem32 : flow->smem64, (cudaStream_t)stream);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail(IKF_ECUDA, "%s: launch failed: %s", name, cudaGetErrorString(e));
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

int ikf_flow_status(IkfFlow* flow, void* stream, uint32_t* status_out) {
  if (!flow || !status_out) return fail(IKF_EINVAL, "ikf_flow_status: bad arguments");
  DeviceGuard guard(flow->device);
  if (!guard.ok) return fail(IKF_ECUDA, "ikf_flow_status: cudaSetDevice(%d) failed", flow->device);
  uint32_t host[2] = {0, 0};
  IKF_CUDA(cudaMemcpyAsync(host, flow->base.status, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  IKF_CUDA(cudaMemsetAsync(flow->base.status, 0, sizeof(uint32_t), (cudaStream_t)stream));
  IKF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  *status_out = host[0];
  return IKF_OK;
}

int ikf_flow_debug_trace(IkfFlow* flow, unsigned long long* dev_stamps, int n_layers) {
  if (!flow || n_layers < 0) return fail(IKF_EINVAL, "ikf_flow_debug_trace: bad arguments");
  flow->trace = n_layers > 0 ? dev_stamps : nullptr;
  flow->trace_layers = n_layers;
  return IKF_OK;
}

int ikf_flow_info(IkfFlow* flow, size_t* packed_weight_bytes, int* grid_ctas_last, int* smem_bytes) {
  if (!flow) return fail(IKF_EINVAL, "ikf_flow_info: flow is NULL");
  if (packed_weight_bytes) *packed_weight_bytes = flow->big_w_bytes + (size_t)2 * flow->desc.nb_nodes * flow->NT * kSmallBytes;
  if (grid_ctas_last) *grid_ctas_last = flow->last_grid;
  if (smem_bytes) *smem_bytes = (int)flow->smem_bytes;
  return IKF_OK;
}

}  // extern "C"
