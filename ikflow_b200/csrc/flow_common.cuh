// Shared device-side pieces of the two flow engines (flow_mma.cuh, flow_umma.cuh): launch parameters, PTX wrappers
// (mbarrier, bulk TMA copies, flags), the swizzled tile layout.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

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

constexpr int kMaxPeers = 8;   // GPUs of one node
constexpr int kMaxFold = 16;   // blocks per launch whose PermuteRandom gathers are folded into the indexing (FlowParams::phys)
constexpr float kLeakySlope = 0.01f;  // nn.LeakyReLU() default, ikflow/model.py:74-83

struct FlowParams {
  int W, s1, s2, dim_cond, nb_nodes, n_big, H, NT, ndof, precision;
  float clamp_scale;  // rnvp_clamp * 0.636, rounded to fp32 the way torch rounds the Python scalar
  const __nv_bfloat16* big_w;  // [subnet][n_big][NT t][NT c][head|tail][64 f][64 k] swizzled
  const float* small;          // [subnet][NT t][kSmallFloats]
  const float* first_jit;      // umma engine: [subnet][KCH c][16 k][64 f] + [64] bias per chunk (just-in-time first layer)
  const int* perm_inv;         // [nb_nodes][kPad]
  const int* perm_fwd;         // [nb_nodes][kPad]: PermuteRandom.perm (forward pass)
  const float* m_fwd;          // [kPad][kPad]: FixedLinearTransform M, out_j = sum_i x_i m_fwd[i][j] + b_j (forward pass)
  const float* m_inv;          // [kPad][kPad]  out_j = sum_i (u_i - b_i) m_inv[i][j]
  const float* flt_b;          // [kPad]
  const float* lo;             // [kPad] joint limits
  const float* hi;
  uint8_t* act;         // [slot][2][NT c][kAChunkStride]: head [RT][64 k] then tail, swizzled
  float* partial;       // [slot][2][NT t][kRTMax r][16 o]
  uint32_t* act_flag;   // [slot][2][NT]
  uint32_t* part_flag;  // [slot][2][NT]
  uint32_t* status;     // [0] status bits, [1] id of the launch that aborted
  uint32_t* status_host;  // the same two words in mapped host memory: the host reads them without synchronising
  uint32_t epoch;       // sequence numbers of this launch start at epoch + 1
  unsigned long long* trace;  // debug: [cta < NT][layer][96] globaltimer stamps of team 0 (NULL = off)
  int trace_layers;
  // timing experiments only (IKFLOW_B200_DEBUG, tcgen05 engine); results may be garbage when non-zero.  Bits: 1 / 2 skip
  // the activation / weight copies, 4 per-chunk SM-clock stamps for the tracer, 32 / 64 / 128 / 256 L2 prefetch distance
  // 0 / 1 / 3 / 4 layers (default 2), 1024 no proxy fence and 2048 no arithmetic in the just-in-time first layer,
  // 4096 thread-per-feature instead of tiled exchanged first layer (valid results), 8192 / 16384 k-split: one split chunk less /
  // more in the first accumulator tile (valid results; measured +-0.2 %).
  int debug;
  const float* in;
  const float* cond;
  float* out;
  int in_ld, cond_ld, cond_rows, cond_cols, out_ld, out_cols;
  int batch, block_first, block_last, finalize, clamp_out, n_rowgroups, slots;
  int row_base;  // row 0 of this launch is row `row_base` of the caller's batch (the condition of row i is row (row_base + i) % cond_rows)
  // tcgen05 engine: CTAs per thread-block cluster (1 = no clusters).  A cluster holds the CTAs with the SAME feature tile
  // t of `cluster` neighbouring teams: they need the same weight chunks at the same time, so each loads 1/cluster of a
  // chunk and multicasts it to all of them (one L2 read instead of `cluster`).  CTA -> (team slot, feature tile t):
  //   c = blockIdx.x / cluster, r = blockIdx.x % cluster (= %cluster_ctarank), t = c % NT, slot = (c / NT) * cluster + r
  int cluster;
  // k-split pairs (flow_umma.cuh, Cfg::KS): cluster = 2, r = the CTA's half kh of the k-chunks and of the rows,
  // slot = c / NT
  int ksplit;
  // k-split: the LAST ks_private k-chunks of every hidden layer are multiplied by BOTH CTAs of a pair, each for its own
  // rows only, so that the hand-over of the split part hides behind them (0: plain split in two halves)
  int ks_private;
  // Fused gather of the batch-sharded solve (tcgen05 engine; n_peers = 0: off).  The ranks of one node each solve a
  // contiguous block of rows; instead of a collective after the kernel, the final epilogue stores its rows straight into
  // the gathered buffer of EVERY rank (peer-mapped pointers: NVLink stores), and the last CTA to finish raises this
  // rank's flag on every rank.
  float* peer_out[kMaxPeers];       // [rank r] base of r's gathered buffer [rows_total][peer_ld] (own rank included)
  uint32_t* peer_flag[kMaxPeers];   // [rank r] r's flag array [n_peers]: entry `peer_rank` = sequence number of my last shard
  int n_peers, peer_rank, peer_row0, peer_ld;
  uint32_t peer_seq;
  uint32_t* peer_counter;           // device counter of finished writer CTAs (monotonic)
  uint32_t peer_count_target;       // its value when the last writer CTA of this launch has counted itself
  // tcgen05 engine: the PermuteRandom gathers never move the flow state.  phys[i][j] = the column of the state array that
  // holds logical column j while the i-th block of this launch is evaluated, phys[n_blocks] = at the end (composed on the
  // host per launch; fold = 0 for launches of more than kMaxFold blocks: those gather physically)
  int fold;
  uint8_t phys[kMaxFold + 1][kPad];
  int forward;         // 1: x -> z with log-det (blocks block_last..block_first ascending), tcgen05 engine only
  float logdet_m;      // FixedLinearTransform.logDetM
  float* logdet_out;   // [batch] (forward pass)
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
// The same for waits that are expected to be long (producer warps waiting for a free stage): sleep between probes so
// that the spinning warp does not take issue slots from the warps that do the arithmetic on the same scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// The same copy delivered to the same shared-memory offset of every CTA of the cluster named in `cta_mask`; each
// destination's mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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
// generic-proxy writes to shared memory -> async-proxy (bulk store) reads of it; the narrow form does not have to
// order the global-memory traffic of the bulk copies that are in flight
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
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

// A wait gave up: mark the launch as aborted for every other waiter, on the device and (plain stores, mapped memory)
// for the host, which checks the word at the next call on the handle without any synchronisation.
__device__ __forceinline__ void report_timeout(uint32_t* status, uint32_t* status_host, uint32_t launch_id) {
  atomicOr(status, IKF_STATUS_SYNC_TIMEOUT);
  atomicExch(status + 1, launch_id);
  if (status_host != nullptr) {
    volatile uint32_t* h = status_host;
    h[1] = launch_id;
    h[0] = h[0] | IKF_STATUS_SYNC_TIMEOUT;
    __threadfence_system();
  }
}
__device__ __forceinline__ void report_nonfinite(uint32_t* status, uint32_t* status_host) {
  atomicOr(status, IKF_STATUS_NONFINITE);
  if (status_host != nullptr) {
    volatile uint32_t* h = status_host;
    h[0] = h[0] | IKF_STATUS_NONFINITE;
  }
}

__device__ __forceinline__ void report_range(uint32_t* status, uint32_t* status_host) {
  atomicOr(status, IKF_STATUS_RANGE);
  if (status_host != nullptr) {
    volatile uint32_t* h = status_host;
    h[0] = h[0] | IKF_STATUS_RANGE;
  }
}

// Wait until *flag has reached `expected` (wrap-safe).  Gives up (and makes every later wait of this launch give up)
// after about a second: the results are then garbage and IKF_STATUS_SYNC_TIMEOUT is reported, but the GPU is not hung.
// The caller issues the acquire fence.
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t expected, uint32_t* status, uint32_t launch_id,
                                          uint32_t* status_host = nullptr) {
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
          report_timeout(status, status_host, launch_id);
          return;
        }
      }
    }
  }
}

__device__ __forceinline__ void cta_coords(const FlowParams& p, int& slot, int& t, int& kh) {
  const int cs = p.cluster > 1 ? p.cluster : 1;
  const int c = blockIdx.x / cs, r = blockIdx.x % cs;
  t = c % p.NT;
  slot = p.ksplit ? c / p.NT : (c / p.NT) * cs + r;
  kh = p.ksplit ? r : 0;
}
// index of this CTA inside team 0 (the team the debug trace follows), or -1
__device__ __forceinline__ int trace_cta(const FlowParams& p) {
  int slot, t, kh;
  cta_coords(p, slot, t, kh);
  return slot == 0 ? t + p.NT * kh : -1;  // (k-split: the second CTA of every pair follows the first NT)
}

constexpr int kTraceEvents = 96;  // stamps per (CTA, layer): 0-15 phase events; per k-chunk i: 16+i landed (MMA warp), 32+i stage free
                                  // (loader), 48+i copies issued, 64+i expect_tx done, 80+i weight copy issued
// per-chunk stamps (events 16..63) use the SM clock: cheap to read, only compared inside one CTA
__device__ __forceinline__ void trace_clk(const FlowParams& p, int layer, int ev) {
  if (p.trace != nullptr && (p.debug & 4) && layer < p.trace_layers) {  // IKFLOW_B200_DEBUG=4: they perturb
    const int tc = trace_cta(p);
    if (tc >= 0) p.trace[((size_t)tc * p.trace_layers + layer) * kTraceEvents + ev] = (unsigned long long)clock64();
  }
}
__device__ __forceinline__ void trace_ev(const FlowParams& p, int layer, int ev) {
  if (p.trace != nullptr && layer < p.trace_layers) {
    const int tc = trace_cta(p);
    if (tc >= 0) {
      unsigned long long tns;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
      p.trace[((size_t)tc * p.trace_layers + layer) * kTraceEvents + ev] = tns;
    }
  }
}

// Explicit shared-space accesses.  The kernels reach their shared memory through a struct reference built from the
// aligned dynamic-smem base, and nvcc then emits GENERIC loads/stores (LD.E / ST.E) for it -- measured 5x slower than
// LDS/STS in the last-layer loop.  These wrappers take 32-bit shared addresses (smem_u32).
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, uint16_t v) {
  asm volatile("st.shared.u16 [%1], %0;" ::"h"(v), "r"(a) : "memory");
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

}  // namespace ikf
