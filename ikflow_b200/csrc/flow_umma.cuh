// Flow engine "umma": the hidden layers on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Same algorithm and exchange protocol as flow_mma.cuh, different tiling:
//   * CTA tile = 128 hidden features (UMMA M = 128, the WEIGHTS are the A operand: nn.Linear stores [out, in], i.e.
//     K-major rows) x RT batch rows (UMMA N = RT, the ACTIVATIONS are the B operand, also K-major), D[128][RT] fp32 in
//     TMEM: lane = feature, column = row;
//   * team = hidden/128 CTAs (8 for the released models), one CTA per SM;
//   * per 64-wide k-chunk: 4 k16 steps x 3 products (head*head, head*tail, tail*head), folded into 8 tcgen05.mma by
//     stacking the activation head and tail along N; issued by one elected thread, operands straight from the swizzled
//     ring stages (K-major SWIZZLE_128B descriptors), stage release and "accumulator ready" signalled by tcgen05.commit
//     on mbarriers;
//   * epilogue warps: tcgen05.ld (thread = feature), bias + LeakyReLU, head/tail split, publish; first and last layer
//     of every subnet in fp32 FMA as before.
// One kernel template, flow_inverse_umma_kernel<RT, JIT, F16, KS, PP>:
//   RT   rows of a row group (of a CTA's share of it): 32 / 64 / 128
//   JIT  just-in-time first layer (RT = 32)                                   -- Cfg, "JIT" below
//   F16  fp16x3 operand format (scaled fp16 tails, correction products in their own accumulator) instead of bf16x3
//   KS   k-split pairs: two CTAs (a cluster) share a 64-row group, each multiplies half of the k-chunks; the kernel of
//        batches up to 576 rows                                                -- Cfg, "KS"
//   PP   ping-pong: two independent row groups per CTA, the SIMT phases of one under the MMAs of the other; the kernel
//        of every larger batch (2 x 32 / 64 / 128 rows)                        -- Cfg, "PP"
// Requires hidden % 128 == 0 and hidden <= 1024.
#pragma once

#include <cuda_fp16.h>

#include <type_traits>

#include "flow_common.cuh"

namespace ikf {
namespace umma {

constexpr int kFTU = 128;                       // hidden features per CTA
constexpr int kWPlaneU = kFTU * kKC * 2;        // one bf16 plane of a weight chunk [128][64]: 16 KB
constexpr int kWChunkU = 2 * kWPlaneU;          // head + tail: 32 KB
constexpr int kRTMaxU = 128;                          // largest row group of this engine
constexpr int kAStrideU = 2 * 2 * kRTMaxU * kKC * 2;  // scratch bytes reserved per producer: 2 k-chunks x (head+tail): 64 KB
// Partial sums of the last layer travel in the "LL" format of NCCL's low-latency protocol: every 16-byte unit holds two
// values and two copies of the exchange's sequence number, {v0, seq, v1, seq}.  An aligned 8-byte pair is written and
// read as one piece, so a reader that finds both sequence numbers in place has the values: no flag, no fence, no
// staging -- the producers store from registers and the consumers poll the data itself.
constexpr int kPartUnits = kPad / 2;            // 16-byte units per row
constexpr int kPartRowBytes = kPartUnits * 16;  // 128
// per (subnet, feature tile) small fp32 parameters:
//   first_wT [16 k][128 f] | first_b [128] | big_b [kMaxBig][128] | last_w [16 o][128 f] (float4 slots swizzled, see
//   flow.cu) | last_b [16]
constexpr int kSmFirstW = 0;
constexpr int kSmFirstB = kSmFirstW + kPad * kFTU;
constexpr int kSmBigB = kSmFirstB + kFTU;
constexpr int kSmLastW = kSmBigB + kMaxBig * kFTU;
constexpr int kSmLastB = kSmLastW + kPad * kFTU;
constexpr int kSmallFloatsU = kSmLastB + kPad;  // 4624
constexpr int kSmallBytesU = kSmallFloatsU * 4;  // 18496
static_assert(kSmallBytesU % 16 == 0, "bulk copies move multiples of 16 bytes");
// "JIT" first layer (RT = 32): every CTA computes the first layer of a subnet for ALL hidden features, one 64-feature
// k-chunk at a time, straight into the activation slot of the ring stage the tensor core is about to read -- while the
// previous chunks are being multiplied.  No publish, no fence, no flag, no poll, no activation copy for that layer: one
// of the three team exchanges of a subnet disappears.  The first-layer weights arrive with the stage:
//   per (subnet, k-chunk): first_w [16 k][64 f] fp32 | first_b [64]
constexpr int kJitMaxK = 12;                       // widest first-layer input (state half + condition) the JIT loop takes
constexpr int kJitChunkFloats = kPad * kKC + kKC;  // 1088
constexpr int kJitChunkBytes = kJitChunkFloats * 4;  // 4352
static_assert(kJitChunkBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

// Operand formats of the hidden-layer products (IkfFlowDesc.precision).  Both split every fp32 operand x into a 16-bit
// head and tail and issue head*head + head*tail + tail*head with fp32 accumulation:
//   bf16x3  tail = bf16(x - head): 16 mantissa bits, the exponent range of fp32 (nothing can overflow or underflow)
//   fp16x3  tail = fp16((x - head) * 2^11): 22 mantissa bits -- as close to the fp32 reference as fp32 is to fp64 -- at
//           the price of fp16's range: |x| must stay below 65504 (hidden activations of a trained network are O(1);
//           beyond it the head is inf, the outputs NaN and IKF_STATUS_NONFINITE is raised).  The scaled tails make the two
//           correction products 2^11 too large, so they get their own accumulator (the second half of the TMEM tile,
//           which the stacked head|tail activation operand fills anyway) and the epilogue adds main + corr * 2^-11.
constexpr float kTailScaleF16 = 2048.f;

// (a0, a1) -> packed heads (a0 in the low half) and packed tails
template <bool F16>
__device__ __forceinline__ void split_pair(float a0, float a1, uint32_t& hb, uint32_t& lb) {
  if constexpr (F16) {
    const __half2 h2 = __floats2half2_rn(a0, a1);
    hb = *reinterpret_cast<const uint32_t*>(&h2);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn((a0 - hf.x) * kTailScaleF16, (a1 - hf.y) * kTailScaleF16);
    lb = *reinterpret_cast<const uint32_t*>(&l2);
  } else {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);  // .x = a0 (low half), one instruction
    hb = *reinterpret_cast<const uint32_t*>(&h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(a0 - __uint_as_float(hb << 16), a1 - __uint_as_float(hb & 0xffff0000u));
    lb = *reinterpret_cast<const uint32_t*>(&l2);
  }
}
// one value -> head in the low half, tail in the high half
template <bool F16>
__device__ __forceinline__ uint32_t split_one(float v) {
  if constexpr (F16) {
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn((v - __half2float(hi)) * kTailScaleF16);
    return (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
  } else {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    return pack_bf16(hi, lo);
  }
}

// KS ("k-split"): a ROW GROUP OF 64 is shared by TWO CTAs per feature tile, a cluster of 2.  CTA (t, kh) multiplies only
// the k-chunks of its half kh of every hidden layer (8 of 16: half the weight bytes through its shared memory, which is
// what paces the hidden layers) against all 64 rows, hands the partial sums of the other CTA's 32 rows over through
// distributed shared memory and finishes its own 32 rows: first layer, epilogue, publication, last layer, coupling and
// flow state are those of a 32-row CTA (RT = 32), at row offset 32 kh inside the team's 64-row exchange tiles.
// PP ("ping-pong", large batches): TWO independent 128-row groups per CTA.  Each has its own four epilogue warps, flow
// state, accumulator tile (TMEM columns 256 h ..), exchange slot and flags; loaders, MMA warp and the operand rings are
// shared and walk the layer jobs in the fixed order (step g, layer l, group h).  While the tensor core multiplies a layer
// of one group, the SIMT phases of the other (drain, publish, last layer, partial sums, coupling, first layer) run --
// phases during which the tensor core of the single-group kernels idles (half of the time at 128 rows).  A group's four
// warps take its 128 rows in two passes of 64.
template <int RT, bool JIT = false, bool KS = false, bool PP = false>
struct Cfg {
  static_assert(!JIT || RT == 32, "the just-in-time first layer is written for 32-row groups");
  static_assert(!KS || (RT == 32 && !JIT), "k-split pairs: 32 rows per CTA, exchanged first layer");
  static_assert(!PP || (!JIT && !KS), "ping-pong: two row groups per CTA, exchanged first layer");
  static constexpr bool kPP = PP;
  static constexpr int kPasses = (PP && RT > 64) ? RT / 64 : 1;  // passes of kEpiRows rows a group's warps make over its rows
  static constexpr int kStateRows = PP ? 2 * RT : RT;         // rows of flow state held by the CTA
  static constexpr int kDomThreads = PP ? 128 : (RT > 64 ? 256 : 128);  // epilogue threads that synchronise with each other
  static constexpr int kXRows = KS ? 2 * RT : RT;     // rows of the exchanged activation tiles = rows of the MMA
  static constexpr int kAPlane = kXRows * kKC * 2;    // one bf16 plane of an activation k-chunk [rows][64]
  static constexpr int kAChunk = 2 * kAPlane;        // head + tail
  static constexpr int kW1Off = kWChunkU + kAChunk;   // JIT: first-layer weights of the chunk's 64 features
  static constexpr int kStage = JIT ? ((kW1Off + kJitChunkBytes + 1023) / 1024) * 1024 : kW1Off;  // 40 KB (RT=32; JIT 45 KB) / 48 KB (RT=64)
  // JIT: the small-parameter block is loaded without its first-layer part
  static constexpr int kSmShift = JIT ? kSmBigB : 0;
  static constexpr int kSmFloats = kSmallFloatsU - kSmShift;
  static constexpr uint32_t kFullCount = JIT ? 2 : 1;  // arrivals per phase of a stage's full barrier
  // Epilogue threads: 128 per group (thread = TMEM lane = hidden feature); a group drains EPI_ROWS accumulator columns.
  // RT = 128 uses two groups (rows 0-63 and 64-127) that work side by side.
  static constexpr int kGroups = (RT > 64 || PP) ? 2 : 1;
  static constexpr int kEpiRows = PP ? (RT < 64 ? RT : 64) : RT / kGroups;
  static constexpr int kEpiWarps = 4 * kGroups;
  static constexpr int kEpiThreads = 32 * kEpiWarps;
  static constexpr int kStages = RT == 32 ? 4 : (RT == 64 ? 3 : 2);  // JIT kernel: stages of the unified ring
  // Kernels without the just-in-time first layer keep the two operands in SEPARATE rings: a stage of the unified ring
  // is 64 KB at 128 rows, two of them fill the shared memory, and with ONE chunk in flight every chunk costs a full L2
  // round trip (1800 cycles against 770 of tensor-core work).  Weights (32 KB per chunk, the same for every row-group
  // size) and activations (256 B per row and chunk) each get the depth the shared memory allows, each has its own
  // loader warp and its own full / empty barriers; the last layer's fp32 scratch (`vt`) lives in the activation ring,
  // which is idle while it is needed (nothing is exchanged between the end of a subnet's second hidden layer and the
  // publication of the next subnet's first layer).
  static constexpr bool kSplit = !JIT;
  static constexpr int kWStages = PP ? (RT == 128 ? 2 : (RT == 64 ? 3 : 4)) : ((RT == 128 || KS) ? 3 : 4);
  static constexpr int kAStages = RT == 128 ? 2 : ((RT == 64 || KS) ? 3 : 4);
  static constexpr int kRecvBytes = KS ? 2 * RT * kFTU * 4 : 16;  // k-split: the peer's partial sums of this CTA's rows, 2 x [32][128] fp32
  // loader warp w owns the ring stages s with s % kLoaders == w: its waits on a stage's barriers are then strictly in
  // order (a waiter may lag an mbarrier by at most one phase)
  // JIT: two loader warps (each owns two stages), so that the CTA stays at 11 warps: with 13 the register file grants
  // only 128 registers per thread and the epilogue spills
  // split rings: one weight loader warp, one activation loader warp (the lanes of a warp take the stages in turn: copies
  // of different threads are processed side by side)
  static constexpr int kLoaders = 2;
  static_assert(!JIT || kStages % kLoaders == 0, "every stage needs exactly one owner");
  static constexpr int kLoaderWarp0 = kEpiWarps;  // first loader warp
  // MMA-issuing warps: warp m takes the k-chunks i = m (mod kMmaWarps) of every hidden layer and accumulates them in its
  // own TMEM tile, the epilogue adds the tiles in a fixed order.  Measured with 2 (round 2, same box): NO gain at any
  // row-group size -- the hidden layers are not paced by the issuing thread but by the shared memory, which feeds the
  // tensor core (44 KB of operand reads per chunk at 32 rows) while the bulk copies write the next chunks into it (40 KB):
  // 84 KB / ~140 B/clk = 600 cycles per chunk against 430 of MMAs alone; 144 KB -> 1030 cycles at 128 rows (measured 840 /
  // 1280).  One warp it stays.
  static constexpr int kMmaWarps = 1;
  static constexpr int kMmaWarp = kEpiWarps + kLoaders;  // first of them
  // registers are granted per 4 warps: up to 12 warps leave 170 registers per thread, 13+ would leave 128
  // JIT: four more SIMT warps that only help with the just-in-time first layer (8 warps x 8 features per chunk)
  static constexpr int kHelpers = JIT ? 4 : 0;
  static constexpr int kHelperWarp0 = kMmaWarp + kMmaWarps;
  static constexpr int kGenWarps = kEpiWarps + kHelpers;
  static constexpr int kGenThreads = 32 * kGenWarps;
  static constexpr int kThreads = (kMmaWarp + kMmaWarps + kHelpers) * 32;
  static constexpr int kVtBytes = 32 * kFTU * 4;         // fp32 [32 rows][128 features] of one group: last-layer operand
  static constexpr int kAccCols = 2 * kXRows;  // one accumulator tile: D[:, 0:2 rows] (N-stacked products)
  static constexpr int kTmemCols = kMmaWarps * kAccCols <= 32 ? 32 : (kMmaWarps * kAccCols <= 64 ? 64 : (kMmaWarps * kAccCols <= 128 ? 128 : (kMmaWarps * kAccCols <= 256 ? 256 : 512)));
  static_assert(kMmaWarps * kAccCols <= 512, "TMEM has 512 columns");
  static_assert(kStage % 1024 == 0, "stages must keep the 1024-byte alignment of the swizzle atoms");
  static_assert(kAChunk % 1024 == 0, "stages must keep the 1024-byte alignment of the swizzle atoms");
  static_assert(JIT || PP || kGroups * kVtBytes <= kAStages * kAChunk, "the last layer's scratch must fit the activation ring");
};

template <int RT, bool JIT = false, bool KS = false, bool PP = false>
struct __align__(1024) Smem {
  using C = Cfg<RT, JIT, KS, PP>;
  // JIT kernel: unified ring [weights head|tail][activations head|tail][first-layer weights]; the others: split rings
  uint8_t ring[JIT ? C::kStages : 1][JIT ? C::kStage : 1024];
  uint8_t wring[JIT ? 1 : C::kWStages][JIT ? 1024 : kWChunkU];   // [head | tail] x [128 features][64 k]
  uint8_t aring[JIT ? 1 : C::kAStages][JIT ? 1024 : C::kAChunk];  // [head | tail] x [RT rows][64 k]
  // per epilogue group: fp32 activations [32 rows][128 features] for the last layer (float4 slots swizzled); the kernels
  // with split rings keep it in `aring` instead
  uint8_t vt[JIT ? C::kGroups : 1][(JIT || PP) ? C::kVtBytes : 1024];  // (ping-pong: ONE tile, taken in turn -- vt_lock)
  uint8_t recv[C::kRecvBytes];  // k-split: written by the peer CTA of the cluster (st.async), 2 x [128 features][32 rows] fp32, 16-byte units swizzled
  float small[2][C::kSmFloats];
  float u[C::kStateRows][kPad];  // flow state (ping-pong: group h at rows RT h ..)
  float cnd[C::kStateRows][8];
  float logdet[C::kStateRows];  // forward pass: log|det J| accumulated over the blocks
  uint8_t phys[kMaxFold + 1][kPad];  // FlowParams::phys
  // input of the current subnet [state half | condition | 0] at its start, output of its last layer at its end
  float a[C::kStateRows][kPad];
  uint64_t full[C::kStages], empty[C::kStages];
  uint64_t wfull[C::kWStages], wempty[C::kWStages], afull[C::kAStages], aempty[C::kAStages];  // split rings
  uint64_t w1full[C::kStages];   // JIT: the first-layer weights of the stage's chunk have landed
  uint64_t w1empty[C::kStages];  // JIT: ... and have been used (the stage's weight/activation areas may still be busy)
  uint64_t small_full[2], small_empty[2];
  uint64_t dfull[2], dempty[2];  // per group (ping-pong: two)
  int vt_lock;
  uint64_t rbar;   // k-split: "the peer's partial sums have arrived" (one phase per hidden layer, 32 KB of st.async bytes)
  uint64_t dpart[2];  // k-split: the accumulator tile of the first / second part of the split chunks is complete
  uint32_t tmem_base;
};

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 64 bf16 (128 B), 8-row swizzle atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);  // start address >> 4
  d |= (uint64_t)1 << 16;                  // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                  // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16 (format 1) or fp16 (format 0), both K-major, M x N
__device__ __host__ constexpr uint32_t make_idesc(int M, int N, bool f16 = false) {
  return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
          tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All tensor-core work of one 64-wide k-chunk in a single asm block.  A tcgen05.mma costs the same ~65-80 cycles for
// every N <= 128 (scripts/ubench/umma_rate.cu), so the three split-operand products are folded into TWO instructions
// per k16 step by stacking the activation head and tail along N (they are adjacent in the stage: rows 0..RT-1 head,
// RT..2RT-1 tail):   D[:, 0:2RT] += W_head * [A_head; A_tail]     (N = 2 RT)
//                    D[:, 0:RT]  += W_tail * A_head               (N = RT)
// and the epilogue adds the two halves of D.  Descriptors advance by one add per k16 step (+32 bytes = +2 in the
// 16-byte address field); the accumulate predicates are set up once.
// tmem_d2: where the W_tail * A_head product goes -- tmem_d (bf16x3: one sum) or tmem_d + RT columns (fp16x3: the
// accumulator of the scaled correction terms, see kTailScaleF16).
__device__ __forceinline__ void mma_chunk_x3(uint32_t tmem_d, uint32_t tmem_d2, uint32_t idesc_2n, uint32_t idesc_n, uint64_t wh,
                                             uint64_t wl, uint64_t a, uint32_t acc_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, p1;\n\t"
      ".reg .b64 wh, wl, a;\n\t"
      "setp.ne.b32 p0, %7, 0;\n\t"
      "setp.eq.b32 p1, %7, %7;\n\t"
      "mov.b64 wh, %4;\n\tmov.b64 wl, %5;\n\tmov.b64 a, %6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, a, %2, p0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], wl, a, %3, p1;\n\t"
      "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, a, %2, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], wl, a, %3, p1;\n\t"
      "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, a, %2, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], wl, a, %3, p1;\n\t"
      "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, a, %2, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], wl, a, %3, p1;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_d2), "r"(idesc_2n), "r"(idesc_n), "l"(wh), "l"(wl), "l"(a), "r"(acc_first)
      : "memory");
}
__device__ __forceinline__ void mma_chunk_x1(uint32_t tmem_d, uint32_t idesc, uint64_t wh, uint64_t ah, uint32_t acc_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, p1;\n\t"
      ".reg .b64 wh, ah;\n\t"
      "setp.ne.b32 p0, %4, 0;\n\t"
      "setp.eq.b32 p1, %4, %4;\n\t"
      "mov.b64 wh, %2;\n\tmov.b64 ah, %3;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, ah, %1, p0;\n\t"
      "add.s64 wh, wh, 2;\n\tadd.s64 ah, ah, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, ah, %1, p1;\n\t"
      "add.s64 wh, wh, 2;\n\tadd.s64 ah, ah, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, ah, %1, p1;\n\t"
      "add.s64 wh, wh, 2;\n\tadd.s64 ah, ah, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, ah, %1, p1;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(idesc), "l"(wh), "l"(ah), "r"(acc_first)
      : "memory");
}
// ... and on the barrier at the same offset in every CTA of the cluster named in `cta_mask`
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// one lane of the (converged) warp; ptxas recognises the pattern and emits the guarded code without per-thread loops
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// two accumulator slices with ONE wait: the second load's latency hides behind the first's.  _issue / _finish: other
// work (remote stores) can be put between the loads and the wait.
__device__ __forceinline__ void tmem_ld32x2_issue(uint32_t taddr0, uint32_t taddr1, uint32_t (&r)[32], uint32_t (&q)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr0));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
        "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
        "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
      : "r"(taddr1));
}
__device__ __forceinline__ void tmem_ld32x2_finish(const uint32_t (&r)[32], const uint32_t (&q)[32], float (&v)[32], float (&w)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  // (the values are consumed through volatile moves below the wait: nothing may be scheduled above it)
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    uint32_t a, b;
    asm volatile("mov.b32 %0, %1;" : "=r"(a) : "r"(r[i]));
    asm volatile("mov.b32 %0, %1;" : "=r"(b) : "r"(q[i]));
    v[i] = __uint_as_float(a), w[i] = __uint_as_float(b);
  }
}
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr0, uint32_t taddr1, float (&v)[32], float (&w)[32]) {
  uint32_t r[32], q[32];
  tmem_ld32x2_issue(taddr0, taddr1, r, q);
  tmem_ld32x2_finish(r, q, v, w);
}

__device__ __forceinline__ void st_ll(void* p, float v0, float v1, uint32_t seq) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%2};" ::"l"(p), "r"(__float_as_uint(v0)), "r"(seq), "r"(__float_as_uint(v1))
               : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void lds128_2x64(uint32_t a, uint64_t& v0, uint64_t& v1) {
  asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v0), "=l"(v1) : "r"(a));
}
__device__ __forceinline__ uint64_t lds64(uint32_t a) {
  uint64_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32u(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64u(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
template <int N>
__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }
// epilogue + helper warps of the just-in-time first layer
// (not inlined: the epilogue and the helper warps then arrive at ONE barrier instruction; compute-sanitizer's synccheck
// reports two bar.sync sites on one barrier as divergence, although the hardware does not care)
template <int N>
__device__ __noinline__ void bar_gen() { asm volatile("bar.sync 5, %0;" ::"n"(N) : "memory"); }
// the four warps of one generation group (0: epilogue warps, 1: helper warps)
__device__ __forceinline__ void bar_gen_group(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(6 + grp) : "memory"); }
// the 128 threads of one epilogue group
__device__ __forceinline__ void bar_group(int h) { asm volatile("bar.sync %0, 128;" ::"r"(3 + h) : "memory"); }
__device__ __forceinline__ void stg64(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void stg128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Just-in-time first layer, one 64-feature k-chunk, computed by a group of four warps: warp j of the group takes
// features 16j .. 16j+15 of all 32 rows.  Lane = (row quad q = lane / 4, feature quad fq = lane % 4): rows q, q+8, q+16,
// q+24 (inputs in registers) x features 16j + 4fq .. +3.  Per k ONE 128-bit shared load per lane (4 distinct addresses
// per warp) -- shared memory is busy feeding the tensor core while this runs, every access is slow, so the loop is
// written for few accesses, all issued up front.  fp32 FMAs (packed pairs) in the same order as the exchanged version:
// bitwise identical activations.  LeakyReLU, bf16 head/tail split, 8-byte stores into the swizzled operand tile (rows
// q + 8 rr: eight different swizzle phases per store instruction).
template <int KB>
__device__ __forceinline__ void jit_chunk_load(uint32_t stage_a, uint32_t w1off, int j, int lane, uint64_t (&w)[KB + 1][2]) {
  const uint32_t w1_a = stage_a + w1off + (16 * j + 4 * (lane & 3)) * 4;
  lds128_2x64(w1_a + (kPad * kKC) * 4, w[KB][0], w[KB][1]);  // bias
#pragma unroll
  for (int k = 0; k < KB; ++k) lds128_2x64(w1_a + (k * kKC) * 4, w[k][0], w[k][1]);
}
template <int KB, int APLANE, bool F16>
__device__ __forceinline__ void jit_chunk_compute(uint32_t stage_a, int j, int lane, const float (&x)[4][KB],
                                                  const uint64_t (&w)[KB + 1][2], uint32_t* status, uint32_t* status_host) {
  const int q = lane >> 2, fq = lane & 3;
  const int f0 = 16 * j + 4 * fq;
  uint64_t acc[4][2];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) acc[rr][0] = w[KB][0], acc[rr][1] = w[KB][1];
#pragma unroll
  for (int k = 0; k < KB; ++k) {
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const uint64_t xk = pack2(x[rr][k], x[rr][k]);
      acc[rr][0] = ffma2(xk, w[k][0], acc[rr][0]);
      acc[rr][1] = ffma2(xk, w[k][1], acc[rr][1]);
    }
  }
  float vmax = 0.f;  // fp16x3: largest magnitude seen (the fp16 head of anything above 65504 is inf); checked once per chunk
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    uint32_t hb[2], lb[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float a0, a1;
      unpack2(acc[rr][c], a0, a1);
      a0 = leaky(a0), a1 = leaky(a1);
      if (F16) vmax = fmaxf(vmax, fmaxf(fabsf(a0), fabsf(a1)));
      split_pair<F16>(a0, a1, hb[c], lb[c]);
    }
    const uint32_t off = tile_off_bytes(q + 8 * rr, f0);
    sts64u(stage_a + kWChunkU + off, hb[0], hb[1]);
    sts64u(stage_a + kWChunkU + APLANE + off, lb[0], lb[1]);
  }
  if (F16 && vmax > 65504.f) report_range(status, status_host);
}

// Exchanged first layer (kernels without the just-in-time version), one tile of 32 rows x 16 features per call, same
// lane mapping and arithmetic as the JIT loop (lane = row quad x feature quad, inputs in registers, one 128-bit weight
// load per k and lane, packed fp32 FMAs in the original order): the head/tail halves go straight to the producer's slot
// of the global scratch ring with 8-byte stores.  The thread-per-feature version it replaces spent most of its time on
// broadcast 128-bit loads of the inputs (a quarter warp per bank phase): 7.2 us per subnet at 128 rows.
template <int KB, int APLANE, bool F16>
__device__ __forceinline__ void first_layer_tile(uint32_t w_a, uint32_t b_a, const float (&x)[4][KB], int rb, int fb, int lane,
                                                 uint8_t* dst_chunk0, int chunk_bytes, uint32_t* status, uint32_t* status_host) {
  const int q = lane >> 2, fq = lane & 3;
  const int f0 = fb * 16 + 4 * fq;  // feature inside the CTA's 128
  uint64_t acc[4][2];
  lds128_2x64(b_a + f0 * 4, acc[0][0], acc[0][1]);
#pragma unroll
  for (int rr = 1; rr < 4; ++rr) acc[rr][0] = acc[0][0], acc[rr][1] = acc[0][1];
#pragma unroll
  for (int k = 0; k < KB; ++k) {
    uint64_t w01, w23;
    lds128_2x64(w_a + (k * kFTU + f0) * 4, w01, w23);
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const uint64_t xk = pack2(x[rr][k], x[rr][k]);
      acc[rr][0] = ffma2(xk, w01, acc[rr][0]);
      acc[rr][1] = ffma2(xk, w23, acc[rr][1]);
    }
  }
  uint8_t* dst = dst_chunk0 + (size_t)(f0 >> 6) * chunk_bytes;
  float vmax = 0.f;  // see jit_chunk_compute
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    uint32_t hb[2], lb[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float a0, a1;
      unpack2(acc[rr][c], a0, a1);
      a0 = leaky(a0), a1 = leaky(a1);
      if (F16) vmax = fmaxf(vmax, fmaxf(fabsf(a0), fabsf(a1)));
      split_pair<F16>(a0, a1, hb[c], lb[c]);
    }
    const uint32_t off = tile_off_bytes(rb * 32 + q + 8 * rr, f0 & 63);
    stg64(dst + off, hb[0], hb[1]);
    stg64(dst + APLANE + off, lb[0], lb[1]);
  }
  if (F16 && vmax > 65504.f) report_range(status, status_host);
}

template <int RT, bool JIT = false, bool F16 = false, bool KS = false, bool PP = false>
__global__ void __launch_bounds__(Cfg<RT, JIT, KS, PP>::kThreads, 1) flow_inverse_umma_kernel(const FlowParams p) {
  using C = Cfg<RT, JIT, KS, PP>;
  constexpr int XR = C::kXRows;  // rows of a row group (of the exchanged tiles, of the MMAs); RT = rows this CTA finishes
  constexpr int kStages = C::kStages;
  constexpr int ER = C::kEpiRows, ET = C::kEpiThreads;
  // Accumulator tiles per hidden layer: k-chunk i goes to tile i % kAcc, the epilogue adds the tiles in a fixed order.
  // The tensor core TRUNCATES its fp32 accumulator after every k16 step (a CPU emulation with round-toward-zero reproduces
  // the measured errors, scripts/precision_study.py), a bias that grows with the number of sequential steps: 64 per layer
  // with one tile.  bf16x3 does not notice (its 16-bit operands dominate its error); fp16x3, whose operands carry 22 bits,
  // is limited by exactly this, so it spreads the chunks over 4 tiles (2 at 128 rows: TMEM has 512 columns) and sums
  // them in fp32 round-to-nearest.
  // k-split: always two tiles, the first and the second half of the CTA's chunks (the first half is handed over to the peer
  // while the second is still being multiplied); 8 chunks per layer and CTA -> the same 16 steps per tile as above
  constexpr int kAcc = KS ? 2 : (PP ? (F16 ? (XR == 128 ? 1 : (XR == 64 ? 2 : 4)) : 1) : (F16 ? (XR == 128 ? 2 : 4) : 1));  // ping-pong: the other half of TMEM belongs to the other group
  constexpr int NG = PP ? 2 : 1;  // independent row groups per CTA
  constexpr int kTmemColsK = KS ? 512 : PP ? (2 * kAcc * C::kAccCols <= 128 ? 128 : (2 * kAcc * C::kAccCols <= 256 ? 256 : 512)) : kAcc * C::kAccCols <= 32 ? 32 : (kAcc * C::kAccCols <= 64 ? 64 : (kAcc * C::kAccCols <= 128 ? 128 : (kAcc * C::kAccCols <= 256 ? 256 : 512)));
  static_assert(NG * kAcc * C::kAccCols <= 512 && C::kMmaWarps == 1, "TMEM has 512 columns; the tiles are dealt by chunk index");
  extern __shared__ uint8_t smem_raw[];
  Smem<RT, JIT, KS, PP>& sm = *reinterpret_cast<Smem<RT, JIT, KS, PP>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int NT = p.NT;            // hidden / 128
  const int KCH = p.H / kKC;      // 64-wide k-chunks per hidden layer (2 per producer)
  int slot, t, kh;                // k-split: kh = this CTA's half of the k-chunks = its half of the row group's rows
  cta_coords(p, slot, t, kh);
  // k-split with private chunks (FlowParams::ks_private = KP > 0): chunks [0, KCH - KP) are split between the two CTAs of
  // the pair (KSH each, multiplied against all 64 rows, partial sums of the peer's rows handed over), chunks [KCH - KP,
  // KCH) are multiplied by BOTH CTAs against their own 32 rows only -- while the hand-over is on its way
  const int KP = KS ? p.ks_private : 0;
  const int KSH = KS ? (KCH - KP) / 2 : 1;  // (1 without k-split: only ever a divisor there)
  const int KCHL = KS ? KSH + KP : KCH;  // k-chunks this CTA multiplies per hidden layer
  // i-th chunk of this CTA's layer job -> k-chunk index (fixed consumption order, rotated by the tile index so that the
  // CTAs of a team do not all pull the same producer's chunk first); i >= KSH: a private chunk
  auto ks_chunk = [&](int i) { return i < KSH ? KSH * kh + (2 * t + i) % KSH : 2 * KSH + (2 * t + (i - KSH)) % (KP > 0 ? KP : 1); };
  const int xrow0 = KS ? RT * kh : 0;   // this CTA's rows inside the row group / the exchanged tiles
  const int FW = KS ? 2 : 1;            // publication flags per feature tile
  // CTAs per cluster: same t, neighbouring teams (FlowParams::cluster).  Read from the parameter bank where needed
  // (the kernel is at its register limit: no function-scope copies).
#define IKF_CS (p.cluster > 1 ? p.cluster : 1)
  const uint32_t launch_id = p.epoch;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], C::kFullCount);
      mbar_init(&sm.empty[s], IKF_CS);  // clusters: a stage is refilled by multicast, so it must be free in EVERY CTA of the cluster
      mbar_init(&sm.w1full[s], 1);
      mbar_init(&sm.w1empty[s], 1);
    }
    for (int s = 0; s < C::kWStages; ++s) {
      mbar_init(&sm.wfull[s], 1);
      mbar_init(&sm.wempty[s], KS ? 1 : IKF_CS);  // (a k-split pair is a cluster too, but its CTAs load different chunks)
    }
    for (int s = 0; s < C::kAStages; ++s) {
      mbar_init(&sm.afull[s], 1);
      mbar_init(&sm.aempty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.small_full[b], 1);
      mbar_init(&sm.small_empty[b], C::kEpiWarps);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.dfull[i], C::kMmaWarps);
      mbar_init(&sm.dempty[i], PP ? C::kEpiWarps / 2 : C::kEpiWarps);
    }
    sm.vt_lock = 0;
    mbar_init(&sm.rbar, 1);
    if (KS) mbar_arrive_expect_tx(&sm.rbar, C::kRecvBytes);  // first phase: two hand-overs of [32 rows][128 features] fp32
    mbar_init(&sm.dpart[0], 1);
    mbar_init(&sm.dpart[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == C::kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                 "n"(kTmemColsK)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();  // the peers' barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  // exchange scratch and flags of group h's (virtual) team slot; ping-pong: team `slot` runs the slots 2 slot, 2 slot + 1
  auto act_slot_of = [&](int h) { return p.act + (size_t)(PP ? 2 * slot + h : slot) * 2 * NT * kAStrideU; };
  auto part_slot_of = [&](int h) {
    return reinterpret_cast<uint8_t*>(p.partial) + (size_t)(PP ? 2 * slot + h : slot) * 2 * NT * kRTMaxU * kPartRowBytes;
  };
  auto aflag_of = [&](int h) { return p.act_flag + (size_t)(PP ? 2 * slot + h : slot) * 2 * NT * FW; };  // [2 buffers][NT tiles][FW]

  const int n_blocks = p.block_first - p.block_last + 1;
  const int steps_per_rg = 2 * n_blocks;
  // clusters keep their CTAs in lock step (the weight stream is shared): every team then walks the same number of row
  // groups, the last ones possibly empty (rows >= batch read as zeros and are not written)
  // (ping-pong: both groups of every CTA walk the same number of row groups, for the shared loaders' and the MMA warp's sake)
  const int my_rgs = PP ? (p.n_rowgroups + 2 * p.slots - 1) / (2 * p.slots)
                        : p.cluster > 1 ? (p.n_rowgroups + p.slots - 1) / p.slots : (p.n_rowgroups - slot + p.slots - 1) / p.slots;
  const int total_steps = my_rgs * steps_per_rg;

  // Just-in-time first layer of subnet step g: generation warp j (epilogue warps 0-3, helper warps 4-7) runs over the
  // layer's k-chunks in ring order.  Called with the subnet's input in sm.a (bar_gen before the call).
  // Subnet evaluated at step g of a row group.  Reverse pass (z -> x): blocks block_first .. block_last descending,
  // subnet1 then subnet2; forward pass (x -> z, FrEIA GLOWCouplingBlock.forward): blocks ascending, subnet2 then subnet1.
  auto step_block = [&](int g) { return p.forward ? p.block_last + (g % steps_per_rg) / 2 : p.block_first - (g % steps_per_rg) / 2; };
  auto step_sidx = [&](int g) { return p.forward ? 1 - (g & 1) : (g & 1); };  // steps_per_rg is even
  auto step_subnet = [&](int g) { return 2 * step_block(g) + step_sidx(g); };

  auto jit_layer_loop_k = [&](auto kb, int grp, int j, int g) {
    constexpr int KB = decltype(kb)::value;  // k extent (inputs and weights are zero-padded, fma(0, 0, acc) leaves acc as it is)
    float xx[4][KB];                         // the inputs of this lane's four rows (lane / 4 + 8 rr)
    const uint32_t xin = smem_u32(&sm.a[0][0]);
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
      for (int k4 = 0; k4 < KB; k4 += 4) {
        const float4 xv = lds128(xin + (((lane >> 2) + 8 * rr) * kPad + k4) * 4);
        xx[rr][k4] = xv.x, xx[rr][k4 + 1] = xv.y, xx[rr][k4 + 2] = xv.z, xx[rr][k4 + 3] = xv.w;
      }
    // the two groups take the chunks alternately: two chunks are always in the making, which hides the latency of
    // the (busy) shared memory
    // cheap per-chunk stamps (IKFLOW_B200_DEBUG=4): row 4g+2 (group 0) / 4g+3... of the trace, SM clock
    unsigned long long* tb = nullptr;
    if (p.trace != nullptr && (p.debug & 4) && slot == 0 && g * 4 + 2 < p.trace_layers && j == 0 && lane == 0)
      tb = p.trace + ((size_t)t * p.trace_layers + g * 4 + 2) * kTraceEvents + grp * 8;
    for (int i = grp; i < KCH; i += 2) {
      const int st = i % kStages;  // every layer starts at ring stage 0 (KCH % kStages == 0, checked at launch)
      const uint32_t stage_a = smem_u32(sm.ring[st]);
      if (tb) tb[16 + (i >> 1)] = clock64();
      // 1. the chunk's first-layer weights have landed: pull them into registers ...
      mbar_wait(&sm.w1full[st], ((uint32_t)g * (uint32_t)(KCH / kStages) + (uint32_t)(i / kStages)) & 1u);
      uint64_t w[KB + 1][2];
      jit_chunk_load<KB>(stage_a, C::kW1Off, j, lane, w);
      // 2. ... while waiting for the tensor core to let go of the stage's activation area (previous use of the stage)
      const uint32_t use = ((uint32_t)g * (uint32_t)p.n_big * (uint32_t)KCH + (uint32_t)i) / kStages;
      if (use > 0) mbar_wait(&sm.empty[st], (use - 1) & 1);
      if (tb) tb[32 + (i >> 1)] = clock64();
      if (!(p.debug & 2048)) jit_chunk_compute<KB, C::kAPlane, F16>(stage_a, j, lane, xx, w, p.status, p.status_host);
      if (tb) tb[48 + (i >> 1)] = clock64();
      if (!(p.debug & 1024)) fence_proxy_async_smem();  // generic-proxy writes -> the tensor core's (async proxy) reads
      bar_gen_group(grp);
      if (tb) tb[64 + (i >> 1)] = clock64();
      if (j == 0 && lane == 0) {
        mbar_arrive(&sm.full[st]);
        mbar_arrive(&sm.w1empty[st]);
        if (i == 0) trace_ev(p, g * 4, 4);
      }
    }
  };
  // Just-in-time first layer of subnet step g: warp j of generation group grp (0: epilogue warps, 1: helper warps).
  // Called with the subnet's input in sm.a (bar_gen before the call).
  // (first-layer inputs wider than kJitMaxK do not fit the register budget of this loop: such models take the
  // exchanged first layer, see flow.cu)
  auto jit_layer_loop = [&](int grp, int j, int g, int /*kin*/) { jit_layer_loop_k(std::integral_constant<int, kJitMaxK>{}, grp, j, g); };

  if (JIT && warp >= C::kHelperWarp0) {
    // ===== helper warps: nothing but the just-in-time first layer =====
    if constexpr (JIT) {
      for (int g = 0; g < total_steps; ++g) {
        const int in_len = step_sidx(g) == 0 ? p.s1 : p.s2;  // steps alternate between the two subnets of a block
        bar_gen<C::kGenThreads>();
        jit_layer_loop(1, warp - C::kHelperWarp0, g, in_len + p.dim_cond);
      }
    }
  } else if (warp >= C::kLoaderWarp0 && warp < C::kLoaderWarp0 + C::kLoaders) {
    // ===== loaders: bulk-TMA producers.  k-chunk number pos = ring_pos + i (counted over the whole launch) goes to ring
    // stage pos % kStages and is loaded by loader warp pos % kLoaders; warp 0 also prefetches the small parameters and
    // pulls the next layer's weights into L2.  Chunks are consumed in the fixed order kc = (2t + i) % KCH, so the
    // accumulation order never depends on timing.
    const int lw = warp - C::kLoaderWarp0;
    uint32_t ring_pos = 0;
    uint32_t act_w[2] = {0, 0};
    uint32_t xchg = 0;
    auto prefetch_small = [&](int g) {
      if (g >= total_steps) return;
      const int b = g & 1;
      if (g >= 2) mbar_wait_relaxed(&sm.small_empty[b], ((g >> 1) - 1) & 1);
      if (lane == 0) {
        const int n = step_subnet(g);
        mbar_arrive_expect_tx(&sm.small_full[b], C::kSmFloats * 4);
        bulk_g2s(sm.small[b], p.small + ((size_t)n * NT + t) * kSmallFloatsU + C::kSmShift, C::kSmFloats * 4, &sm.small_full[b]);
      }
      __syncwarp();
    };
    // Hidden-layer weights into L2, two layers ahead of their use.  A weight slice is read by the CTAs t of ALL teams at
    // about the same time, so the teams share the work: team `slot` prefetches the k-chunks k = slot (mod slots) of its
    // CTAs' slices.  (Every CTA prefetching its whole slice -- 512 KB per layer through the SM's bulk-copy engine, next
    // to 640 KB of real loads -- delayed the loads: DRAM-sourced weights cost 4.5 us per subnet against L2-resident ones.)
    auto prefetch_layer = [&](int q) {
      const int g2 = q / p.n_big, l2 = q % p.n_big;
      if (g2 >= total_steps) return;
      const int n2 = step_subnet(g2);
      const uint8_t* wnext =
          reinterpret_cast<const uint8_t*>(p.big_w) + (((size_t)n2 * p.n_big + l2) * NT + t) * KCH * kWChunkU;
      for (int k = lane; k < KCH; k += 32)
        if (k % p.slots == slot && (!KS || (k < 2 * KSH ? k / KSH == kh : (k & 1) == kh)))  // (k-split: own half of the split part, every other private chunk)
          bulk_prefetch_l2(wnext + (size_t)k * kWChunkU, kWChunkU);
    };
    if constexpr (!JIT) {
      // ===== split rings (see Cfg::kSplit): loader warp 0 streams the weights, loader warp 1 the exchanged activations;
      // chunk number pos (counted over the whole launch) goes to stage pos % depth of its ring, issued by lane = stage =====
      if (lw == 0) {
        prefetch_small(0);
        uint32_t pos = 0;
        for (int g = 0; g < total_steps; ++g) {
          const int n = step_subnet(g);
          for (int l = 0; l < p.n_big; ++l) {
            const uint8_t* wbase =
                reinterpret_cast<const uint8_t*>(p.big_w) + (((size_t)n * p.n_big + l) * NT + t) * KCH * kWChunkU;
            for (int hh = 0; hh < NG; ++hh)  // ping-pong: the layer's weights once per group (the groups are out of phase)
            for (int i = 0; i < KCHL; ++i, ++pos) {
              const int st = pos % C::kWStages;
              const uint32_t use = pos / C::kWStages;
              // fixed consumption order (the accumulation order never depends on timing), starting at the CTA's own chunks;
              // k-split: inside this CTA's half of the layer
              const int kc = KS ? ks_chunk(i) : (2 * t + i) % KCH;
              if (use > 0) mbar_wait_relaxed(&sm.wempty[st], (use - 1) & 1);  // clusters: free in EVERY CTA of the cluster
              if (p.trace != nullptr && lane == 0 && i < 16) trace_clk(p, g * 4 + l, 32 + i);
              if (lane == st) {
                mbar_arrive_expect_tx(&sm.wfull[st], kWChunkU);
                if (!KS && p.cluster > 1) {  // this CTA's share of the chunk, delivered to all CTAs of the cluster
                  const uint32_t share = (uint32_t)kWChunkU / (uint32_t)p.cluster;
                  const uint32_t off = (blockIdx.x % (uint32_t)p.cluster) * share;
                  bulk_g2s_multicast(sm.wring[st] + off, wbase + (size_t)kc * kWChunkU + off, share, &sm.wfull[st], (uint16_t)((1u << p.cluster) - 1u));
                } else {
                  bulk_g2s(sm.wring[st], wbase + (size_t)kc * kWChunkU, kWChunkU, &sm.wfull[st]);
                }
                if (i == 0 && hh == 0) trace_ev(p, g * 4 + l, 0);
              }
              __syncwarp();
              if (p.trace != nullptr && lane == 0 && i < 16) trace_clk(p, g * 4 + l, 80 + i);
              if (i == 0 && hh == 0) {  // after the layer's first copy is on its way: next layers into L2
                const int q = g * p.n_big + l;
                const int dist = (p.debug & 32) ? 0 : (p.debug & 64) ? 1 : (p.debug & 128) ? 3 : (p.debug & 256) ? 4 : 2;
                if (q == 0) for (int d = 1; d < dist; ++d) prefetch_layer(d);
                if (dist > 0) prefetch_layer(q + dist);
              }
            }
            if (l == 0) prefetch_small(g + 1);
          }
          if (p.n_big == 0) prefetch_small(g + 1);
        }
      } else {
        uint32_t pos = 0;
        bool gave_up = false;
        uint32_t act_wg[2][2] = {{0, 0}, {0, 0}}, xchgg[2] = {0, 0};  // per group
        for (int g = 0; g < total_steps; ++g) {
          for (int l = 0; l < p.n_big; ++l)
          for (int hh = 0; hh < NG; ++hh) {
            const int buf = xchgg[hh] & 1;
            const uint32_t expected = p.epoch + 1 + act_wg[hh][buf];
            const uint8_t* abase = act_slot_of(hh) + (size_t)buf * NT * kAStrideU;
            const uint32_t* my_flag = aflag_of(hh) + buf * NT * FW + (lane < NT * FW ? lane : 0);
            uint32_t ready = gave_up ? 0xffffffffu : 0u;  // bit: that producer CTA has published (warp-uniform); k-split: two per tile
            for (int i = 0; i < KCHL; ++i, ++pos) {
              const int st = pos % C::kAStages;
              const uint32_t use = pos / C::kAStages;
              const int kc = KS ? ks_chunk(i) : (2 * t + i) % KCH;
              const int c = kc >> 1;
              const bool priv = KS && i >= KSH;  // a private chunk: this CTA's own 32 rows only
              // k-split: both halves of the producer tile's rows (a private chunk: the half that holds this CTA's rows)
              const uint32_t need = KS ? (priv ? (1u << (2 * c + kh)) : (3u << (2 * c))) : (1u << c);
              if (use > 0) mbar_wait_relaxed(&sm.aempty[st], (use - 1) & 1);
              if (p.trace != nullptr && lane == 0 && i < 16) trace_clk(p, g * 4 + l, 64 + i);
              uint32_t spins = 0;
              long long t0 = 0;
              while ((ready & need) != need) {
                bool ok = false;
                if (lane < NT * FW && !((ready >> lane) & 1u)) ok = (int32_t)(ld_relaxed(my_flag) - expected) >= 0;
                ready |= __ballot_sync(0xffffffffu, ok);
                if ((ready & need) == need) break;
                ++spins;
                if (spins == 64) t0 = clock64();
                if (spins > 64) {
                  __nanosleep(20);
                  if ((spins & 255u) == 0) {
                    int bail = 0;
                    if (lane == 0) {
                      if (ld_relaxed(p.status + 1) == launch_id) bail = 1;
                      else if (clock64() - t0 > 2500000000LL) {
                        report_timeout(p.status, p.status_host, launch_id);
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
              if (lane == st) {
                const uint8_t* src = abase + (size_t)c * kAStrideU + (size_t)(kc & 1) * C::kAChunk;
                if (priv) {
                  // the own rows of the head and of the tail plane, stacked in the stage as [head 32 rows | tail 32 rows]
                  mbar_arrive_expect_tx(&sm.afull[st], C::kAChunk / 2);
                  bulk_g2s(sm.aring[st], src + (size_t)xrow0 * (kKC * 2), C::kAPlane / 2, &sm.afull[st]);
                  bulk_g2s(sm.aring[st] + C::kAPlane / 2, src + C::kAPlane + (size_t)xrow0 * (kKC * 2), C::kAPlane / 2, &sm.afull[st]);
                } else {
                  mbar_arrive_expect_tx(&sm.afull[st], C::kAChunk);
                  bulk_g2s(sm.aring[st], src, C::kAChunk, &sm.afull[st]);
                }
                if (i == 0 && hh == 0) trace_ev(p, g * 4 + l, 1);
              }
              __syncwarp();
              if (p.trace != nullptr && lane == 0 && i < 16) trace_clk(p, g * 4 + l, 48 + i);
            }
            ++act_wg[hh][buf];
            ++xchgg[hh];
          }
        }
      }
    } else {
      if (lw == 0) prefetch_small(0);
      bool gave_up = false;
      for (int g = 0; g < total_steps; ++g) {
        const int n = step_subnet(g);
        for (int l = 0; l < p.n_big; ++l) {
          const int buf = xchg & 1;
          const uint32_t expected = p.epoch + 1 + act_w[buf];
          const uint8_t* wbase =
              reinterpret_cast<const uint8_t*>(p.big_w) + (((size_t)n * p.n_big + l) * NT + t) * KCH * kWChunkU;
          const uint8_t* abase = act_slot_of(0) + (size_t)buf * NT * kAStrideU;
          const uint32_t* my_flag = aflag_of(0) + buf * NT + (lane < NT ? lane : 0);
          bool prefetched = lw != C::kLoaders - 1;  // the last loader warp pulls weights into L2, see prefetch_layer
          // JIT: the activations of the first hidden layer are written into the stage by this CTA's own SIMT warps;
          // the loader brings the first-layer weights of the chunk's 64 features instead, and nothing is exchanged
          const bool jit_layer = JIT && l == 0;
          const uint8_t* w1base = reinterpret_cast<const uint8_t*>(p.first_jit) + (size_t)n * KCH * kJitChunkBytes;
          uint32_t ready = (gave_up || jit_layer) ? 0xffffffffu : 0u;  // bit c: producer c has published (warp-uniform)
          for (int i = (lw + C::kLoaders - (int)(ring_pos % C::kLoaders)) % C::kLoaders; i < KCH; i += C::kLoaders) {
            const uint32_t pos = ring_pos + i;
            const int st = pos % kStages;
            const uint32_t use = pos / kStages;
            // everything that does not depend on the stage being free is computed before the wait
            const int kc = (2 * t + i) % KCH;
            const int c = kc >> 1;
            const void* wsrc = wbase + (size_t)kc * kWChunkU;
            const void* asrc = abase + (size_t)c * kAStrideU + (size_t)(kc & 1) * C::kAChunk;
            // lane 0 copies the weights, lane 1 the activations: copies of one thread are processed one after the other,
            // copies of different threads side by side (scripts/ubench/ingest2.cu)
            const void* my_src = lane == 0 ? wsrc : (jit_layer ? (const void*)(w1base + (size_t)kc * kJitChunkBytes) : asrc);
            uint8_t* my_dst = sm.ring[st] + (lane == 0 ? 0 : (jit_layer ? C::kW1Off : kWChunkU));
            const uint32_t my_bytes = lane == 0 ? (uint32_t)kWChunkU : (jit_layer ? (uint32_t)kJitChunkBytes : (uint32_t)C::kAChunk);
            uint64_t* my_bar = (jit_layer && lane == 1) ? &sm.w1full[st] : &sm.full[st];
            if (jit_layer) {
              // the first-layer weights have their own, earlier, hand-over: their area of the stage is free as soon as the
              // SIMT warps have used it, a chunk time or more before the tensor core lets go of the rest of the stage
              const uint32_t w1use = (uint32_t)g * (uint32_t)(KCH / kStages) + (uint32_t)(i / kStages);
              if (w1use > 0) mbar_wait_relaxed(&sm.w1empty[st], (w1use - 1) & 1);
              if (lane == 1) {
                mbar_arrive_expect_tx(&sm.w1full[st], kJitChunkBytes);
                bulk_g2s(my_dst, my_src, my_bytes, my_bar);
              }
              __syncwarp();
            }
            if (use > 0) mbar_wait_relaxed(&sm.empty[st], (use - 1) & 1);
            if (p.trace != nullptr && lane == 0 && i < 16) trace_clk(p, g * 4 + l, 32 + i);
            // JIT kernel: the first two chunks of an exchanged layer (i = 0, 1 <-> kc = 2t, 2t+1) are this CTA's own output;
            // its epilogue warps write them into the stage themselves (and arrive on the full barrier), so that the tensor
            // core starts on them while the publish / fence / flag / poll round trip of the exchange is still under way
            const bool own_chunk = JIT && !jit_layer && i < 2;
            const bool skip_a = jit_layer || own_chunk;  // no activation copy by the loader
            // one look at the flags: if the producer is already done, weights and activations go out together
            if (!skip_a && !((ready >> c) & 1u)) {
              bool ok = false;
              if (lane < NT && !((ready >> lane) & 1u)) ok = (int32_t)(ld_relaxed(my_flag) - expected) >= 0;
              ready |= __ballot_sync(0xffffffffu, ok);
            }
            const bool a_now = !skip_a && ((ready >> c) & 1u);
            if (p.trace != nullptr && lane == 0 && i < 16) trace_clk(p, g * 4 + l, 64 + i);
            // No ordering is needed between lane 0's expect_tx and lane 1's copy: the phase cannot complete before the
            // (single) pending arrival, which is the expect_tx itself, whatever the transient sign of the tx-count.
            if (JIT) {
              // the full barrier of a stage takes two arrivals per phase: the weights' expect_tx, and the SIMT warps' "the
              // activations are written" in a JIT layer (a second plain arrival of the loader in the other layers)
              if (lane == 0) mbar_arrive_expect_tx(&sm.full[st], kWChunkU + (skip_a ? 0 : C::kAChunk));
              if (lane == 2 && !skip_a) mbar_arrive(&sm.full[st]);
            } else {
              if (lane == 0) mbar_arrive_expect_tx(&sm.full[st], (p.debug & 1 ? 0 : C::kAChunk) + (p.debug & 2 ? 0 : kWChunkU));
            }
            if (p.cluster > 1 && lane == 0) {
              // this CTA's share of the weight chunk, delivered to every CTA of the cluster (all of them wait for the same
              // chunk in the same ring stage; the `empty` barrier above has told us that the stage is free in all of them)
              const uint32_t share = (uint32_t)kWChunkU / (uint32_t)p.cluster;
              const uint32_t off = (blockIdx.x % (uint32_t)p.cluster) * share;
              bulk_g2s_multicast(my_dst + off, (const uint8_t*)my_src + off, share, my_bar, (uint16_t)((1u << p.cluster) - 1u));
            } else if (lane < (a_now ? 2 : 1) && !((p.debug >> (1 - lane)) & 1)) {
              bulk_g2s(my_dst, my_src, my_bytes, my_bar);
            }
            __syncwarp();
            if (!prefetched) {  // after this warp's first copies of the layer are on their way
              prefetched = true;
              const int q = g * p.n_big + l;
              const int dist = (p.debug & 32) ? 0 : (p.debug & 64) ? 1 : (p.debug & 128) ? 3 : (p.debug & 256) ? 4 : 2;
              if (q == 0) for (int d = 1; d < dist; ++d) prefetch_layer(d);
              if (dist > 0) prefetch_layer(q + dist);
            }
            if (p.trace != nullptr && lane == 0 && i < 16) trace_clk(p, g * 4 + l, 80 + i);
            if (lane == 0 && i == 0) trace_ev(p, g * 4 + l, 0);
            if (!a_now && !skip_a) {
              uint32_t spins = 0;
              long long t0 = 0;
              while (!((ready >> c) & 1u)) {
                bool ok = false;
                if (lane < NT && !((ready >> lane) & 1u)) ok = (int32_t)(ld_relaxed(my_flag) - expected) >= 0;
                ready |= __ballot_sync(0xffffffffu, ok);
                if ((ready >> c) & 1u) break;
                ++spins;
                if (spins == 64) t0 = clock64();
                if (spins > 64) {
                  __nanosleep(20);
                  if ((spins & 255u) == 0) {
                    int bail = 0;
                    if (lane == 0) {
                      if (ld_relaxed(p.status + 1) == launch_id) bail = 1;
                      else if (clock64() - t0 > 2500000000LL) {
                        report_timeout(p.status, p.status_host, launch_id);
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
              if (lane == 0 && i == 0) trace_ev(p, g * 4 + l, 3);
              // lane 1, not lane 0: lane 0's weight copy may still be in flight, and a thread's copies are processed in order
              if (lane == 1 && !(p.debug & 1)) bulk_g2s(my_dst, my_src, my_bytes, &sm.full[st]);
              __syncwarp();
            }
            if (lane == 0) {
              if (i == 0) trace_ev(p, g * 4 + l, 1);
              if (i == KCH - 1) trace_ev(p, g * 4 + l, 2);
              if (p.trace != nullptr && i < 16) trace_clk(p, g * 4 + l, 48 + i);
            }
          }
          if (lw == 0 && l == 0) prefetch_small(g + 1);
          ring_pos += KCH;
          if (!jit_layer) {  // a JIT layer is not an exchange
            ++act_w[buf];
            ++xchg;
          }
        }
        if (lw == 0 && p.n_big == 0) prefetch_small(g + 1);
      }
    }
  } else if (warp >= C::kMmaWarp && warp < C::kMmaWarp + C::kMmaWarps) {
    // ===== MMA issuers (see Cfg::kMmaWarps): the whole warp runs the loop (warp-uniform control flow), one elected lane drives the tensor
    // core.  With `if (lane == 0)` around the loop ptxas wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY
    // loop ("once per active thread"); behind elect.sync the UTCHMMAs are emitted back to back: 840 -> 560 cycles per
    // k-chunk for the issuing thread (scripts/ubench/umma_loop.cu), which is what paces the hidden layers. =====
    {
      constexpr uint32_t idesc = make_idesc(kFTU, XR, F16);       // N = rows of the row group
      constexpr uint32_t idesc2 = make_idesc(kFTU, 2 * XR, F16);  // N = 2 x rows: activation head and tail stacked
      uint32_t ring_pos = 0;
      uint32_t layers[2] = {0, 0};  // hidden layers issued so far, per group
      const bool x3 = p.precision != IKF_PRECISION_BF16X1;
      const uint64_t d_wh0 = make_desc(smem_u32(sm.ring[0])), d_wl0 = make_desc(smem_u32(sm.ring[0]) + kWPlaneU);
      const uint64_t d_a0 = make_desc(smem_u32(sm.ring[0]) + kWChunkU);
      const int mw = warp - C::kMmaWarp;  // this warp's chunks: i = mw (mod kMmaWarps), its accumulator tile: mw
      const uint32_t tmem_u0 = __shfl_sync(0xffffffffu, tmem, 0);
      constexpr uint32_t kCorrOff = F16 ? XR : 0;  // fp16x3: the scaled correction terms have their own accumulator
      for (int g = 0; g < total_steps; ++g) {
        for (int l = 0; l < p.n_big; ++l)
        for (int hh = 0; hh < NG; ++hh) {  // ping-pong: the layer jobs alternate between the two groups
          if (layers[hh] > 0) mbar_wait(&sm.dempty[hh], (layers[hh] - 1) & 1);  // the epilogue has drained the accumulators
          tc_fence_after();
          if constexpr (C::kSplit) {
            // split rings: chunk number ring_pos + i sits in weight stage (ring_pos + i) % kWStages and activation stage
            // (ring_pos + i) % kAStages
            const uint64_t d_w0 = make_desc(smem_u32(&sm.wring[0][0])), d_as0 = make_desc(smem_u32(&sm.aring[0][0]));
            for (int i = mw; i < KCHL; i += C::kMmaWarps) {
              const uint32_t pos = ring_pos + (uint32_t)i;
              const int sw = pos % C::kWStages, sa = pos % C::kAStages;
              mbar_wait(&sm.wfull[sw], (pos / C::kWStages) & 1);
              mbar_wait(&sm.afull[sa], (pos / C::kAStages) & 1);
              tc_fence_after();
              if (p.trace != nullptr && lane == 0 && i < 16 && hh == 0) trace_clk(p, g * 4 + l, 16 + i);
              const uint64_t dwh = d_w0 + (uint64_t)(sw * (kWChunkU >> 4)), da = d_as0 + (uint64_t)(sa * (C::kAChunk >> 4));
              if (elect_one()) {
                // accumulator tile of this chunk (k-split: first / second half of the CTA's chunks)
                // k-split: tiles 0 / 1 = first / second half of the CTA's split chunks, handed over one after the other; the private
                // chunks (N = the CTA's own 32 rows) go to a third tile behind them
                const bool priv = KS && KP > 0 && i >= KSH;
                // split chunks of the first tile (IKFLOW_B200_DEBUG 8192 / 16384: one less / one more, for A/B runs; valid results)
                const int KH1 = (KSH + 1) / 2 - ((p.debug & 8192) ? 1 : 0) + ((p.debug & 16384) ? 1 : 0);
                const int tile = KS ? (priv ? 2 : (i >= KH1 ? 1 : 0)) : i % kAcc;
                const bool accum = KS ? (i != 0 && i != KH1 && i != KSH) : i >= kAcc;
                const uint32_t tmem_u = tmem_u0 + (uint32_t)(tile * C::kAccCols) + (PP ? (uint32_t)(hh * kAcc * C::kAccCols) : 0u);
                if (priv) {
                  constexpr uint32_t idesc_p = make_idesc(kFTU, RT, F16), idesc2_p = make_idesc(kFTU, 2 * RT, F16);
                  if (x3)
                    mma_chunk_x3(tmem_u, tmem_u + (F16 ? RT : 0), idesc2_p, idesc_p, dwh, dwh + (uint64_t)(kWPlaneU >> 4), da, accum);
                  else
                    mma_chunk_x1(tmem_u, idesc_p, dwh, da, accum);
                } else if (x3)
                  mma_chunk_x3(tmem_u, tmem_u + kCorrOff, idesc2, idesc, dwh, dwh + (uint64_t)(kWPlaneU >> 4), da, accum);
                else
                  mma_chunk_x1(tmem_u, idesc, dwh, da, accum);
                // a tile of the split part is complete: its hand-over starts now
                if (KS && i == KH1 - 1) mma_commit(&sm.dpart[0]);
                if (KS && i == KSH - 1) mma_commit(&sm.dpart[1]);
                // both stages are free once these MMAs have read them (clusters: the weight stage is refilled by every CTA)
                if (!KS && p.cluster > 1) mma_commit_multicast(&sm.wempty[sw], (uint16_t)((1u << p.cluster) - 1u)); else mma_commit(&sm.wempty[sw]);
                mma_commit(&sm.aempty[sa]);
              }
              __syncwarp();
            }
            ring_pos += KCHL;
          } else {
            // unified ring (just-in-time kernel; every layer starts at stage 0): stage index = i % kStages is a compile-time
            // constant inside the unrolled group, so the 3 descriptors of every stage stay in uniform registers (428 instead
            // of 459 cycles of issue per chunk, scripts/ubench/umma_loop.cu); warp mw owns the stages s = mw (mod 2)
            static_assert(C::kSplit || kStages % C::kMmaWarps == 0, "stages are dealt to the MMA warps");
            for (int i0 = 0; i0 < KCH; i0 += kStages) {
              const uint32_t par = (ring_pos / kStages) & 1;
#pragma unroll
              for (int s = 0; s < kStages; ++s) {
                if ((s % C::kMmaWarps) != mw) continue;
                mbar_wait(&sm.full[s], par);
                tc_fence_after();
                if (p.trace != nullptr && lane == 0 && i0 + s < 16) trace_clk(p, g * 4 + l, 16 + i0 + s);
                constexpr uint64_t kStageOff = (uint64_t)(C::kStage >> 4);  // stage offset in the 16-byte address field
                if (elect_one()) {
                  static_assert(kStages % kAcc == 0, "the tile of a stage must not depend on the group");
                  const uint32_t tmem_u = tmem_u0 + (uint32_t)((s % kAcc) * C::kAccCols);  // accumulator tile of this chunk
                  if (x3)
                    mma_chunk_x3(tmem_u, tmem_u + kCorrOff, idesc2, idesc, d_wh0 + s * kStageOff, d_wl0 + s * kStageOff, d_a0 + s * kStageOff, (i0 + s) >= kAcc);
                  else
                    mma_chunk_x1(tmem_u, idesc, d_wh0 + s * kStageOff, d_a0 + s * kStageOff, (i0 + s) >= kAcc);
                  // the stage is free once these MMAs have read it (clusters: tell every CTA that refills it)
                  if (p.cluster > 1) mma_commit_multicast(&sm.empty[s], (uint16_t)((1u << p.cluster) - 1u)); else mma_commit(&sm.empty[s]);
                }
                __syncwarp();
              }
              ring_pos += kStages;
            }
          }
          if (elect_one()) mma_commit(&sm.dfull[hh]);  // accumulator complete
          __syncwarp();
          if (lane == 0 && hh == 0) trace_ev(p, g * 4 + l, 8);
          ++layers[hh];
        }
      }
    }
  } else {
    // ===== epilogue / SIMT warps: thread (f, h) owns hidden feature 128 t + f (TMEM lane f) for the rows of group h =====
    const int f = tid & 127;
    const int h = tid >> 7;    // epilogue group: rows h*ER .. h*ER + ER - 1
    const int sub = f >> 6;    // which of the CTA's two 64-wide k-chunks
    const int kf = f & 63;
    // Synchronisation domain: all epilogue threads -- or, ping-pong, the 128 threads of this thread's row group, which has
    // its own flow state, accumulator tile, exchange slot and flags and runs out of phase with the other group
    constexpr int DT = C::kDomThreads;
    const int dt = PP ? f : tid;  // index inside the domain
    const int gh = PP ? h : 0;    // this thread's row group inside the CTA
    auto bar_dom = [&]() {
      if constexpr (PP) bar_group(h); else bar_epi<ET>();
    };
    float (*su)[kPad] = sm.u + (PP ? h * RT : 0);
    float (*sa)[kPad] = sm.a + (PP ? h * RT : 0);
    float (*scnd)[8] = sm.cnd + (PP ? h * RT : 0);
    float* sld = sm.logdet + (PP ? h * RT : 0);
    uint8_t* const act_slot = act_slot_of(gh);
    uint8_t* const part_slot = part_slot_of(gh);
    uint32_t* const aflag = aflag_of(gh);
    int row0 = PP ? 0 : h * ER;   // first row of the thread's current 64 (ping-pong: of the current pass)
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (PP ? (uint32_t)(h * kAcc * C::kAccCols) : 0u);
    uint32_t part_w[2] = {0, 0};
    uint32_t pxchg = 0;
    uint32_t act_w[2] = {0, 0};  // publishes so far into each activation scratch buffer
    uint32_t axchg = 0;          // activation exchanges so far
    uint32_t layers = 0;         // hidden layers drained so far
    const bool x3 = p.precision != IKF_PRECISION_BF16X1;
    const uint32_t vt_a = PP ? smem_u32(sm.vt[0]) : JIT ? smem_u32(sm.vt[h]) : smem_u32(&sm.aring[0][0]) + h * C::kVtBytes;
    const int j8 = lane & 7;     // publish: row inside an 8-row block after the lane transpose
    const uint32_t xin_a = smem_u32(&sa[0][0]);
    const bool tiled_first = !JIT && (KS || !(p.debug & 4096));
    // k-split: shared::cluster addresses of the peer CTA's receive buffer and barrier
    uint32_t peer_recv = 0, peer_rbar = 0;
    if constexpr (KS) {
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_recv) : "r"(smem_u32(sm.recv)), "r"((uint32_t)(kh ^ 1)));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_rbar) : "r"(smem_u32(&sm.rbar)), "r"((uint32_t)(kh ^ 1)));
    }

    // u <- u[:, table]
    auto permute_state = [&](const int* table) {
      constexpr int kPer = (RT * kPad + DT - 1) / DT;
      float tmp[kPer];
#pragma unroll
      for (int c = 0; c < kPer; ++c) {
        const int i = dt + c * DT;
        tmp[c] = i < RT * p.W ? su[i / p.W][table[i % p.W]] : 0.f;
      }
      bar_dom();
#pragma unroll
      for (int c = 0; c < kPer; ++c) {
        const int i = dt + c * DT;
        if (i < RT * p.W) su[i / p.W][i % p.W] = tmp[c];
      }
      bar_dom();
    };

    // (every synchronisation domain writes the whole table -- identical values --, so that the barriers of its own row
    // group's load order the writes before the first use)
    for (int i = dt; i < (kMaxFold + 1) * kPad; i += DT) sm.phys[i / kPad][i % kPad] = p.fold ? p.phys[i / kPad][i % kPad] : (uint8_t)(i % kPad);
    int g = 0;
    for (int rgi = 0; rgi < my_rgs; ++rgi) {
      // (>= n_rowgroups: an empty row group of a cluster / of a ping-pong CTA in lock step)
      const int rg = PP ? 2 * slot + h + rgi * 2 * p.slots : slot + rgi * p.slots;
      for (int i = dt; i < RT * kPad; i += DT) {
        const int r = i / kPad, j = i % kPad;
        const int row = rg * XR + xrow0 + r;
        float uv = 0.f, cv = 0.f;
        if (row < p.batch) {
          if (j < p.W) uv = p.in[(size_t)row * p.in_ld + j];
          if (j < p.cond_cols) cv = p.cond[(size_t)((p.row_base + row) % p.cond_rows) * p.cond_ld + j];
        }
        su[r][j] = uv;
        if (j < 8) scnd[r][j] = cv;
      }
      bar_dom();
      if (p.forward) {
        // FixedLinearTransform forward, x.mm(M) + b (FrEIA; ikflow/model.py:197), and a fresh log-det accumulator
        constexpr int kPer = (RT * kPad + DT - 1) / DT;
        float tmp[kPer];
#pragma unroll
        for (int c = 0; c < kPer; ++c) {
          const int i = dt + c * DT, r = i / kPad, j = i % kPad;
          float o = 0.f;
          if (i < RT * kPad && j < p.W) {
            for (int k = 0; k < p.W; ++k) o = fmaf(su[r][k], p.m_fwd[k * kPad + j], o);
            o += p.flt_b[j];
          }
          tmp[c] = o;
        }
        bar_dom();
#pragma unroll
        for (int c = 0; c < kPer; ++c) {
          const int i = dt + c * DT;
          if (i < RT * kPad) su[i / kPad][i % kPad] = tmp[c];
        }
        for (int r = dt; r < RT; r += DT) sld[r] = p.logdet_m;
        bar_dom();
      }

      for (int bi = 0; bi < n_blocks; ++bi) {
        const int blk = p.forward ? p.block_last + bi : p.block_first - bi;
        // PermuteRandom (forward: x[:, perm] before the block; reverse: x[:, perm_inv] after it) is folded into the indexing
        // of the state: logical column j of this block sits in su[.][ph[j]]
        const uint8_t* ph = sm.phys[p.fold ? bi : 0];
        if (p.forward && !p.fold) permute_state(p.perm_fwd + blk * kPad);
        for (int step = 0; step < 2; ++step, ++g) {
          const int sidx = p.forward ? 1 - step : step;
          const int sb = g & 1;
          const uint32_t sp_a = smem_u32(sm.small[sb]);  // explicit shared-space accesses (see lds128)
          const int in_off = sidx == 0 ? 0 : p.s1;
          const int in_len = sidx == 0 ? p.s1 : p.s2;
          const int tg_off = sidx == 0 ? p.s1 : 0;
          const int tg_len = sidx == 0 ? p.s2 : p.s1;
          const int kin = in_len + p.dim_cond;
          // subnet input [state half | condition | 0] (sm.a doubles as this buffer until the last layer)
          for (int i = dt; i < RT * kPad; i += DT) {
            const int r = i / kPad, k = i % kPad;
            sa[r][k] = k < in_len ? su[r][ph[in_off + k]] : (k < kin ? scnd[r][k - in_len] : 0.f);
          }
          mbar_wait(&sm.small_full[sb], (g >> 1) & 1);
          bar_dom();
          if (tid == 0) trace_ev(p, g * 4, 2);

          float v[ER];  // activations of feature f for the rows of this group
          if constexpr (JIT) {
            // ---- first layer, just in time (see jit_chunk): the 4 epilogue warps and the 4 helper warps share every chunk ----
            if (tid == 0) trace_ev(p, g * 4, 10);
            bar_gen<C::kGenThreads>();  // the helpers may read the subnet's input
            jit_layer_loop(0, warp, g, kin);
          } else if (tiled_first) {
            // ---- first layer, tile by tile, published as it is computed (first_layer_tile) ----
            constexpr int kTilesPerWarp = (RT / 32) * (kFTU / 16) / (PP ? 4 : C::kEpiWarps);  // (ping-pong: the group's four warps)
            const int tile0 = (PP ? (warp & 3) : warp) * kTilesPerWarp;
            const int rb = tile0 / (kFTU / 16), fb0 = tile0 % (kFTU / 16);
            uint8_t* dst0 = act_slot + ((size_t)(axchg & 1) * NT + t) * kAStrideU;
            const int rb_out = KS ? kh : rb;  // 32-row block inside the exchanged tile (k-split: this CTA's half of the rows)
            auto run = [&](auto kb) {
              constexpr int KB = decltype(kb)::value;
              float x[4][KB];
#pragma unroll
              for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                for (int k4 = 0; k4 < KB; k4 += 4) {
                  const float4 xv = lds128(xin_a + ((rb * 32 + (lane >> 2) + 8 * rr) * kPad + k4) * 4);
                  x[rr][k4] = xv.x, x[rr][k4 + 1] = xv.y, x[rr][k4 + 2] = xv.z, x[rr][k4 + 3] = xv.w;
                }
#pragma unroll
              for (int ti = 0; ti < kTilesPerWarp; ++ti)
                first_layer_tile<KB, C::kAPlane, F16>(sp_a + (kSmFirstW - C::kSmShift) * 4, sp_a + (kSmFirstB - C::kSmShift) * 4, x, rb_out, fb0 + ti,
                                                 lane, dst0, C::kAChunk, p.status, p.status_host);
            };
            if (kin <= 12) run(std::integral_constant<int, 12>{});
            else run(std::integral_constant<int, kPad>{});
            if (tid == 0) trace_ev(p, g * 4, 3);
          } else {
          // ---- first layer: fp32 FMA, one thread per feature (kept for A/B runs: IKFLOW_B200_DEBUG=4096) ----
            const float b0 = lds32(sp_a + (kSmFirstB - C::kSmShift + f) * 4);
#pragma unroll
            for (int r = 0; r < ER; ++r) v[r] = b0;
            for (int k4 = 0; k4 < kin; k4 += 4) {  // the input is zero-padded to 16 columns, the weights too
              const uint32_t wa = sp_a + (kSmFirstW - C::kSmShift + k4 * kFTU + f) * 4;
              const float w0 = lds32(wa), w1 = lds32(wa + kFTU * 4), w2 = lds32(wa + 2 * kFTU * 4), w3 = lds32(wa + 3 * kFTU * 4);
              // every epilogue warp is alone on its scheduler: issue the loads of 8 rows back to back, then the 32 FMAs
              // that consume them, so that the shared-memory latency is paid once per batch
#pragma unroll
              for (int r0 = 0; r0 < ER; r0 += 8) {
                float4 x[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) x[r] = lds128(xin_a + ((row0 + r0 + r) * kPad + k4) * 4);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                  v[r0 + r] = fmaf(x[r].x, w0, v[r0 + r]);
                  v[r0 + r] = fmaf(x[r].y, w1, v[r0 + r]);
                  v[r0 + r] = fmaf(x[r].z, w2, v[r0 + r]);
                  v[r0 + r] = fmaf(x[r].w, w3, v[r0 + r]);
                }
              }
            }
#pragma unroll
            for (int r = 0; r < ER; ++r) v[r] = leaky(v[r]);
          }

          const int pb = pxchg & 1;
          const uint32_t pexp = p.epoch + 1 + part_w[pb];  // sequence number of this subnet's partial-sum exchange
          const int n_units = tg_len;                       // 2 * tg_len outputs, two per 16-byte unit
          for (int l = JIT ? 1 : 0; l <= p.n_big; ++l) {
            if (l > 0 && tid == 0) trace_ev(p, g * 4 + l - 1, 6);
            if (!KS && l > 0) {
              // ---- hidden layer l-1: its accumulator is complete ----
              mbar_wait(&sm.dfull[gh], layers & 1);
              tc_fence_after();
              if (tid == 0) trace_ev(p, g * 4 + l - 1, 7);
            }
            if (l == p.n_big) {
              if constexpr (PP) {
                // the last layer's transposed tile is shared by the two groups (their last layers are half a cycle apart)
                if (dt == 0) {
                  const long long tl = clock64();
                  while (atomicCAS(&sm.vt_lock, 0, 1) != 0) {
                    __nanosleep(64);
                    if (clock64() - tl > 4000000000LL) __trap();
                  }
                  __threadfence_block();
                }
                bar_dom();
              }
              if (tid == 0) trace_ev(p, g * 4 + 3, 11);
            }
            float pub_max = 0.f;  // fp16x3: largest magnitude published (checked once, after the stores)
            const int buf = axchg & 1;
            // ping-pong: the four warps of a group take its 128 rows in two passes of 64
#pragma unroll 1
            for (int pass = 0; pass < C::kPasses; ++pass) {
              if constexpr (PP) row0 = pass * ER;
              if (l > 0) {
                // ---- drain the accumulator (the rows of this pass) ----
                if constexpr (KS) {
                  // ---- k-split: this CTA's two tiles hold the partial sums (first / second half of its half of k) of ALL 64
                  //      rows.  The other CTA's 32 rows go to its receive buffers through distributed shared memory ([row]
                  //      [feature]: a warp stores 128 contiguous bytes), the own 32 rows stay in registers.  The first tile
                  //      is handed over while the tensor core still works on the second: only the second hand-over is exposed ----
#pragma unroll
                  for (int m = 0; m < 2; ++m) {
                    mbar_wait(&sm.dpart[m], layers & 1);  // tile m of the split part is complete
                    tc_fence_after();
                    if (tid == 0 && m == 1) trace_ev(p, g * 4 + l - 1, 7);
                    // the peer's rows first (it waits for them); the loads of the own rows are in flight while those go out
                    float tmp[32], tmp2[32];
                    const int cp = RT * (kh ^ 1), co = RT * kh;
                    if (x3) {
                      tmem_ld32x2(taddr + m * C::kAccCols + cp, taddr + m * C::kAccCols + XR + cp, tmp, tmp2);
#pragma unroll
                      for (int r = 0; r < 32; ++r) tmp[r] = F16 ? fmaf(tmp2[r], 1.f / kTailScaleF16, tmp[r]) : tmp[r] + tmp2[r];
                    } else {
                      tmem_ld32(taddr + m * C::kAccCols + cp, tmp);
                    }
                    if (tid == 0 && m == 1) trace_ev(p, g * 4 + l - 1, 11);
                    uint32_t o1[32], o2[32];
                    if (x3) tmem_ld32x2_issue(taddr + m * C::kAccCols + co, taddr + m * C::kAccCols + XR + co, o1, o2);
                    // st.async: every 16-byte store signals its bytes on the peer's barrier when it lands -- no release (which
                    // would wait for the acknowledgements of the stores before the arrival even leaves)
#pragma unroll
                    for (int j = 0; j < RT / 4; ++j)
                      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(
                                       peer_recv + (uint32_t)((m * kFTU + f) * (RT * 4) + ((j ^ (f & 7)) << 4))),
                                   "r"(__float_as_uint(tmp[4 * j])), "r"(__float_as_uint(tmp[4 * j + 1])), "r"(__float_as_uint(tmp[4 * j + 2])),
                                   "r"(__float_as_uint(tmp[4 * j + 3])), "r"(peer_rbar)
                                   : "memory");
                    if (tid == 0 && m == 1) trace_ev(p, g * 4 + l - 1, 12);
                    if (x3) {
                      tmem_ld32x2_finish(o1, o2, tmp, tmp2);
#pragma unroll
                      for (int r = 0; r < 32; ++r) tmp[r] = F16 ? fmaf(tmp2[r], 1.f / kTailScaleF16, tmp[r]) : tmp[r] + tmp2[r];
                    } else {
                      tmem_ld32(taddr + m * C::kAccCols + co, tmp);
                    }
                    if (tid == 0 && m == 1) trace_ev(p, g * 4 + l - 1, 13);
#pragma unroll
                    for (int r = 0; r < 32; ++r) v[r] = m == 0 ? tmp[r] : v[r] + tmp[r];
                  }
                  mbar_wait(&sm.dfull[0], layers & 1);
                  if (KP > 0) {
                    // the private chunks' tile: this CTA's own rows only (nothing to hand over)
                    tc_fence_after();
                    float tmp[32], tmp2[32];
                    if (x3) {
                      tmem_ld32x2(taddr + 2 * C::kAccCols, taddr + 2 * C::kAccCols + RT, tmp, tmp2);
#pragma unroll
                      for (int r = 0; r < 32; ++r) v[r] += F16 ? fmaf(tmp2[r], 1.f / kTailScaleF16, tmp[r]) : tmp[r] + tmp2[r];
                    } else {
                      tmem_ld32(taddr + 2 * C::kAccCols, tmp);
#pragma unroll
                      for (int r = 0; r < 32; ++r) v[r] += tmp[r];
                    }
                  }
                } else {
#pragma unroll
                for (int c0 = 0; c0 < ER; c0 += 32) {
                  // the accumulator tiles (k-chunks i = m mod kAcc), added in a fixed order
#pragma unroll
                  for (int m = 0; m < kAcc; ++m) {
                    if (m >= KCH) break;  // a layer of fewer chunks than tiles (hidden = 128) leaves the rest untouched
                    float tmp[32], tmp2[32];
                    // main: bf16x3 W_head*A_head + W_tail*A_head, fp16x3 W_head*A_head; second: bf16x3 W_head*A_tail, fp16x3
                    // 2^11 (W_head*A_tail + W_tail*A_head)
                    if (x3) {
                      tmem_ld32x2(taddr + m * C::kAccCols + row0 + c0, taddr + m * C::kAccCols + RT + row0 + c0, tmp, tmp2);
#pragma unroll
                      for (int r = 0; r < 32; ++r) tmp[r] = F16 ? fmaf(tmp2[r], 1.f / kTailScaleF16, tmp[r]) : tmp[r] + tmp2[r];
                    } else {
                      tmem_ld32(taddr + m * C::kAccCols + row0 + c0, tmp);
                    }
#pragma unroll
                    for (int r = 0; r < 32; ++r) v[c0 + r] = m == 0 ? tmp[r] : v[c0 + r] + tmp[r];
                  }
                }
                }
                if (pass == C::kPasses - 1) {
                  tc_fence_before();
                  __syncwarp();
                  if (lane == 0) mbar_arrive(&sm.dempty[gh]);
                }
                if constexpr (KS) {
                  // the peer's partial sums of my rows (one barrier phase per hidden layer, in lock step: neither CTA can run a
                  // layer ahead, it needs the other's publication first)
                  {
                    uint32_t ok = 0;
                    const long long tw = clock64();
                    while (!ok) {
                      asm volatile("{\n\t.reg .pred pp;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 pp, [%1], %2;\n\tselp.u32 %0, 1, 0, pp;\n\t}"
                                   : "=r"(ok) : "r"(smem_u32(&sm.rbar)), "r"(layers & 1u) : "memory");
                      if (!ok && clock64() - tw > 4000000000LL) __trap();
                    }
                  }
                  if (tid == 0) trace_ev(p, g * 4 + l - 1, 14);
                  if (tid == 0) mbar_arrive_expect_tx(&sm.rbar, C::kRecvBytes);  // the next layer's phase
                  const uint32_t rv = smem_u32(sm.recv) + (uint32_t)(f * (RT * 4));
#pragma unroll
                  for (int m = 0; m < 2; ++m) {
#pragma unroll
                    for (int j = 0; j < RT / 4; ++j) {
                      const float4 x = lds128(rv + (uint32_t)(m * kFTU * RT * 4 + ((j ^ (f & 7)) << 4)));
                      v[4 * j] += x.x, v[4 * j + 1] += x.y, v[4 * j + 2] += x.z, v[4 * j + 3] += x.w;
                    }
                  }
                }
                const float bb = lds32(sp_a + (kSmBigB - C::kSmShift + (l - 1) * kFTU + f) * 4);
#pragma unroll
                for (int r = 0; r < ER; ++r) v[r] = leaky(v[r] + bb);
                if (tid == 0) trace_ev(p, g * 4 + l - 1, 9);
              }
              if (l < p.n_big) {
                // ---- publish straight from the registers: bf16 head/tail split, 8x8 transpose across the 8 lanes that hold
                //      8 consecutive features (so that every lane owns one 16-byte chunk = 8 features of one row), two
                //      16-byte global stores per lane and 8-row block into the producer's slot of the scratch ring
                //      [k-chunk 2][head|tail][RT rows][64 k, swizzled]; then ONE release flag per CTA ----
                uint8_t* dst = act_slot + ((size_t)buf * NT + t) * kAStrideU + (size_t)sub * C::kAChunk;
                const int kf8 = kf & ~7;
                if (!(l == 0 && tiled_first))  // (the tiled first layer has stored its output already)
#pragma unroll
                for (int r0 = 0; r0 < ER; r0 += 8) {
                  uint32_t wd[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    if (F16) pub_max = fmaxf(pub_max, fabsf(v[r0 + i]));
                    wd[i] = split_one<F16>(v[r0 + i]);  // head in the low half, tail in the high half
                  }
                  // before: lane j8 (feature kf8 + j8) holds rows r0..r0+7; after: lane j8 holds row r0 + j8, features kf8..kf8+7
#pragma unroll
                  for (int st = 4; st >= 1; st >>= 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      if ((i & st) == 0) {
                        const bool up = (j8 & st) != 0;
                        const uint32_t send = up ? wd[i] : wd[i + st];
                        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, st);
                        if (up) wd[i] = recv; else wd[i + st] = recv;
                      }
                    }
                  }
                  const uint32_t off = tile_off_bytes(xrow0 + row0 + r0 + j8, kf8);
                  stg128(dst + off, __byte_perm(wd[0], wd[1], 0x5410), __byte_perm(wd[2], wd[3], 0x5410),
                         __byte_perm(wd[4], wd[5], 0x5410), __byte_perm(wd[6], wd[7], 0x5410));
                  stg128(dst + C::kAPlane + off, __byte_perm(wd[0], wd[1], 0x7632), __byte_perm(wd[2], wd[3], 0x7632),
                         __byte_perm(wd[4], wd[5], 0x7632), __byte_perm(wd[6], wd[7], 0x7632));
                  if constexpr (JIT) {  // ... and into this CTA's own ring stages 0 / 1 (the next layer starts at stage 0)
                    const uint32_t own_a = smem_u32(sm.ring[sub]) + kWChunkU + off;
                    sts128u(own_a, __byte_perm(wd[0], wd[1], 0x5410), __byte_perm(wd[2], wd[3], 0x5410),
                            __byte_perm(wd[4], wd[5], 0x5410), __byte_perm(wd[6], wd[7], 0x5410));
                    sts128u(own_a + C::kAPlane, __byte_perm(wd[0], wd[1], 0x7632), __byte_perm(wd[2], wd[3], 0x7632),
                            __byte_perm(wd[4], wd[5], 0x7632), __byte_perm(wd[6], wd[7], 0x7632));
                  }
                }
              } else {
                // ---- last layer: fp32, through a transposed copy of the activations; every group works on its own rows, 32
                //      at a time ----
                {
                  // thread = (k quarter kq, row pair rp, output group og): 2 rows x 8 outputs over 32 of the 128 features,
                  // then a 4-lane shuffle reduction over the k quarters
                  constexpr int OUTS = 8;
                  const int kq = f & 3, rp = (f >> 2) & 15, og = f >> 6;
                  const uint32_t vrow = vt_a + (2 * rp) * kFTU * 4;
                  const uint32_t wrow = sp_a + (kSmLastW - C::kSmShift + og * OUTS * kFTU) * 4;
                  const int osw = (og * OUTS) >> 2;  // (o >> 2) = osw + (oo >> 2)
#pragma unroll
                  for (int ps = 0; ps < ER / 32; ++ps) {
                    // [32][128] fp32; float4 slot j4 of row r sits at (j4 & ~7) | ((j4 ^ (j4 >> 3) ^ ((r >> 1) << 2)) & 7): the 8
                    // lanes of a quarter warp of the reader (4 k-quarters x 2 row pairs) then hit 8 different bank groups
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                      sts32(vt_a + (r * kFTU + (((((f >> 2) & ~7) | (((f >> 2) ^ (f >> 5) ^ ((r >> 1) << 2)) & 7)) << 2) | (f & 3))) * 4,
                            v[ps * 32 + r]);
                    bar_group(h);
                    float po[2][OUTS];
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                      for (int oo = 0; oo < OUTS; ++oo) po[q][oo] = 0.f;
#pragma unroll 2
                    for (int jj = 0; jj < 8; ++jj) {
                      const int slot_ = (kq << 3) | ((jj ^ kq ^ (rp << 2)) & 7);  // both rows of the pair share (row >> 1)
                      // all loads of the step first, then the FMAs (see the first layer)
                      float4 w[OUTS];
                      const float4 x0 = lds128(vrow + (slot_ << 4));
                      const float4 x1 = lds128(vrow + kFTU * 4 + (slot_ << 4));
#pragma unroll
                      for (int oo = 0; oo < OUTS; ++oo)
                        w[oo] = lds128(wrow + (oo * kFTU + (((kq << 3) | ((jj ^ kq ^ (osw + (oo >> 2))) & 7)) << 2)) * 4);
#pragma unroll
                      for (int oo = 0; oo < OUTS; ++oo) {
                        po[0][oo] = fmaf(x0.x, w[oo].x, po[0][oo]);
                        po[1][oo] = fmaf(x1.x, w[oo].x, po[1][oo]);
                        po[0][oo] = fmaf(x0.y, w[oo].y, po[0][oo]);
                        po[1][oo] = fmaf(x1.y, w[oo].y, po[1][oo]);
                        po[0][oo] = fmaf(x0.z, w[oo].z, po[0][oo]);
                        po[1][oo] = fmaf(x1.z, w[oo].z, po[1][oo]);
                        po[0][oo] = fmaf(x0.w, w[oo].w, po[0][oo]);
                        po[1][oo] = fmaf(x1.w, w[oo].w, po[1][oo]);
                      }
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                      for (int oo = 0; oo < OUTS; ++oo) {
                        float x = po[q][oo];
                        x += __shfl_xor_sync(0xffffffffu, x, 1);
                        x += __shfl_xor_sync(0xffffffffu, x, 2);
                        po[q][oo] = x;
                      }
                    // all four lanes of a k-quarter group hold the sums: lane kq publishes outputs og*8 + 2kq, +1 of both rows,
                    // straight from its registers, in the LL format
                    if (og * 4 + kq < n_units) {
#pragma unroll
                      for (int q = 0; q < 2; ++q) {
                        const float v0 = kq == 0 ? po[q][0] : kq == 1 ? po[q][2] : kq == 2 ? po[q][4] : po[q][6];
                        const float v1 = kq == 0 ? po[q][1] : kq == 1 ? po[q][3] : kq == 2 ? po[q][5] : po[q][7];
                        st_ll(part_slot + (((size_t)pb * NT + t) * XR + xrow0 + row0 + ps * 32 + 2 * rp + q) * kPartRowBytes + (og * 4 + kq) * 16,
                              v0, v1, pexp);
                      }
                    }
                    if (ps + 1 < ER / 32 || pass + 1 < C::kPasses) bar_group(h);  // the tile is free for the next 32 rows
                  }
                }
              }
            }
            if (l > 0) ++layers;
            if (l < p.n_big) {
              if (F16 && pub_max > 65504.f) report_range(p.status, p.status_host);  // the fp16 head of such a value is inf
              if constexpr (JIT) fence_proxy_async_smem();  // the tensor core reads the own chunks through the async proxy
              bar_dom();
              if constexpr (JIT) {
                if (tid == 0) {
                  mbar_arrive(&sm.full[0]);
                  mbar_arrive(&sm.full[1]);
                }
              }
              // st.release is cumulative over the stores the barrier ordered before it.  It is NOT optional: a relaxed
              // flag store lets consumers read stale chunks (scripts/stress_flow.py); its MEMBAR.GPU costs ~1 us.
              if (dt == 0) {  // (ping-pong: thread 0 of the group)
                if (tid == 0) trace_ev(p, g * 4 + l, 4);
                st_release(aflag + (buf * NT + t) * FW + (KS ? kh : 0), p.epoch + 1 + act_w[buf]);
                if (tid == 0) trace_ev(p, g * 4 + l, 5);
              }
              ++act_w[buf];
              ++axchg;
              if (tid == 0) trace_ev(p, g * 4 + l, 10);
            } else if constexpr (PP) {
              bar_dom();
              if (dt == 0) {
                __threadfence_block();
                atomicExch(&sm.vt_lock, 0);
              }
            }
          }
          if (tid == 0) {
            trace_ev(p, g * 4 + 3, 14);
            trace_ev(p, g * 4 + 3, 15);
          }
          // ---- sum the team's partial sums in a fixed order (bitwise identical replicas), polling the data itself ----
          for (int i = dt; i < RT * 4; i += DT) {
            const int r = i >> 2, o4 = i & 3;
            float4 acc = lds128(sp_a + (kSmLastB - C::kSmShift + 4 * o4) * 4);
            const uint8_t* src = part_slot + ((size_t)pb * NT * XR + xrow0 + r) * kPartRowBytes + (2 * o4) * 16;
            const bool need0 = 2 * o4 < n_units, need1 = 2 * o4 + 1 < n_units;
            for (int c0 = 0; c0 < NT; c0 += 8) {
              uint4 x0[8], x1[8];
              uint32_t spins = 0;
              long long t0 = 0;
              while (true) {
                bool ok = true;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  x0[c] = make_uint4(0u, pexp, 0u, pexp);
                  x1[c] = make_uint4(0u, pexp, 0u, pexp);
                  if (c0 + c < NT) {
                    if (need0) x0[c] = ld_ll(src + (size_t)(c0 + c) * XR * kPartRowBytes);
                    if (need1) x1[c] = ld_ll(src + (size_t)(c0 + c) * XR * kPartRowBytes + 16);
                  }
                }
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  ok = ok && x0[c].y == pexp && x0[c].w == pexp && x1[c].y == pexp && x1[c].w == pexp;
                if (ok) break;
                // bounded: a lost producer must surface as IKF_STATUS_SYNC_TIMEOUT, not as a hung GPU
                ++spins;
                if (spins == 64) t0 = clock64();
                if (spins > 64) {
                  __nanosleep(20);
                  if ((spins & 255u) == 0) {
                    if (ld_relaxed(p.status + 1) == launch_id) break;
                    if (clock64() - t0 > 2500000000LL) {
                      report_timeout(p.status, p.status_host, launch_id);
                      break;
                    }
                  }
                }
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                acc.x += __uint_as_float(x0[c].x);
                acc.y += __uint_as_float(x0[c].z);
                acc.z += __uint_as_float(x1[c].x);
                acc.w += __uint_as_float(x1[c].z);
              }
            }
            sts128(smem_u32(&sa[r][4 * o4]), acc);
          }
          if (tid == 0) trace_ev(p, g * 4 + 3, 12);
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.small_empty[sb]);
          ++part_w[pb];
          ++pxchg;
          bar_dom();
          for (int i = dt; i < RT * tg_len; i += DT) {
            const int r = i / tg_len, j = i % tg_len;
            const float sc = p.clamp_scale * atanf(sa[r][j]);
            const float tr = sa[r][tg_len + j];
            float& uu = su[r][ph[tg_off + j]];
            uu = p.forward ? fmaf(uu, expf(sc), tr) : (uu - tr) * expf(-sc);
          }
          if (p.forward)  // log-det of the block: sum of the (clamped) scales, one thread per row, fixed order
            for (int r = dt; r < RT; r += DT) {
              float acc = sld[r];
              for (int j = 0; j < tg_len; ++j) acc += p.clamp_scale * atanf(sa[r][j]);
              sld[r] = acc;
            }
          bar_dom();
          if (tid == 0) trace_ev(p, g * 4 + 3, 13);
        }
        if (!p.forward && !p.fold) permute_state(p.perm_inv + blk * kPad);
      }

      if (t == 0) {
        const uint8_t* phf = sm.phys[p.fold ? n_blocks : 0];  // where the logical columns sit at the end
        for (int i = dt; i < RT * p.out_cols; i += DT) {
          const int r = i / p.out_cols, j = i % p.out_cols;
          const int row = rg * XR + xrow0 + r;
          if (row >= p.batch) continue;
          float o;
          if (p.finalize) {
            o = 0.f;
            for (int k = 0; k < p.W; ++k) o = fmaf(su[r][phf[k]] - p.flt_b[k], p.m_inv[k * kPad + j], o);
            // NaN / inf are reported BEFORE the clamp, and NaN survives it as in torch.clamp (robot.clamp_to_joint_limits)
            if (!isfinite(o)) report_nonfinite(p.status, p.status_host);
            if (p.clamp_out && j < p.ndof && o == o) o = fminf(fmaxf(o, p.lo[j]), p.hi[j]);
          } else {
            o = su[r][phf[j]];
            if (!isfinite(o)) report_nonfinite(p.status, p.status_host);
          }
          p.out[(size_t)row * p.out_ld + j] = o;
          // fused gather: the same value into the gathered buffer of every rank of the node (NVLink stores)
          for (int r = 0; r < p.n_peers; ++r) p.peer_out[r][(size_t)(p.peer_row0 + row) * p.peer_ld + j] = o;
        }
        if (p.forward && p.logdet_out != nullptr)
          for (int r = dt; r < RT; r += DT)
            if (rg * XR + xrow0 + r < p.batch) p.logdet_out[rg * XR + xrow0 + r] = sld[r];
      }
      bar_dom();
    }
  }

  if (p.n_peers > 0 && t == 0 && warp < C::kEpiWarps) {
    // fused gather: this CTA's rows are on their way to every rank.  The last writer CTA of the launch publishes the
    // shard: fence (system scope, cumulative over the stores the barrier / the counter ordered before it), then this rank's
    // sequence number into the flag array of every rank.
    bar_epi<ET>();
    if (tid == 0) {
      __threadfence_system();
      const uint32_t old = atomicAdd(p.peer_counter, 1u);
      if (old + 1u == p.peer_count_target) {
        __threadfence_system();  // ONE system-scope fence orders every shard store before the flags (a release per flag
                                 // would wait for the NVLink round trip of the previous one: 8 x ~2 us at 8 ranks)
        for (int r = 0; r < p.n_peers; ++r) st_relaxed_sys(p.peer_flag[r] + p.peer_rank, p.peer_seq);
        // ... and, as the last CTA alive, waits for the shards of the other ranks: when this kernel ends, this rank's
        // gathered tensor is complete -- no second launch, no collective.  (Every rank publishes before it waits, so the
        // ranks cannot wait for each other in a circle; the wait is bounded like all others.)
        const uint32_t* mine = p.peer_flag[p.peer_rank];
        const long long t0 = clock64();
        for (int r = 0; r < p.n_peers; ++r) {
          uint32_t spins = 0;
          while ((int32_t)(ld_acquire_sys(mine + r) - p.peer_seq) < 0) {
            if (++spins > 32) __nanosleep(40);
            if ((spins & 1023u) == 0 && clock64() - t0 > 6000000000LL) {  // ~3 s: a peer never published
              report_timeout(p.status, p.status_host, launch_id);
              break;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == C::kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemColsK) : "memory");
  }
}

#undef IKF_CS

}  // namespace umma
}  // namespace ikf
