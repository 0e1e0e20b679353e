// What does the MMA-issuing thread's per-chunk protocol cost?  One CTA, kernel-like loop over 64-wide k-chunks
// (8 stacked MMAs per chunk, M = 128, N = 2 RT / RT), ring stages of 40 KB.
//   mode 0: MMAs only                      mode 1: + tcgen05.commit per chunk (nobody waits on it)
//   mode 2: + full/empty handshake with 4 "loader" warps that only arrive (no data movement)
//   mode 3: mode 2 with real 40 KB bulk copies per stage (32 KB + 8 KB, L2 resident)
//   mode 4: mode 3, but commit/handshake every 2 chunks (80 KB stages, 2 stages, 4 x 20 KB copies)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t par) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t par) { while (!mbar_try(bar, par)) {} }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0; d |= (uint64_t)((saddr >> 4) & 0x3fff); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__device__ __forceinline__ void mma_chunk_x3(uint32_t tmem_d, uint32_t idesc_2n, uint32_t idesc_n, uint64_t wh, uint64_t wl, uint64_t a, uint32_t acc_first) {
  asm volatile(
      "{\n\t.reg .pred p0, p1;\n\t.reg .b64 wh, wl, a;\n\tsetp.ne.b32 p0, %6, 0;\n\tsetp.eq.b32 p1, %6, %6;\n\t"
      "mov.b64 wh, %3;\n\tmov.b64 wl, %4;\n\tmov.b64 a, %5;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, a, %1, p0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], wl, a, %2, p1;\n\t"
      "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, a, %1, p1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], wl, a, %2, p1;\n\t"
      "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, a, %1, p1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], wl, a, %2, p1;\n\t"
      "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], wh, a, %1, p1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], wl, a, %2, p1;\n\t"
      "}\n" ::"r"(tmem_d), "r"(idesc_2n), "r"(idesc_n), "l"(wh), "l"(wl), "l"(a), "r"(acc_first) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
constexpr int kStage = 40960;
extern __shared__ __align__(1024) uint8_t smem_raw[];
template <bool kConstTmem, bool kElect, bool kStatic>
__global__ void k(int mode, int chunks, int RT, const uint8_t* src, unsigned long long* out) {
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(sm + 5 * kStage);
  uint64_t* empty = full + 4;
  uint64_t* done = empty + 4;
  uint32_t* slot = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int per = 1;          // chunks per stage
  constexpr int nst = 4;          // stages
  const int stage_bytes = per * kStage;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_real = *slot;
  const uint32_t tmem = kConstTmem ? 0u : tmem_real;  // the only allocation of the only CTA on the SM starts at column 0
  if (kConstTmem && tmem_real != 0) __trap();
  const int nsteps = chunks / per;
  if (warp == 0 && (kElect || lane == 0)) {
    const uint32_t idN = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t id2N = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * RT) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t wh0 = make_desc(smem_u32(sm)), wl0 = make_desc(smem_u32(sm) + 16384), a0 = make_desc(smem_u32(sm) + 32768);
    long long t0 = clock64();
    if (!kStatic) for (int c = 0; c < nsteps; ++c) {
      const int s = c % nst;
      if (mode >= 2) { mbar_wait(&full[s], (c / nst) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      for (int q = 0; q < per; ++q) {
        const uint64_t soff = (uint64_t)((s * stage_bytes + q * kStage) >> 4);
        if (kElect) { if (elect_one()) mma_chunk_x3(tmem, id2N, idN, wh0 + soff, wl0 + soff, a0 + soff, 1); }
        else mma_chunk_x3(tmem, id2N, idN, wh0 + soff, wl0 + soff, a0 + soff, 1);
      }
      if (mode >= 1) { if (kElect) { if (elect_one()) commit(&empty[s]); } else commit(&empty[s]); }
      if (kElect) __syncwarp();
    }
    if (kStatic) for (int c4 = 0; c4 < nsteps; c4 += 4) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (mode >= 2) { mbar_wait(&full[s], (c4 / 4) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        const uint64_t soff = (uint64_t)((s * kStage) >> 4);
        if (elect_one()) {
          mma_chunk_x3(tmem, id2N, idN, wh0 + soff, wl0 + soff, a0 + soff, 1);
          if (mode >= 1) commit(&empty[s]);
        }
        __syncwarp();
      }
    }
    if (kElect) { if (elect_one()) commit(done); } else commit(done);
    mbar_wait(done, 0);
    long long t1 = clock64();
    if (lane == 0) out[0] = (unsigned long long)(t1 - t0);
  } else if (warp >= 1 && warp <= 4 && mode >= 2) {
    const int lw = warp - 1;
    for (int c = lw; c < nsteps; c += 4) {
      const int s = c % nst;
      const int use = c / nst;
      if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
      if (mode == 2) { if (lane == 0) mbar_arrive(&full[s]); }
      else {
        if (lane == 0) mbar_expect(&full[s], stage_bytes);
        __syncwarp();
        const uint8_t* g = src + (size_t)(c % 64) * 81920;
        if (mode == 3) {
          if (lane == 0) bulk_g2s(sm + s * stage_bytes, g, 32768, &full[s]);
          if (lane == 1) bulk_g2s(sm + s * stage_bytes + 32768, g + 32768, 8192, &full[s]);
        } else {
          if (lane < 4) bulk_g2s(sm + s * stage_bytes + lane * 20480, g + lane * 20480, 20480, &full[s]);
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_real) : "memory");
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 64); unsigned long long h;
  uint8_t* src; cudaMalloc(&src, 64 * 81920); cudaMemset(src, 0, 64 * 81920);
  size_t smem = 5 * kStage + 2048;
  cudaFuncSetAttribute(k<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("mode,RT,cycles per chunk\n");
  for (int mode : {0, 3, 16, 17, 18, 19, 32, 33, 34, 35}) for (int RT : {32, 64}) {
    const int chunks = 4096;
    for (int rep = 0; rep < 2; ++rep) { if (mode & 32) k<true, true, true><<<1, 160, smem>>>(mode & 7, chunks, RT, src, d); else if (mode & 16) k<true, true, false><<<1, 160, smem>>>(mode & 7, chunks, RT, src, d); else if (mode & 8) k<true, false, false><<<1, 160, smem>>>(mode & 7, chunks, RT, src, d); else k<false, false, false><<<1, 160, smem>>>(mode, chunks, RT, src, d); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%d,%d,%.1f\n", mode, RT, h / (double)chunks);
  }
  return 0;
}
