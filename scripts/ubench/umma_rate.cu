// How long does the tensor core take per 64-wide k-chunk (4 k16 steps x 3 products, M = 128, both operands in shared
// memory, K-major SWIZZLE_128B) as a function of N?  One CTA, operands resident (no loads), many repetitions.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t par) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0; d |= (uint64_t)((saddr >> 4) & 0x3fff); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
extern __shared__ __align__(1024) uint8_t smem_raw[];
// mode 0: 3 products per k16 (wl*ah, wh*al, wh*ah) with N columns; mode 1: stacked: wh*[ah;al] (2N columns) + wl*ah (N)
__global__ void k(int N, int M, int mode, int chunks, unsigned long long* out) {
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(sm + 200 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idN = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t id2N = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t wh = smem_u32(sm), wl = wh + 16384, ah = wl + 16384, al = ah + N * 128;  // [ah; al] contiguous
    long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) {
      for (int kk = 0; kk < 4; ++kk) {
        if (mode == 0) {
          umma(tmem, make_desc(wl + kk * 32), make_desc(ah + kk * 32), idN, 1);
          umma(tmem, make_desc(wh + kk * 32), make_desc(al + kk * 32), idN, 1);
          umma(tmem, make_desc(wh + kk * 32), make_desc(ah + kk * 32), idN, 1);
        } else if (mode == 1) {
          umma(tmem, make_desc(wh + kk * 32), make_desc(ah + kk * 32), id2N, 1);
          umma(tmem, make_desc(wl + kk * 32), make_desc(ah + kk * 32), idN, 1);
        } else if (mode == 2) {  // three independent accumulators
          umma(tmem, make_desc(wl + kk * 32), make_desc(ah + kk * 32), idN, 1);
          umma(tmem + N, make_desc(wh + kk * 32), make_desc(al + kk * 32), idN, 1);
          umma(tmem + 2 * N, make_desc(wh + kk * 32), make_desc(ah + kk * 32), idN, 1);
        } else {  // six: also alternate by k16 parity
          const uint32_t o = (kk & 1) * 3 * N;
          umma(tmem + o, make_desc(wl + kk * 32), make_desc(ah + kk * 32), idN, 1);
          umma(tmem + o + N, make_desc(wh + kk * 32), make_desc(al + kk * 32), idN, 1);
          umma(tmem + o + 2 * N, make_desc(wh + kk * 32), make_desc(ah + kk * 32), idN, 1);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    while (!mbar_try(bar, 0)) {}
    long long t1 = clock64();
    out[0] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 64); unsigned long long h;
  size_t smem = 202 * 1024 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("M,N,mode,cycles per chunk (4 k16 steps),cycles per MMA\n");
  for (int M : {128}) for (int mode = 0; mode < 4; ++mode) for (int N : {16, 32, 64, 128, 256}) {
    if (mode == 1 && 2 * N > 256) continue;
    if (mode == 2 && 3 * N > 512) continue;
    if (mode == 3 && 6 * N > 512) continue;
    const int chunks = 2000;
    k<<<1, 128, smem>>>(N, M, mode, chunks, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%d,%d,%d,%.1f,%.1f\n", M, N, mode, h / (double)chunks, h / (double)chunks / (mode == 0 ? 12 : 8));
  }
  return 0;
}
