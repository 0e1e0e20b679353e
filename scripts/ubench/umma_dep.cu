// Is the ~70-80 cycle cost of a small-N tcgen05.mma a dependency latency (same accumulator) or a throughput floor?
// 8 MMAs per chunk (N-stacked form: wh x [a; a'] with N2, wl x a with N1), fully unrolled asm, descriptors precomputed.
// variant 0: one accumulator; 1: two accumulators alternating by k16 step; 2: four accumulators.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t par) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0; d |= (uint64_t)((saddr >> 4) & 0x3fff); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
#define MMA(D, A, B, I) "tcgen05.mma.cta_group::1.kind::f16 [" D "], " A ", " B ", " I ", p1;\n\t"
template <int V>
__device__ __forceinline__ void chunk(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t i2, uint32_t i1, uint64_t wh, uint64_t wl, uint64_t a) {
  if (V == 0)
    asm volatile("{\n\t.reg .pred p1;\n\t.reg .b64 wh, wl, a;\n\tsetp.eq.b32 p1, %0, %0;\n\tmov.b64 wh, %6;\n\tmov.b64 wl, %7;\n\tmov.b64 a, %8;\n\t"
                 MMA("%0", "wh", "a", "%4") MMA("%0", "wl", "a", "%5") "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
                 MMA("%0", "wh", "a", "%4") MMA("%0", "wl", "a", "%5") "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
                 MMA("%0", "wh", "a", "%4") MMA("%0", "wl", "a", "%5") "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
                 MMA("%0", "wh", "a", "%4") MMA("%0", "wl", "a", "%5") "}\n"
                 ::"r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(i2), "r"(i1), "l"(wh), "l"(wl), "l"(a) : "memory");
  else if (V == 1)
    asm volatile("{\n\t.reg .pred p1;\n\t.reg .b64 wh, wl, a;\n\tsetp.eq.b32 p1, %0, %0;\n\tmov.b64 wh, %6;\n\tmov.b64 wl, %7;\n\tmov.b64 a, %8;\n\t"
                 MMA("%0", "wh", "a", "%4") MMA("%0", "wl", "a", "%5") "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
                 MMA("%1", "wh", "a", "%4") MMA("%1", "wl", "a", "%5") "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
                 MMA("%0", "wh", "a", "%4") MMA("%0", "wl", "a", "%5") "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
                 MMA("%1", "wh", "a", "%4") MMA("%1", "wl", "a", "%5") "}\n"
                 ::"r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(i2), "r"(i1), "l"(wh), "l"(wl), "l"(a) : "memory");
  else
    asm volatile("{\n\t.reg .pred p1;\n\t.reg .b64 wh, wl, a;\n\tsetp.eq.b32 p1, %0, %0;\n\tmov.b64 wh, %6;\n\tmov.b64 wl, %7;\n\tmov.b64 a, %8;\n\t"
                 MMA("%0", "wh", "a", "%4") MMA("%1", "wl", "a", "%5") "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
                 MMA("%2", "wh", "a", "%4") MMA("%3", "wl", "a", "%5") "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
                 MMA("%0", "wh", "a", "%4") MMA("%1", "wl", "a", "%5") "add.s64 wh, wh, 2;\n\tadd.s64 wl, wl, 2;\n\tadd.s64 a, a, 2;\n\t"
                 MMA("%2", "wh", "a", "%4") MMA("%3", "wl", "a", "%5") "}\n"
                 ::"r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(i2), "r"(i1), "l"(wh), "l"(wl), "l"(a) : "memory");
}
extern __shared__ __align__(1024) uint8_t smem_raw[];
template <int V>
__global__ void k(int N, int chunks, unsigned long long* out) {
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(sm + 100 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (warp == 0 && lane == 0) {
    const uint32_t i1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t i2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | (8u << 24);
    const uint64_t wh = make_desc(smem_u32(sm)), wl = make_desc(smem_u32(sm) + 16384), a = make_desc(smem_u32(sm) + 32768);
    long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) chunk<V>(tmem, tmem + 128, tmem + 256, tmem + 384, i2, i1, wh, wl, a);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    while (!mbar_try(bar, 0)) {}
    out[0] = (unsigned long long)(clock64() - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 64); unsigned long long h;
  size_t smem = 102 * 1024 + 1024;
  printf("N,variant,cycles per chunk (8 MMAs),cycles per MMA\n");
  for (int N : {16, 32, 64}) for (int v = 0; v < 3; ++v) {
    const int chunks = 4000;
    if (v == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<0><<<1, 128, smem>>>(N, chunks, d); }
    if (v == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<1><<<1, 128, smem>>>(N, chunks, d); }
    if (v == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<2><<<1, 128, smem>>>(N, chunks, d); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%d,%d,%.1f,%.1f\n", N, v, h / (double)chunks, h / (double)chunks / 8);
  }
  return 0;
}
