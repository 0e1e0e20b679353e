// Does cp.async.bulk.prefetch.L2 work?  Cold region (L2 flushed by a 512 MB write), optional prefetch of S bytes, a
// delay, then a timed cp.async.bulk load of the same bytes.  Also: prefetch.global.L2 per 128-byte line as alternative.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t par) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory"); }
extern __shared__ __align__(1024) uint8_t smem[];
// mode 0: no prefetch; 1: bulk prefetch; 2: prefetch.global.L2 by 32 lanes, one per 128-byte line
__global__ void k(const uint8_t* src, int S, int mode, int delay_ns, long long* out) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint8_t* buf = smem + 1024;
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  src += (size_t)blockIdx.x * S;
  if (mode == 1 && lane == 0) bulk_prefetch_l2(src, S);
  if (mode == 2) for (int o = lane * 128; o < S; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); } while (t - t0 < (unsigned long long)delay_ns);
  __syncwarp();
  const long long c0 = clock64();
  if (lane == 0) { mbar_expect(bar, S); bulk_g2s(buf, src, S, bar); }
  while (!mbar_try(bar, 0)) {}
  const long long c1 = clock64();
  if (lane == 0) out[blockIdx.x] = c1 - c0;
}
int main() {
  const size_t span = (size_t)64 << 20;
  uint8_t* src; cudaMalloc(&src, span); cudaMemset(src, 1, span);
  uint8_t* flush; cudaMalloc(&flush, (size_t)512 << 20);
  long long* d; cudaMalloc(&d, 1024 * 8); long long h[1024];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("grid,S,mode,delay_ns,load cycles median,max\n");
  for (int grid : {1, 128}) for (int S : {32768}) for (int delay : {3000, 10000}) for (int mode = 0; mode < 3; ++mode) {
    long long med = 0, mx = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemset(flush, rep, (size_t)512 << 20);
      cudaDeviceSynchronize();
      k<<<grid, 32, 100 * 1024>>>(src + (size_t)rep * (8 << 20), S, mode, delay, d);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
      cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      // crude: median over CTAs of the last rep
      for (int i = 0; i < grid; ++i) for (int j = i + 1; j < grid; ++j) if (h[j] < h[i]) { long long tt = h[i]; h[i] = h[j]; h[j] = tt; }
      med = h[grid / 2]; mx = h[grid - 1];
    }
    printf("%d,%d,%d,%d,%lld,%lld\n", grid, S, mode, delay, med, mx);
  }
  return 0;
}
