// How long does the ISSUE of a cp.async.bulk take for the issuing thread (clock64 around the instruction), and when
// does the data land?  One CTA, one warp; S bytes per copy; L lanes of the warp issue one copy each in the same
// divergent region.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t par) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory"); }
extern __shared__ __align__(1024) uint8_t smem[];
__global__ void k(const uint8_t* src, int S, int L, int reps, long long* out) {  // blockDim.x / 32 warps stream concurrently
  const int warp = threadIdx.x >> 5;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem) + warp;
  uint8_t* buf = smem + 1024 + warp * (S * L);
  const int lane = threadIdx.x & 31;
  src += (size_t)warp * 65536;
  if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  long long t_exp = 0, t_issue = 0, t_land = 0;
  for (int r = 0; r < reps; ++r) {
    const uint8_t* s = src + (size_t)(r % 16) * 262144 + (size_t)lane * S;
    __syncwarp();
    const long long t0 = clock64();
    if (lane == 0) mbar_expect(bar, (uint32_t)(S * L));
    const long long t1 = clock64();
    if (lane < L) bulk_g2s(buf + lane * S, s, S, bar);
    __syncwarp();
    const long long t2 = clock64();
    while (!mbar_try(bar, r & 1)) {}
    const long long t3 = clock64();
    if (r >= 4) { t_exp += t1 - t0; t_issue += t2 - t1; t_land += t3 - t2; }
  }
  if (threadIdx.x == 0) { out[0] = t_exp / (reps - 4); out[1] = t_issue / (reps - 4); out[2] = t_land / (reps - 4); }
}
int main() {
  uint8_t* src; cudaMalloc(&src, 16 * 262144 + 65536 * 8); cudaMemset(src, 1, 16 * 262144 + 65536 * 8);
  long long* d; cudaMalloc(&d, 64); long long h[3];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("warps,S,lanes,expect_tx cycles,issue cycles,issue->landed cycles\n");
  for (int W : {1, 2, 4}) for (int S : {8192, 20480, 32768}) for (int L : {1, 2}) {
    if (S * L * W > 190 * 1024) continue;
    k<<<1, 32 * W, 200 * 1024>>>(src, S, L, 68, d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("%d,%d,%d,%lld,%lld,%lld\n", W, S, L, h[0], h[1], h[2]);
  }
  return 0;
}
