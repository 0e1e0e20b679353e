// Microbenchmark 2: does issuing cp.async.bulk from several threads / warps of one CTA run copies concurrently?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t par) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory"); }
extern __shared__ __align__(1024) uint8_t smem[];
// T issuing threads (thread 32*w lane 0 if warps!=0 else lanes 0..T-1 of warp 0); each streams per_thr bytes with copies
// of S bytes, D in flight, into its own smem region.  split=1: one logical copy of S bytes is issued as P pieces.
__global__ void k(const uint8_t* base, size_t span, size_t per_thr, int S, int D, int T, int by_warp, int P, unsigned long long* ns_out) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* buf = smem + 1024;
  if (threadIdx.x == 0) { for (int i = 0; i < D * T; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  int me = -1;
  if (by_warp) { if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < T) me = threadIdx.x >> 5; }
  else if (threadIdx.x < T) me = threadIdx.x;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  if (me >= 0) {
    const size_t n = per_thr / S;
    size_t off = (((size_t)blockIdx.x * T + me) * per_thr) % span;
    uint64_t* mb = bars + me * D;
    uint8_t* mybuf = buf + (size_t)me * D * S;
    for (size_t i = 0; i < n + D; ++i) {
      const int s = i % D;
      if (i >= (size_t)D) { const uint32_t par = ((i / D) - 1) & 1; while (!mbar_try(&mb[s], par)) {} }
      if (i < n) {
        mbar_expect(&mb[s], S);
        for (int q = 0; q < P; ++q) bulk_g2s(mybuf + (size_t)s * S + q * (S / P), base + off + q * (S / P), S / P, &mb[s]);
        off += S; if (off + S > span) off = 0;
      }
    }
  }
  __syncthreads();
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0) ns_out[blockIdx.x] = t1 - t0;
}
int main() {
  const size_t span = (size_t)64 << 20;
  uint8_t* d; cudaMalloc(&d, span); cudaMemset(d, 1, span);
  unsigned long long* ns; cudaMalloc(&ns, 1024 * 8);
  unsigned long long h[1024];
  printf("grid,S,D,T,by_warp,P,GB/s per CTA,GB/s total\n");
  int grids[] = {1, 148};
  for (int gi = 0; gi < 2; ++gi) for (int by_warp = 0; by_warp < 2; ++by_warp) {
    int cfg[][4] = {{16384,2,1,1},{16384,2,2,1},{16384,2,4,1},{16384,1,4,1},{16384,1,6,1},{8192,2,4,1},{8192,1,8,1},{4096,2,8,1},{65536,1,1,1},{65536,2,1,1},{32768,2,2,1},{32768,1,3,1},{16384,2,1,2},{16384,2,1,4},{32768,2,1,4},{65536,2,1,8},{49152,2,1,1},{49152,2,2,1},{32768,1,4,1},{32768,1,6,1},{16384,1,8,1},{16384,1,12,1},{8192,1,16,1},{8192,2,12,1},{40960,1,4,1},{40960,1,5,1},{20480,1,8,1},{20480,2,4,1}};
    for (auto& c : cfg) {
      int S = c[0], D = c[1], T = c[2], P = c[3];
      size_t sm = 1024 + (size_t)S * D * T;
      if (sm > 200 * 1024) continue;
      size_t per_thr = (((size_t)16 << 20) / T / S) * S;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      int grid = grids[gi];
      for (int rep = 0; rep < 2; ++rep) k<<<grid, 256, sm>>>(d, span, per_thr, S, D, T, by_warp, P, ns);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, ns, grid * 8, cudaMemcpyDeviceToHost);
      unsigned long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
      size_t tot = (per_thr / S) * S * T;
      printf("%d,%d,%d,%d,%d,%d,%.1f,%.1f\n", grid, S, D, T, by_warp, P, tot / (double)h[grid / 2], grid * (double)tot / mx);
    }
  }
  return 0;
}
