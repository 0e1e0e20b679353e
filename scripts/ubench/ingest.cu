// Microbenchmark: how fast can ONE SM pull bytes from L2/HBM into shared memory, by mechanism?
//   mode 0: cp.async.bulk (1-D bulk TMA), copy size S, D copies in flight per CTA
//   mode 1: cp.async 16 B (LDGSTS) from 256 threads, groups in flight
//   mode 2: ld.global.v4 -> st.shared from 256 threads (unrolled x8)
// Usage: ingest <grid> <ctas_per_sm_hint: smem KB per CTA> ; prints GB/s per CTA and total.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t par) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory"); }

extern __shared__ __align__(1024) uint8_t smem[];

// each CTA reads `per_cta` bytes starting at base + cta*stride (wraps inside `span` bytes)
__global__ void k_bulk(const uint8_t* base, size_t span, size_t per_cta, int S, int D, unsigned long long* ns_out) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* buf = smem + 1024;
  if (threadIdx.x == 0) { for (int i = 0; i < D; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    const size_t n = per_cta / S;
    size_t off = ((size_t)blockIdx.x * per_cta) % span;
    for (size_t i = 0; i < n + D; ++i) {
      const int s = i % D;
      if (i >= (size_t)D) { const uint32_t par = ((i / D) - 1) & 1; while (!mbar_try(&bars[s], par)) {} }
      if (i < n) { mbar_expect(&bars[s], S); bulk_g2s(buf + (size_t)s * S, base + off, S, &bars[s]); off += S; if (off + S > span) off = 0; }
    }
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    ns_out[blockIdx.x] = t1 - t0;
  }
}

__global__ void k_ldg(const uint8_t* base, size_t span, size_t per_cta, unsigned long long* ns_out) {
  uint4* buf = reinterpret_cast<uint4*>(smem + 1024);
  unsigned long long t0, t1;
  __syncthreads();
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  const size_t off0 = ((size_t)blockIdx.x * per_cta) % span;
  const uint4* src = reinterpret_cast<const uint4*>(base + off0);
  const size_t n16 = per_cta / 16;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = threadIdx.x; i + 7 * blockDim.x < n16; i += 8 * blockDim.x) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldcg(src + i + j * blockDim.x);
#pragma unroll
    for (int j = 0; j < 8; ++j) buf[(threadIdx.x + j * blockDim.x) % 4096] = v[j];
  }
  __syncthreads();
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0) ns_out[blockIdx.x] = t1 - t0 + (acc.x & 0);
}

int main(int argc, char** argv) {
  const size_t span = (size_t)64 << 20;  // 64 MB: L2 resident after the first pass
  uint8_t* d; cudaMalloc(&d, span); cudaMemset(d, 1, span);
  unsigned long long* ns; cudaMalloc(&ns, 1024 * 8);
  unsigned long long h[1024];
  const size_t per_cta = (size_t)8 << 20;
  int grids[] = {1, 16, 74, 148, 296};
  printf("mode,grid,S,D,smemKB,GB/s per CTA (median),GB/s total\n");
  for (int gi = 0; gi < 5; ++gi) {
    int grid = grids[gi];
    int Ss[] = {2048, 8192, 16384, 32768};
    for (int si = 0; si < 4; ++si) {
      int Ds[] = {1, 2, 4, 6, 12};
      for (int di = 0; di < 5; ++di) {
        int S = Ss[si], D = Ds[di];
        size_t sm = 1024 + (size_t)S * D;
        if (sm > 100 * 1024) continue;
        if (grid == 296 && sm > 100 * 1024) continue;
        cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        for (int rep = 0; rep < 2; ++rep) k_bulk<<<grid, 32, sm>>>(d, span, per_cta, S, D, ns);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, ns, grid * 8, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
        printf("bulk,%d,%d,%d,%zu,%.1f,%.1f\n", grid, S, D, sm / 1024, per_cta / (double)h[grid / 2], grid * per_cta / (double)mx);
      }
    }
    for (int threads = 128; threads <= 512; threads *= 2) {
      size_t sm = 1024 + 65536;
      cudaFuncSetAttribute(k_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      for (int rep = 0; rep < 2; ++rep) k_ldg<<<grid, threads, sm>>>(d, span, per_cta, ns);
      cudaDeviceSynchronize();
      cudaMemcpy(h, ns, grid * 8, cudaMemcpyDeviceToHost);
      unsigned long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
      printf("ldg%d,%d,16,8,%zu,%.1f,%.1f\n", threads, grid, sm / 1024, per_cta / (double)h[grid / 2], grid * per_cta / (double)mx);
    }
  }
  return 0;
}
