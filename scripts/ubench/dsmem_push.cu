// Experiment: all-gather inside a thread-block cluster by pushing shared memory to the peers with
// cp.async.bulk.shared::cluster.shared::cta (completion signalled on the DESTINATION CTA's mbarrier).
// Every CTA pushes S bytes to each of the C CTAs of its cluster per round; one lane per destination.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t par) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t par) { long long t0 = clock64(); while (!mbar_try(bar, par)) { if (clock64() - t0 > 2000000000LL) __trap(); } }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void push(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
extern __shared__ __align__(1024) uint8_t smem[];
__global__ void k(int S, int rounds, int C, unsigned long long* ns_out, unsigned* check) {
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint8_t* outbox = smem + 1024;
  uint8_t* inbox = outbox + S;  // [C][S]
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < S / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(outbox)[i] = rank * 1000 + i;
  asm volatile("fence.proxy.async;" ::: "memory");
  cluster.sync();
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x == 0) mbar_expect(bar, (uint32_t)S * C);
    if (threadIdx.x < C) {
      const uint32_t dst = mapa(smem_u32(inbox + (size_t)rank * S), threadIdx.x);
      const uint32_t rb = mapa(smem_u32(bar), threadIdx.x);
      push(dst, smem_u32(outbox), S, rb);
    }
    if (threadIdx.x == 0) mbar_wait(bar, r & 1);
    cluster.sync();  // flow control for the benchmark: nobody overwrites an inbox that is still being checked
  }
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0) {
    ns_out[blockIdx.x] = t1 - t0;
    unsigned bad = 0;
    for (int c = 0; c < C; ++c) for (int i = 0; i < S / 4; i += 97) if (reinterpret_cast<uint32_t*>(inbox + (size_t)c * S)[i] != c * 1000 + i) ++bad;
    check[blockIdx.x] = bad;
  }
}
int main() {
  unsigned long long* ns; unsigned* chk; cudaMalloc(&ns, 1024 * 8); cudaMalloc(&chk, 1024 * 4);
  unsigned long long h[1024]; unsigned hc[1024];
  printf("cluster,S,grid,us per round (incl. cluster.sync),GB/s received per CTA,bad\n");
  int Cs[] = {8, 16};
  for (int ci = 0; ci < 2; ++ci) for (int S : {2048, 8192, 16384}) for (int grid : {0, 1}) {
    int C = Cs[ci];
    size_t sm = 1024 + (size_t)S * (C + 1);
    if (sm > 220 * 1024) continue;
    int g = grid == 0 ? C : (C == 8 ? 128 : 128);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (C > 8) cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = sm;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int rounds = 200;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, S, rounds, C, ns, chk);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%d,%d,%d,err %s\n", C, S, g, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    cudaMemcpy(h, ns, g * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hc, chk, g * 4, cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; unsigned bad = 0; for (int i = 0; i < g; ++i) { if (h[i] > mx) mx = h[i]; bad += hc[i]; }
    printf("%d,%d,%d,%.2f,%.1f,%u\n", C, S, g, mx / 1e3 / rounds, (double)S * C * rounds / mx, bad);
  }
  return 0;
}
