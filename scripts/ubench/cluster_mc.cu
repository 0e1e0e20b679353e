// Experiment: inside a cluster of 8 CTAs
//  (1) producer-side push through L2: bulk store smem->global, wait, then ONE multicast bulk load global->smem of all 8
//      CTAs (cp.async.bulk...multicast::cluster) completing on the same-offset mbarrier of every destination CTA;
//  (2) tcgen05.commit with .multicast::cluster as an 8-way "stage free" signal.
// Measures the latency of a full 8-way all-gather of S bytes per CTA done that way, and checks the data.
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
extern __shared__ __align__(1024) uint8_t smem[];
__global__ void __cluster_dims__(8, 1, 1) k(uint8_t* scratch, int S, int rounds, unsigned long long* ns_out, unsigned* check) {
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const uint32_t cid = blockIdx.x / 8;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);       // all 8 chunks landed
  uint64_t* freeb = full + 1;                               // all 8 CTAs have consumed (multicast commit target)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* outbox = smem + 1024;
  uint8_t* inbox = outbox + S;  // [8][S]
  uint8_t* mine = scratch + ((size_t)cid * 8 + rank) * S;
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(freeb, 8); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(tslot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  cluster.sync();
  unsigned long long t0, t1;
  unsigned bad = 0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (int r = 0; r < rounds; ++r) {
    for (int i = threadIdx.x; i < S / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(outbox)[i] = (r << 20) + rank * 10000 + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect(full, (uint32_t)S * 8);
      // 1) publish through L2
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine), "r"(smem_u32(outbox)), "r"(S) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      // 2) wait until every CTA of the cluster has released the inbox of the previous round
      if (r > 0) mbar_wait(freeb, (r - 1) & 1);
      // 3) one multicast load: my chunk into slot [rank] of all 8 inboxes, completing on every CTA's `full`
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                   ::"r"(smem_u32(inbox + (size_t)rank * S)), "l"(mine), "r"(S), "r"(smem_u32(full)), "h"((uint16_t)0xff) : "memory");
    }
    if (threadIdx.x == 0) mbar_wait(full, r & 1);
    __syncthreads();
    if (threadIdx.x < 8) {  // check a few words of every chunk
      const uint32_t* w = reinterpret_cast<const uint32_t*>(inbox + (size_t)threadIdx.x * S);
      for (int i = 0; i < S / 4; i += 61) if (w[i] != (uint32_t)((r << 20) + threadIdx.x * 10000 + i)) ++bad;
    }
    __syncthreads();
    // 4) "consumed": multicast commit -> arrives on `freeb` of all 8 CTAs
    if (threadIdx.x == 0)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(freeb)), "h"((uint16_t)0xff) : "memory");
  }
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
  unsigned tot = __syncthreads_count(bad != 0);
  if (threadIdx.x == 0) { ns_out[blockIdx.x] = t1 - t0; check[blockIdx.x] = tot; }
  cluster.sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(*tslot) : "memory");
}
int main() {
  unsigned long long* ns; unsigned* chk; uint8_t* scratch;
  cudaMalloc(&ns, 1024 * 8); cudaMalloc(&chk, 1024 * 4); cudaMalloc(&scratch, 64 << 20);
  unsigned long long h[1024]; unsigned hc[1024];
  printf("S,grid,us per round,bad\n");
  for (int S : {2048, 8192, 16384}) for (int g : {8, 128}) {
    size_t sm = 1024 + (size_t)S * 9;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int rounds = 300;
    k<<<g, 128, sm>>>(scratch, S, rounds, ns, chk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%d,%d,err %s\n", S, g, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, ns, g * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hc, chk, g * 4, cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; unsigned bad = 0; for (int i = 0; i < g; ++i) { if (h[i] > mx) mx = h[i]; bad += hc[i]; }
    printf("%d,%d,%.2f,%u\n", S, g, mx / 1e3 / rounds, bad);
  }
  int nc = 0; cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(128); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, k, &cfg);
  printf("max active clusters of 8 (320 threads, 200 KB smem): %d (%s)\n", nc, cudaGetErrorString(e));
  return 0;
}
