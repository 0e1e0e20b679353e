// Split-K reduction between the two CTAs of a cluster: how long does it take to hand half of an fp32 accumulator tile
// (128 features x 32 rows = 16 KB) to the peer CTA through distributed shared memory and to synchronise?
//   variant 0: st.shared::cluster (16-byte stores from registers) + barrier.cluster
//   variant 1: same stores + remote mbarrier arrive (release.cluster) / local wait  (no full cluster barrier)
//   variant 2: local st.shared only + __syncthreads (the no-exchange baseline)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_reduce dsmem_reduce.cu ; run: ./dsmem_reduce
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int VARIANT>
__global__ void __launch_bounds__(128) k(float* out, long long* cycles, int iters) {
  __shared__ __align__(16) float buf[2][128 * 32];  // double buffered [feature][32 rows]
  __shared__ __align__(8) unsigned long long bar[2];
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x;
  const uint32_t peer = cl.block_rank() ^ 1;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(128));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cl.sync();
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = tid * 0.001f + i;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int b = it & 1;
    const uint32_t local = smem_u32(&buf[b][tid * 32]);
    const uint32_t remote = VARIANT == 2 ? local : mapa(local, peer);
#pragma unroll
    for (int i = 0; i < 32; i += 4) st_cluster_v4(remote + i * 4, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    if (VARIANT == 0) {
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else if (VARIANT == 1) {
      const uint32_t rbar = mapa(smem_u32(&bar[b]), peer);
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
      uint32_t ok = 0;
      const uint32_t parity = (it >> 1) & 1;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"(parity) : "memory");
      }
    } else {
      __syncthreads();
    }
    // consume what the peer sent: the same layout, own feature
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 r = *reinterpret_cast<const float4*>(&buf[b][tid * 32 + i]);
      acc += r.x + r.y + r.z + r.w;
      v[i] += 1e-6f * r.x;
    }
  }
  const long long t1 = clock64();
  cl.sync();
  if (tid == 0) cycles[blockIdx.x] = (t1 - t0) / iters;
  out[blockIdx.x * 128 + tid] = acc;
}

template <int VARIANT>
void run(int grid, const char* name) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, grid * 128 * 4);
  cudaMalloc(&cyc, grid * 8);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int iters = 2000;
  cudaLaunchKernelEx(&cfg, k<VARIANT>, out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  long long mn = 1 << 30, mx = 0, sum = 0;
  for (int i = 0; i < grid; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
  printf("%-44s grid %3d: cycles per 16 KB hand-over  min %lld  mean %lld  max %lld   (%s)\n", name, grid, mn, sum / grid, mx, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int grid : {2, 128}) {
    run<2>(grid, "local st.shared + __syncthreads (baseline)");
    run<0>(grid, "st.shared::cluster + barrier.cluster");
    run<1>(grid, "st.shared::cluster + remote mbarrier arrive");
  }
  return 0;
}
