// Experiment: one CTA, D[128 x N] (fp32, TMEM) = sum over K chunks of W[128 x 64] * X[N x 64]^T with bf16 head/tail
// operands (3 products), operands in shared memory in the K-major SWIZZLE_128B layout written by the host.
// Validates: smem descriptors, instruction descriptor, K advance inside the swizzle atom, tcgen05.commit -> mbarrier,
// tcgen05.ld 32x32b layout.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t par) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t par) { long long t0 = clock64(); while (!mbar_try(bar, par)) { if (clock64() - t0 > 2000000000LL) __trap(); } }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory"); }

// K-major, SWIZZLE_128B, rows of 64 bf16 (128 B), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);          // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset = 1024 B
  d |= (uint64_t)1 << 46;                          // version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// kind::f16, A = B = bf16, D = f32, K-major both, M = 128, N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

template <int N>
__global__ void __launch_bounds__(192) k(const uint8_t* w, const uint8_t* x, float* out, int nchunks) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int WB = 128 * 128;           // one bf16 plane of a W chunk [128][64]
  constexpr int XB = N * 128;             // one plane of an X chunk [N][64]
  uint8_t* sw = sm;                       // [head | tail]
  uint8_t* sx = sm + 2 * WB;              // [head | tail]
  uint64_t* bars = (uint64_t*)(sm + 2 * WB + 2 * XB);  // [0] full, [1] mma done
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = make_idesc(128, N);
  for (int c = 0; c < nchunks; ++c) {
    if (warp == 5 && lane == 0) {
      if (c > 0) mbar_wait(&bars[1], (c - 1) & 1);     // previous MMAs have read the operands
      mbar_expect(&bars[0], 2 * WB + 2 * XB);
      bulk_g2s(sw, w + (size_t)c * 2 * WB, 2 * WB, &bars[0]);
      bulk_g2s(sx, x + (size_t)c * 2 * XB, 2 * XB, &bars[0]);
    }
    if (warp == 4 && lane == 0) {
      mbar_wait(&bars[0], c & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aw = smem_u32(sw), ax = smem_u32(sx);
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t wh = make_desc(aw + kk * 32), wl = make_desc(aw + WB + kk * 32);
        const uint64_t xh = make_desc(ax + kk * 32), xl = make_desc(ax + XB + kk * 32);
        umma(tmem, wl, xh, idesc, (c | kk) != 0);
        umma(tmem, wh, xl, idesc, 1);
        umma(tmem, wh, xh, idesc, 1);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
      if (c == nchunks - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[2])) : "memory");
    }
  }
  if (warp < 4) {
    mbar_wait(&bars[2], 0);  // committed once, after the last chunk
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[N];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    if constexpr (N == 32) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                     "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(taddr));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int f = warp * 32 + lane;
    for (int j = 0; j < N; ++j) out[(size_t)f * N + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64) : "memory");
}

static uint16_t bf16_rn(float f) { uint32_t x; memcpy(&x, &f, 4); uint32_t lsb = (x >> 16) & 1; x += 0x7fff + lsb; return (uint16_t)(x >> 16); }
static float bf16_f(uint16_t b) { uint32_t x = (uint32_t)b << 16; float f; memcpy(&f, &x, 4); return f; }
static uint32_t tile_off(int row, int k) { return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2); }

int main() {
  constexpr int N = 32, NC = 16, K = NC * 64;
  std::vector<float> W(128 * K), X(N * K);
  srand(1);
  for (auto& v : W) v = (rand() / (float)RAND_MAX - 0.5f) / 16.f;
  for (auto& v : X) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  std::vector<uint8_t> wp((size_t)NC * 2 * 128 * 128), xp((size_t)NC * 2 * N * 128);
  for (int c = 0; c < NC; ++c) {
    for (int r = 0; r < 128; ++r) for (int e = 0; e < 64; ++e) {
      float v = W[(size_t)r * K + c * 64 + e]; uint16_t h = bf16_rn(v), l = bf16_rn(v - bf16_f(h));
      memcpy(&wp[(size_t)c * 2 * 16384 + tile_off(r, e)], &h, 2); memcpy(&wp[(size_t)c * 2 * 16384 + 16384 + tile_off(r, e)], &l, 2);
    }
    for (int r = 0; r < N; ++r) for (int e = 0; e < 64; ++e) {
      float v = X[(size_t)r * K + c * 64 + e]; uint16_t h = bf16_rn(v), l = bf16_rn(v - bf16_f(h));
      memcpy(&xp[(size_t)c * 2 * N * 128 + tile_off(r, e)], &h, 2); memcpy(&xp[(size_t)c * 2 * N * 128 + N * 128 + tile_off(r, e)], &l, 2);
    }
  }
  uint8_t *dw, *dx; float* dout;
  cudaMalloc(&dw, wp.size()); cudaMalloc(&dx, xp.size()); cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(dw, wp.data(), wp.size(), cudaMemcpyHostToDevice); cudaMemcpy(dx, xp.data(), xp.size(), cudaMemcpyHostToDevice);
  size_t smem = 2 * 16384 + 2 * N * 128 + 64 + 1024;
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<N><<<1, 192, smem>>>(dw, dx, dout, NC);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> out(128 * N);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxerr32 = 0, maxref = 0;
  for (int f = 0; f < 128; ++f) for (int j = 0; j < N; ++j) {
    double ref = 0, ref3 = 0;
    for (int kk = 0; kk < K; ++kk) {
      float w = W[(size_t)f * K + kk], x = X[(size_t)j * K + kk];
      ref += (double)w * x;
      float wh = bf16_f(bf16_rn(w)), wl = bf16_f(bf16_rn(w - wh)), xh = bf16_f(bf16_rn(x)), xl = bf16_f(bf16_rn(x - xh));
      ref3 += (double)wh * xh + (double)wh * xl + (double)wl * xh;
    }
    maxerr = fmax(maxerr, fabs(out[f * N + j] - ref3)); maxerr32 = fmax(maxerr32, fabs(out[f * N + j] - ref)); maxref = fmax(maxref, fabs(ref));
  }
  printf("max |D - bf16x3 exact| = %.3e, max |D - fp64 exact| = %.3e, max |ref| = %.3f, D[0][0..3] = %f %f %f %f\n", maxerr, maxerr32, maxref, out[0], out[1], out[2], out[3]);
  return 0;
}
