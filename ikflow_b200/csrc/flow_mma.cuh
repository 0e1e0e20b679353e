// Flow engine "mma": warp-level mma.sync tensor-core tiles, 64-feature CTA tiles, teams of hidden/64 CTAs, two CTAs
// per SM.  Handles every supported architecture (hidden a multiple of 64); see flow.cu for the algorithm.
#pragma once

#include "flow_common.cuh"

namespace ikf {

template <int RT>
struct FlowCfg {
  static constexpr int kMI = RT / 32;                 // m16 tiles per warp (warp tile = 16*kMI rows x 32 features)
  static constexpr int kATileBytes = RT * kKC * 2;    // one bf16 plane of an activation chunk
  static constexpr int kAChunkBytes = 2 * kATileBytes;
  static constexpr int kStageBytes = kAChunkBytes + kWChunkBytes;
  static constexpr int kStages = RT == 64 ? 2 : 3;
};

template <int RT>
struct __align__(1024) FlowSmem {
  using C = FlowCfg<RT>;
  uint8_t ring[C::kStages][C::kStageBytes];  // [activation head|tail][weight head|tail]
  uint8_t staging[C::kAChunkBytes];          // outgoing activation chunk; reused as the last layer's warp partials
  float small[2][kSmallFloats];
  float u[RT][kPad];   // flow state
  float cnd[RT][8];    // condition
  float a[RT][kPad];   // output of the last layer of the current subnet
  uint64_t full[C::kStages], empty[C::kStages];
  uint64_t small_full[2], small_empty[2];
  uint64_t staging_free;
};

// ---------------------------------------------------------------------------------------------------------------------
// the kernel

template <int RT>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) flow_inverse_kernel(const FlowParams p) {
  using C = FlowCfg<RT>;
  constexpr int MI = C::kMI;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  FlowSmem<RT>& sm =
      *reinterpret_cast<FlowSmem<RT>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int NT = p.NT;
  const int slot = blockIdx.x / NT;
  const int t = blockIdx.x % NT;
  const uint32_t launch_id = p.epoch;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], kComputeWarps);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.small_full[b], 1);
      mbar_init(&sm.small_empty[b], kComputeWarps);
    }
    mbar_init(&sm.staging_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  uint8_t* act_slot = p.act + (size_t)slot * 2 * NT * kAChunkStride;
  float* part_slot = p.partial + (size_t)slot * 2 * NT * kRTMax * kPad;
  uint32_t* aflag = p.act_flag + (size_t)slot * 2 * NT;
  uint32_t* pflag = p.part_flag + (size_t)slot * 2 * NT;

  const int n_blocks = p.block_first - p.block_last + 1;
  const int steps_per_rg = 2 * n_blocks;  // subnets per row group
  const int my_rgs = (p.n_rowgroups - slot + p.slots - 1) / p.slots;
  const int total_steps = my_rgs * steps_per_rg;

  // The three roles walk the same schedule: step g = (row group, block, subnet); within a step the layers in order.
  // Activation exchange number x uses scratch buffer x % 2; its flags carry epoch + 1 + (writes so far to that buffer).

  if (warp == kLoaderWarp) {
    // ===== loader: bulk-TMA producer for the small-parameter blocks and the weight/activation ring =====
    // The whole warp polls (lane c watches the flag of chunk c, relaxed loads); lane 0 issues the copies.  Weight
    // chunks are issued as soon as their stage is free, the activation chunk of a stage follows when its producer has
    // published it -- in the fixed order c = t, t+1, ... so that the fp32 accumulation order (and therefore the
    // result) never depends on timing.
    uint32_t ring_pos = 0;
    uint32_t act_w[2] = {0, 0};  // writes so far into each activation scratch buffer
    uint32_t xchg = 0;           // activation exchanges so far
    auto prefetch_small = [&](int g) {
      if (g >= total_steps) return;
      const int b = g & 1;
      if (g >= 2) mbar_wait(&sm.small_empty[b], ((g >> 1) - 1) & 1);
      if (lane == 0) {
        const int in_rg = g % steps_per_rg;
        const int blk = p.block_first - in_rg / 2;
        const int n = 2 * blk + (in_rg & 1);
        mbar_arrive_expect_tx(&sm.small_full[b], kSmallBytes);
        bulk_g2s(sm.small[b], p.small + ((size_t)n * NT + t) * kSmallFloats, kSmallBytes, &sm.small_full[b]);
      }
      __syncwarp();
    };
    prefetch_small(0);
    for (int g = 0; g < total_steps; ++g) {
      const int in_rg = g % steps_per_rg;
      const int blk = p.block_first - in_rg / 2;
      const int n = 2 * blk + (in_rg & 1);
      if (p.n_big == 0) prefetch_small(g + 1);
      for (int l = 0; l < p.n_big; ++l) {
        // input of hidden layer l = exchange number xchg (written by the previous layer of every team member)
        const int buf = xchg & 1;
        const uint32_t expected = p.epoch + 1 + act_w[buf];
        const uint8_t* wbase =
            reinterpret_cast<const uint8_t*>(p.big_w) + (((size_t)n * p.n_big + l) * NT + t) * NT * kWChunkBytes;
        const uint8_t* abase = act_slot + (size_t)buf * NT * kAChunkStride;
        const uint32_t* my_flag = aflag + buf * NT + (lane < NT ? lane : 0);
        {
          // The model (203 MB for Panda) does not stay in L2 between calls, so every weight byte comes from HBM once
          // per launch; pull the NEXT hidden layer's slice of this CTA into L2 now, one whole layer ahead of its use,
          // so that the ring refills at L2 latency instead of DRAM latency.
          int n2 = n, l2 = l + 1;
          if (l2 == p.n_big) {
            l2 = 0;
            n2 = -1;
            if (g + 1 < total_steps) {
              const int in_rg2 = (g + 1) % steps_per_rg;
              n2 = 2 * (p.block_first - in_rg2 / 2) + (in_rg2 & 1);
            }
          }
          if (n2 >= 0) {
            const uint8_t* wnext =
                reinterpret_cast<const uint8_t*>(p.big_w) + (((size_t)n2 * p.n_big + l2) * NT + t) * NT * kWChunkBytes;
            for (int k = lane; k < NT; k += 32) bulk_prefetch_l2(wnext + (size_t)((t + k) % NT) * kWChunkBytes, kWChunkBytes);
          }
        }
        uint32_t ready = 0;  // bit c: chunk c has been published (warp-uniform)
        int issued_w = 0, issued_a = 0;
        bool gave_up = false;
        uint32_t spins = 0;
        long long t0 = 0;
        while (issued_a < NT) {
          // 1) which stages are free for the next weight chunks?  lane k looks at chunk issued_w + k
          int n_w = 0;
          {
            bool free_ = false;
            if (lane < kStages && issued_w + lane < NT) {
              const uint32_t pos = ring_pos + issued_w + lane;
              const uint32_t use = pos / kStages;
              free_ = use == 0 || mbar_test_wait(&sm.empty[pos % kStages], (use - 1) & 1);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, free_);
            n_w = __ffs(~m) - 1;  // consecutive free stages starting at chunk issued_w
          }
          // 2) poll the flags that are still outstanding
          if (!gave_up && ready != 0xffffffffu) {
            bool ok = false;
            if (lane < NT && !((ready >> lane) & 1u)) ok = (int32_t)(ld_relaxed(my_flag) - expected) >= 0;
            ready |= __ballot_sync(0xffffffffu, ok);
            if (__popc(ready) == NT) ready = 0xffffffffu;
          }
          // 3) activation chunks, in order, for stages whose weight copy is (being) issued
          int n_go = 0;
          while (n_go < kStages && issued_a + n_go < issued_w + n_w && ((ready >> ((t + issued_a + n_go) % NT)) & 1u))
            ++n_go;
          if (n_w > 0 || n_go > 0) {
            // One bulk copy keeps its issuing thread busy for ~0.4 us whatever its size, but copies issued by
            // different lanes run concurrently (measured: scripts/ubench/ingest2.cu) -- so every pending copy gets
            // its own lane and all of them go out in one instruction: lanes 0.. the weight chunks, lanes 8.. the
            // activation chunks.
            const bool do_w = lane < n_w;
            const bool do_a = lane >= 8 && lane < 8 + n_go;
            void* dst = nullptr;
            const void* src = nullptr;
            uint32_t bytes = 0;
            uint64_t* bar = nullptr;
            if (do_w) {
              const int iw = issued_w + lane;
              const int st = (ring_pos + iw) % kStages;
              mbar_arrive_expect_tx(&sm.full[st], C::kStageBytes);
              dst = sm.ring[st] + C::kAChunkBytes;
              src = wbase + (size_t)((t + iw) % NT) * kWChunkBytes;
              bytes = kWChunkBytes;
              bar = &sm.full[st];
            }
            if (do_a) {
              const int ia = issued_a + (lane - 8);
              const int st = (ring_pos + ia) % kStages;
              dst = sm.ring[st];
              src = abase + (size_t)((t + ia) % NT) * kAChunkStride;
              bytes = C::kAChunkBytes;
              bar = &sm.full[st];
            }
            __syncwarp();
            // The activation chunks were written by bulk stores that completed before their flag was released and are
            // read here by bulk copies from L2 (no L1 in the path); the copies are issued after (and control-dependent
            // on) the flag loads.  No proxy fence here: fence.proxy.async drains the bulk copies already in flight,
            // which serialised the whole ring (measured: 0.8 us per chunk with the fence).
            if (do_w || do_a) bulk_g2s(dst, src, bytes, bar);
            __syncwarp();
            if (lane == 0) {
              if (issued_w == 0 && n_w > 0) trace_ev(p, g * 4 + l, 0);
              if (issued_a == 0 && n_go > 0) trace_ev(p, g * 4 + l, 1);
              if (issued_a + n_go == NT) trace_ev(p, g * 4 + l, 2);
            }
            const bool first = issued_a == 0 && n_go > 0;
            issued_w += n_w;
            issued_a += n_go;
            spins = 0;
            if (l == 0 && first) prefetch_small(g + 1);
          } else {
            // nothing to do yet: back off a little; give up after about a second (see wait_flag)
            ++spins;
            if (spins == 64) t0 = clock64();
            if (spins > 64) {
              __nanosleep(20);
              if ((spins & 255u) == 0 && !gave_up) {
                int bail = 0;
                if (lane == 0) {
                  if (ld_relaxed(p.status + 1) == launch_id) bail = 1;
                  else if (clock64() - t0 > 2500000000LL) {
                    report_timeout(p.status, p.status_host, launch_id);
                    bail = 1;
                  }
                }
                if (__shfl_sync(0xffffffffu, bail, 0)) {
                  gave_up = true;
                  ready = 0xffffffffu;
                }
              }
            }
          }
        }
        ring_pos += NT;
        ++act_w[buf];
        ++xchg;
      }
    }
  } else if (warp == kStorerWarp) {
    // ===== storer: publishes this CTA's activation chunk to the team =====
    uint32_t act_w[2] = {0, 0};
    uint32_t xchg = 0;
    for (int g = 0; g < total_steps; ++g) {
      for (int l = 0; l < p.n_big; ++l) {
        const int buf = xchg & 1;
        bar_staged_sync();  // compute warps have written + proxy-fenced the staging buffer
        if (lane == 0) {
          uint8_t* dst = act_slot + ((size_t)buf * NT + t) * kAChunkStride;
          trace_ev(p, g * 4 + l, 3);
          bulk_s2g(dst, sm.staging, C::kAChunkBytes);
          bulk_commit();
          bulk_wait_all();
          trace_ev(p, g * 4 + l, 4);
          st_release(aflag + buf * NT + t, p.epoch + 1 + act_w[buf]);
          trace_ev(p, g * 4 + l, 5);
          mbar_arrive(&sm.staging_free);
        }
        __syncwarp();
        ++act_w[buf];
        ++xchg;
      }
    }
  } else {
    // ===== compute warps: 2 x 2 over the RT x 64 tile, warp tile (RT/2) rows x 32 features =====
    const int warp_m = warp >> 1;
    const int warp_n = warp & 1;
    const int gq = lane >> 2, tq = lane & 3;
    const int row_base = warp_m * (16 * MI);  // + mi * 16 + gq (+ 8)
    const int col_base = warp_n * 32;         // + ni * 8 + 2 * tq (+ 1)
    uint32_t ring_pos = 0;
    uint32_t part_w[2] = {0, 0};
    uint32_t pxchg = 0;
    uint32_t staged = 0;  // chunks handed to the storer so far

    // per-lane ldmatrix offsets inside a tile (k16 step added later)
    const int a_row_in = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int a_kc = lane >> 4;  // which 16-byte k chunk of the k16 step
    const int b_row_in = (lane & 7) + (lane >> 4) * 8;
    const int b_kc = (lane >> 3) & 1;

    int g = 0;
    for (int rg = slot; rg < p.n_rowgroups; rg += p.slots) {
      // ---- load the flow state and the condition of this row group ----
      for (int i = tid; i < RT * kPad; i += kComputeThreads) {
        const int r = i / kPad, j = i % kPad;
        const int row = rg * RT + r;
        float uv = 0.f, cv = 0.f;
        if (row < p.batch) {
          if (j < p.W) uv = p.in[(size_t)row * p.in_ld + j];
          if (j < p.cond_cols) cv = p.cond[(size_t)(row % p.cond_rows) * p.cond_ld + j];
        }
        sm.u[r][j] = uv;
        if (j < 8) sm.cnd[r][j] = cv;
      }
      bar_compute();

      for (int blk = p.block_first; blk >= p.block_last; --blk) {
        for (int sidx = 0; sidx < 2; ++sidx, ++g) {
          const int sb = g & 1;
          const float* sp = sm.small[sb];
          mbar_wait(&sm.small_full[sb], (g >> 1) & 1);
          // subnet1 reads the first half and transforms the second; subnet2 the other way round
          const int in_off = sidx == 0 ? 0 : p.s1;
          const int in_len = sidx == 0 ? p.s1 : p.s2;
          const int tg_off = sidx == 0 ? p.s1 : 0;
          const int tg_len = sidx == 0 ? p.s2 : p.s1;

          // v[mi][ni][e]: activation (after bias + LeakyReLU) of row row_base + mi*16 + gq + 8*(e>>1),
          //               feature col_base + ni*8 + 2*tq + (e&1)  -- the mma accumulator fragment layout
          float v[MI][4][4];

          // ---- first layer: fp32 SIMT from the replicated state ----
          {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
              const float2 b2 = *reinterpret_cast<const float2*>(sp + kSmallFirstB + col_base + ni * 8 + 2 * tq);
#pragma unroll
              for (int mi = 0; mi < MI; ++mi) {
                v[mi][ni][0] = b2.x;
                v[mi][ni][1] = b2.y;
                v[mi][ni][2] = b2.x;
                v[mi][ni][3] = b2.y;
              }
            }
            const int kin = in_len + p.dim_cond;
            for (int k = 0; k < kin; ++k) {
              float x[MI][2];
#pragma unroll
              for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int r = row_base + mi * 16 + gq + 8 * h;
                  x[mi][h] = k < in_len ? sm.u[r][in_off + k] : sm.cnd[r][k - in_len];
                }
              const float* wr = sp + kSmallFirstW + k * kFT + col_base + 2 * tq;
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) {
                const float2 w2 = *reinterpret_cast<const float2*>(wr + ni * 8);
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                  v[mi][ni][0] = fmaf(x[mi][0], w2.x, v[mi][ni][0]);
                  v[mi][ni][1] = fmaf(x[mi][0], w2.y, v[mi][ni][1]);
                  v[mi][ni][2] = fmaf(x[mi][1], w2.x, v[mi][ni][2]);
                  v[mi][ni][3] = fmaf(x[mi][1], w2.y, v[mi][ni][3]);
                }
              }
            }
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
              for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[mi][ni][e] = leaky(v[mi][ni][e]);
          }

          for (int l = 0; l <= p.n_big; ++l) {
            if (l > 0) {
              // ---- hidden layer l-1: bf16x3 tensor-core tiles over the NT k-chunks ----
#pragma unroll
              for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                  for (int e = 0; e < 4; ++e) v[mi][ni][e] = 0.f;
              if (tid == 0) trace_ev(p, g * 4 + l - 1, 6);
              for (int i = 0; i < NT; ++i) {
                const int s = ring_pos % kStages;
                mbar_wait(&sm.full[s], (ring_pos / kStages) & 1);
                if (tid == 0 && i == 0) trace_ev(p, g * 4 + l - 1, 7);
                const uint32_t a_hi = smem_u32(sm.ring[s]);
                const uint32_t a_lo = a_hi + C::kATileBytes;
                const uint32_t w_hi = a_hi + C::kAChunkBytes;
                const uint32_t w_lo = w_hi + kWTileBytes;
#pragma unroll
                for (int kk = 0; kk < kKC / 16; ++kk) {
                  uint32_t ah[MI][4], al[MI][4], bh[2][4], bl[2][4];
#pragma unroll
                  for (int mi = 0; mi < MI; ++mi) {
                    const int row = row_base + mi * 16 + a_row_in;
                    const uint32_t off = row * 128 + ((((kk * 2 + a_kc) ^ (row & 7)) & 7) << 4);
                    ldsm_x4(a_hi + off, ah[mi]);
                    ldsm_x4(a_lo + off, al[mi]);
                  }
#pragma unroll
                  for (int nj = 0; nj < 2; ++nj) {
                    const int row = col_base + nj * 16 + b_row_in;
                    const uint32_t off = row * 128 + ((((kk * 2 + b_kc) ^ (row & 7)) & 7) << 4);
                    ldsm_x4(w_hi + off, bh[nj]);
                    ldsm_x4(w_lo + off, bl[nj]);
                  }
#pragma unroll
                  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
                      const int nj = ni >> 1, o = (ni & 1) * 2;
                      if (p.precision == IKF_PRECISION_BF16X3) {
                        mma_bf16(v[mi][ni], al[mi], bh[nj][o], bh[nj][o + 1]);
                        mma_bf16(v[mi][ni], ah[mi], bl[nj][o], bl[nj][o + 1]);
                      }
                      mma_bf16(v[mi][ni], ah[mi], bh[nj][o], bh[nj][o + 1]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[s]);
                ++ring_pos;
              }
              if (tid == 0) trace_ev(p, g * 4 + l - 1, 8);
              const float* bb = sp + kSmallBigB + (l - 1) * kFT + col_base + 2 * tq;
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) {
                const float2 b2 = *reinterpret_cast<const float2*>(bb + ni * 8);
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                  v[mi][ni][0] = leaky(v[mi][ni][0] + b2.x);
                  v[mi][ni][1] = leaky(v[mi][ni][1] + b2.y);
                  v[mi][ni][2] = leaky(v[mi][ni][2] + b2.x);
                  v[mi][ni][3] = leaky(v[mi][ni][3] + b2.y);
                }
              }
            }

            if (l < p.n_big) {
              // ---- publish: split into bf16 head/tail, stage the swizzled chunk, hand it to the storer ----
              if (staged > 0) mbar_wait(&sm.staging_free, (staged - 1) & 1);
#pragma unroll
              for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    const float f0 = v[mi][ni][2 * h], f1 = v[mi][ni][2 * h + 1];
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(f0), h1 = __float2bfloat16_rn(f1);
                    const __nv_bfloat16 l0 = __float2bfloat16_rn(f0 - __bfloat162float(h0));
                    const __nv_bfloat16 l1 = __float2bfloat16_rn(f1 - __bfloat162float(h1));
                    const uint32_t off = tile_off_bytes(row_base + mi * 16 + gq + 8 * h, col_base + ni * 8 + 2 * tq);
                    *reinterpret_cast<uint32_t*>(sm.staging + off) = pack_bf16(h0, h1);
                    *reinterpret_cast<uint32_t*>(sm.staging + C::kATileBytes + off) = pack_bf16(l0, l1);
                  }
              fence_proxy_async_smem();
              bar_staged_arrive();
              if (tid == 0) trace_ev(p, g * 4 + l, 10);
              ++staged;
            }
          }

          // ---- last layer: this CTA's 64 features of every output, in fp32 straight from the activations ----
          if (tid == 0) trace_ev(p, g * 4 + 3, 11);
          const int pb = pxchg & 1;
          if (staged > 0) mbar_wait(&sm.staging_free, (staged - 1) & 1);  // the storer is done with the buffer
          float* wpart = reinterpret_cast<float*>(sm.staging);           // [2 warp_n][RT][16]
          {
#pragma unroll
            for (int o4 = 0; o4 < 4; ++o4) {
              float po[MI][2][4];
#pragma unroll
              for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                  for (int oo = 0; oo < 4; ++oo) po[mi][h][oo] = 0.f;
#pragma unroll
              for (int oo = 0; oo < 4; ++oo) {
                const float* wr = sp + kSmallLastW + (o4 * 4 + oo) * kFT + col_base + 2 * tq;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                  const float2 w2 = *reinterpret_cast<const float2*>(wr + ni * 8);
#pragma unroll
                  for (int mi = 0; mi < MI; ++mi) {
                    po[mi][0][oo] = fmaf(v[mi][ni][0], w2.x, po[mi][0][oo]);
                    po[mi][0][oo] = fmaf(v[mi][ni][1], w2.y, po[mi][0][oo]);
                    po[mi][1][oo] = fmaf(v[mi][ni][2], w2.x, po[mi][1][oo]);
                    po[mi][1][oo] = fmaf(v[mi][ni][3], w2.y, po[mi][1][oo]);
                  }
                }
              }
#pragma unroll
              for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                  for (int oo = 0; oo < 4; ++oo) {
                    float s = po[mi][h][oo];
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    po[mi][h][oo] = s;
                  }
                  if (tq == o4) {
                    const int r = row_base + mi * 16 + gq + 8 * h;
                    *reinterpret_cast<float4*>(&wpart[(warp_n * RT + r) * kPad + 4 * o4]) =
                        make_float4(po[mi][h][0], po[mi][h][1], po[mi][h][2], po[mi][h][3]);
                  }
                }
            }
          }
          bar_compute();
          for (int i = tid; i < RT * 4; i += kComputeThreads) {
            const int r = i >> 2, o4 = i & 3;
            const float4 x0 = *reinterpret_cast<const float4*>(&wpart[r * kPad + 4 * o4]);
            const float4 x1 = *reinterpret_cast<const float4*>(&wpart[(RT + r) * kPad + 4 * o4]);
            float* dst = part_slot + (((size_t)pb * NT + t) * kRTMax + r) * kPad + 4 * o4;
            __stcg(reinterpret_cast<float4*>(dst), make_float4(x0.x + x1.x, x0.y + x1.y, x0.z + x1.z, x0.w + x1.w));
          }
          bar_compute();
          const uint32_t pexp = p.epoch + 1 + part_w[pb];
          if (tid == 0) st_release(pflag + pb * NT + t, pexp);  // cumulative over the barrier: one fence per CTA
          if (warp == 0) {
            for (int c = lane; c < NT; c += 32) wait_flag(pflag + pb * NT + c, pexp, p.status, launch_id, p.status_host);
            __threadfence();  // acquire
          }
          bar_compute();
          if (tid == 0) trace_ev(p, g * 4 + 3, 12);
          for (int i = tid; i < RT * 4; i += kComputeThreads) {
            // fixed summation order over the team: every CTA obtains bitwise identical coefficients
            const int r = i >> 2, o4 = i & 3;
            float4 s = *reinterpret_cast<const float4*>(sp + kSmallLastB + 4 * o4);
            const float* src = part_slot + ((size_t)pb * NT * kRTMax + r) * kPad + 4 * o4;
            for (int c = 0; c < NT; ++c) {
              const float4 x = __ldcg(reinterpret_cast<const float4*>(src + (size_t)c * kRTMax * kPad));
              s.x += x.x;
              s.y += x.y;
              s.z += x.z;
              s.w += x.w;
            }
            *reinterpret_cast<float4*>(&sm.a[r][4 * o4]) = s;
          }
          // this subnet's small parameters are no longer needed
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.small_empty[sb]);
          ++part_w[pb];
          ++pxchg;
          bar_compute();
          // ---- affine coupling, reverse direction: y = (x - t) * exp(-clamp * 0.636 * atan(s)) ----
          for (int i = tid; i < RT * tg_len; i += kComputeThreads) {
            const int r = i / tg_len, j = i % tg_len;
            const float sc = p.clamp_scale * atanf(sm.a[r][j]);
            const float tr = sm.a[r][tg_len + j];
            sm.u[r][tg_off + j] = (sm.u[r][tg_off + j] - tr) * expf(-sc);
          }
          bar_compute();
          if (tid == 0) trace_ev(p, g * 4 + 3, 13);
        }
        // ---- PermuteRandom reverse: u = u[:, perm_inv] ----
        {
          constexpr int kPer = (RT * kPad + kComputeThreads - 1) / kComputeThreads;
          float tmp[kPer];
#pragma unroll
          for (int c = 0; c < kPer; ++c) {
            const int i = tid + c * kComputeThreads;
            tmp[c] = i < RT * p.W ? sm.u[i / p.W][p.perm_inv[blk * kPad + i % p.W]] : 0.f;
          }
          bar_compute();
#pragma unroll
          for (int c = 0; c < kPer; ++c) {
            const int i = tid + c * kComputeThreads;
            if (i < RT * p.W) sm.u[i / p.W][i % p.W] = tmp[c];
          }
          bar_compute();
        }
      }

      // ---- write this row group (team member 0 only; all replicas are identical) ----
      if (t == 0) {
        for (int i = tid; i < RT * p.out_cols; i += kComputeThreads) {
          const int r = i / p.out_cols, j = i % p.out_cols;
          const int row = rg * RT + r;
          if (row >= p.batch) continue;
          float o;
          if (p.finalize) {
            // FixedLinearTransform reverse (x - b) @ M_inv, slice, joint-limit clamp (ikflow_solver.py:98-102)
            o = 0.f;
            for (int k = 0; k < p.W; ++k) o = fmaf(sm.u[r][k] - p.flt_b[k], p.m_inv[k * kPad + j], o);
            // NaN / inf are reported BEFORE the clamp, and NaN survives it as in torch.clamp
            if (!isfinite(o)) report_nonfinite(p.status, p.status_host);
            if (p.clamp_out && j < p.ndof && o == o) o = fminf(fmaxf(o, p.lo[j]), p.hi[j]);
          } else {
            o = sm.u[r][j];
            if (!isfinite(o)) report_nonfinite(p.status, p.status_host);
          }
          p.out[(size_t)row * p.out_ld + j] = o;
        }
      }
      bar_compute();
    }
  }
}

}  // namespace ikf
