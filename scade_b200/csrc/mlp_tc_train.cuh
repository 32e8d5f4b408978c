// Tensor-core backward of the field network (autograd of model/run_nerf_helpers.py:223-247 as used by
// run_scade_scannet.py:985), included into namespace scade::tc of mlp_tc.cu.
//
//   forward (kStash)  : nerf_mlp_tc_pp_kernel<true> leaves every layer's fp16 activations, the ReLU sign masks and the alpha
//                       pre-activation in the stash (TrainLayout).
//   dgrad             : nerf_mlp_tc_dgrad_kernel -- the same SM-pair ping-pong machinery as the forward (weight ring fed by TMA,
//                       tcgen05.mma.cta_group::2, accumulators in TMEM) walking the layers backwards with B = W^T:
//                       dZ_v -> d_feature -> dH_{D-1} (+ d_alpha * w_alpha) -> mask -> dZ_{D-1} -> ... -> dZ_0.
//                       The chain never leaves the SM; every dZ tile is also bulk-stored to the stash for the weight gradients.
//   wgrad             : nerf_mlp_tc_wgrad_kernel -- dW_l = dZ_l^T X_l with the POINT axis as K: both operands are the stashed
//                       [128 points x 64 features] chunks read as MN-major SWIZZLE_128B tiles (no transposition anywhere);
//                       accumulators [256 x (256 + 64)] fp32 in TMEM per SM pair, one pass over a contiguous span of the
//                       (layer, tile) work line per cluster, then red.global.add into the fp32 gradient tensors.
//                       The encoding chunk's constant-1 column turns the bias gradients into one more MMA column.
//   heads             : alpha_linear / rgb_linear gradients (skinny) on CUDA cores from the stash.
//
// Gradients travel as fp16 scaled by a power of two chosen per call (absmax kernel) so that the largest gradient entering the
// chain sits in [32, 64): 2^10 of headroom before fp16 saturates (conversions use .satfinite, never inf) and 2^20 below it in
// fp16's normal range; what falls under 2^-30 of the maximum is flushed, which fp32 accumulation of the same sums would lose too.
// The weight gradients are un-scaled in fp32.  (bf16 gradients would need no scaling, but tcgen05 kind::f16 rejects mixed
// bf16 x fp16 operands -- "illegal instruction" on B200 -- and the stashed activations are the forward's fp16 operand tiles.)
#pragma once

// scale = 2^(6 - (e+1)) with e = exponent of max|d_out|  ->  max scaled |d_out| in [32, 64)
__device__ __forceinline__ float grad_scale(uint32_t maxbits, float* inv) {
  const int e = (int)(maxbits >> 23) - 127;
  if (maxbits == 0u || e < -110 || e > 110) { *inv = 1.0f; return 1.0f; }
  *inv = exp2f((float)(e - 5));
  return exp2f((float)(5 - e));
}

// max over points of (|d_rgb_raw|, |d_alpha|) with d_alpha = d_sigma * softplus'(alpha): the magnitudes that actually enter the
// GEMM chain.  (d_sigma itself is useless for this: compute_weights' 1e10 "last interval" (RS:514-515) produces d_sigma outliers
// many orders of magnitude above the rest, which softplus' then multiplies by ~1e-8.)
__global__ void absmax_kernel(const float4* __restrict__ x, const float* __restrict__ alpha_pre, int64_t n, uint32_t* __restrict__ out) {
  uint32_t m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    const float bx = alpha_pre[i] * 10.0f;
    const float da = v.w * (bx > 20.0f ? 1.0f : sigmoidf_(bx));
    const float mx = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(da)));
    if (mx < 3.0e38f) m = max(m, __float_as_uint(mx));          // ignore inf / nan rows (they saturate downstream)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(out, m);
}

struct BwdArgs {
  const uint8_t* packed;          // forward stream; the fp32 tail (w_alpha, w_rgb) sits at tail_off
  unsigned long long tail_off;
  uint8_t* ws;
  TrainLayout L;
  const float4* d_out;            // [P] (d_rgb_raw3, d_sigma)
  int64_t P;
  int64_t n_pairs;
};

// ---- dgrad chain -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PP_THREADS, 1) nerf_mlp_tc_dgrad_kernel(const __grid_constant__ BwdArgs a,
                                                                          const __grid_constant__ NetPlan plan,
                                                                          const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int64_t unit0 = cluster_id_x(), n_units = num_clusters_x();
  const int64_t n_steps = (a.n_pairs + 1) / 2;

  const uint32_t bar_full = sbase + OFF_BAR, bar_empty = bar_full + 8 * NUM_STAGES;
  const uint32_t bar_acc = bar_empty + 16 * NUM_STAGES, bar_aready = bar_acc + 16;     // (8 * NUM_STAGES bytes after bar_empty are unused)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (3 * NUM_STAGES + 4));

  if (threadIdx.x == 0) {
    for (int s = 0; s < NUM_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_acc + 8 * t, 1);
      mbar_init(bar_aready + 8 * t, PP_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) pp_weight_producer(plan, &tmap, sbase, bar_full, bar_empty, cta_rank, unit0, n_steps, n_units);
  } else if (warp == 1) {
    if (cta_rank == 0) pp_mma_issuer(plan, sbase, tmem_base, bar_full, bar_empty, bar_acc, bar_aready, unit0, n_steps, n_units);
  } else if (warp >= PP_EPI_WARP0) {
    // ================= prologue / epilogue warps: thread == one row x 128 columns =================
    const int ew = warp - PP_EPI_WARP0;
    const int tile = ew >> 3;
    const int quarter = warp & 3;
    const int half = (ew >> 2) & 1;
    const int row = quarter * 32 + lane;
    uint8_t* a_tile = smem + OFF_A + tile * 4 * CHUNK_BYTES;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + tile * W;
    const PackedTail* tail = reinterpret_cast<const PackedTail*>(a.packed + a.tail_off);
    const uint32_t my_acc = bar_acc + 8 * tile, my_aready = bar_aready + 8 * tile;
    const uint32_t aready_target = cta_rank != 0 ? map_to_cta(my_aready, 0) : my_aready;
    uint32_t acc_phase = 0;
    const int D = a.L.D;
    float inv_unused;
    const float gscale = grad_scale(*reinterpret_cast<const uint32_t*>(a.ws + a.L.gs), &inv_unused);
    auto signal_a_ready = [&]() {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank != 0) mbar_arrive_remote_relaxed(aready_target);
        else mbar_arrive(my_aready);
      }
    };

    for (int64_t step = unit0; step < n_steps; step += n_units) {
      const int64_t pair = 2 * step + cta_rank;
      const int64_t p_raw = pair * (TILES * TILE_M) + tile * TILE_M + row;
      const bool live = p_raw < a.P;
      const int64_t tile_g = 2 * pair + tile;
      const size_t grow = (size_t)(tile_g * TILE_M + row);

      // ---- prologue: heads backward.  dZ_v = (d_rgb W_rgb) * [z_v > 0]  (H:239-241); d_alpha = d_sigma softplus'(alpha)  (H:242)
      float d_alpha;
      {
        float4 g = live ? __ldg(a.d_out + p_raw) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float al = reinterpret_cast<const float*>(a.ws + a.L.alpha)[grow];
        const uint4 mv = *reinterpret_cast<const uint4*>(a.ws + a.L.maskv + grow * 16);
        const float bx = al * 10.0f;
        d_alpha = g.w * (bx > 20.0f ? 1.0f : sigmoidf_(bx)) * gscale;
        const float dr = g.x * gscale, dg = g.y * gscale, db = g.z * gscale;
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        uint8_t* chunk = a_tile + 2 * half * CHUNK_BYTES;       // views-bwd reads activation chunks 0 and 2 (a_step = 2)
#pragma unroll
        for (int w2 = 0; w2 < 2; ++w2) {
          const uint32_t word = half == 0 ? (w2 == 0 ? mv.x : mv.y) : (w2 == 0 ? mv.z : mv.w);
          const float4* wr = tail->w_rgb + 64 * half + 32 * w2;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            uint32_t qq[4];
#pragma unroll
            for (int ss = 0; ss < 2; ++ss) {
              const int s = 2 * pc + ss;
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 w = __ldg(wr + 4 * s + e);
                v[e] = fmaf(dr, w.x, fmaf(dg, w.y, db * w.z));
              }
              uint32_t m01, m23;
              inactive_masks(word, s, &m01, &m23);
              qq[2 * ss] = pack_sat_f16x2(v[0], v[1]) & ~m01;
              qq[2 * ss + 1] = pack_sat_f16x2(v[2], v[3]) & ~m23;
            }
            *reinterpret_cast<uint4*>(chunk + sw128_offset(row, w2 * 4 + pc)) = make_uint4(qq[0], qq[1], qq[2], qq[3]);
          }
        }
      }
      signal_a_ready();
      if (lane == 0) {
        bulk_s2g(a.ws + a.L.dzv + (size_t)(tile_g * 2 + half) * CHUNK_BYTES + quarter * 4096,
                 smem_u32(a_tile) + 2 * half * CHUNK_BYTES + quarter * 4096, 4096);
        bulk_commit();
      }

      // ---- layers, backwards: j = 0 views^T -> d_feature; j = 1 feature^T (+ alpha) -> dZ_{D-1}; j >= 2 pts_{D+1-j}^T -> dZ_{D-j}
      // The body is instantiated twice (with / without the alpha_linear term) so that the 8 common layers carry no predicated
      // FFMA / LDG; addresses are formed from per-thread constants hoisted out of the loops.
      const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
      const uint32_t rx16 = (uint32_t)(row & 7) << 4;
      uint8_t* const my_chunks = a_tile + 2 * half * CHUNK_BYTES + row_off;          // this thread's row in chunks 2h, 2h+1
      const float4* const wa4 = reinterpret_cast<const float4*>(tail->w_alpha) + half * 32;
      auto layer_epilogue = [&](auto alpha_tag, const uint4 mk) {
        constexpr bool kAlpha = decltype(alpha_tag)::value;
        uint32_t rbuf[2][32];
        const uint32_t t_col = t_lane + half * 128;
        tmem_ld32(t_col, rbuf[0]);
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t* r = rbuf[c4 & 1];
          tmem_ld_wait();
          if (c4 + 1 < 4) tmem_ld32(t_col + (c4 + 1) * 32, rbuf[(c4 + 1) & 1]);
          const uint32_t word = c4 == 0 ? mk.x : (c4 == 1 ? mk.y : (c4 == 2 ? mk.z : mk.w));
          uint8_t* const dst = my_chunks + (c4 >> 1) * CHUNK_BYTES;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            uint32_t qq[4];
#pragma unroll
            for (int ss = 0; ss < 2; ++ss) {
              const int s = 2 * pc + ss;
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = __uint_as_float(r[4 * s + e]);
              if (kAlpha) {                                                           // alpha_linear^T  (H:233)
                const float4 w = __ldg(wa4 + c4 * 8 + s);
                v[0] = fmaf(d_alpha, w.x, v[0]); v[1] = fmaf(d_alpha, w.y, v[1]);
                v[2] = fmaf(d_alpha, w.z, v[2]); v[3] = fmaf(d_alpha, w.w, v[3]);
              }
              uint32_t m01, m23;
              inactive_masks(word, s, &m01, &m23);                                    // j == 0: word = 0 -> nothing masked
              qq[2 * ss] = pack_sat_f16x2(v[0], v[1]) & ~m01;
              qq[2 * ss + 1] = pack_sat_f16x2(v[2], v[3]) & ~m23;
            }
            *reinterpret_cast<uint4*>(dst + ((uint32_t)(((c4 & 1) * 4 + pc) << 4) ^ rx16)) = make_uint4(qq[0], qq[1], qq[2], qq[3]);
          }
        }
      };
      for (int j = 0; j < plan.n_layers; ++j) {
        uint4 mk = make_uint4(0u, 0u, 0u, 0u);
        if (j >= 1) mk = *reinterpret_cast<const uint4*>(a.ws + a.L.maskh[D - j] + ((size_t)(tile_g * 2 + half) * TILE_M + row) * 16);
        mbar_wait(my_acc, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (j == 1) layer_epilogue(std::true_type{}, mk);
        else layer_epilogue(std::false_type{}, mk);
        if (j + 1 < plan.n_layers) {
          signal_a_ready();
        } else {
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
        }
        if (lane == 0) {
          uint8_t* dst = a.ws + (j == 0 ? a.L.dzf : a.L.dz[D - j]) + (size_t)(tile_g * 4 + 2 * half) * CHUNK_BYTES + quarter * 4096;
          const uint32_t src = smem_u32(a_tile) + 2 * half * CHUNK_BYTES + quarter * 4096;
          bulk_s2g(dst, src, 4096);
          bulk_s2g(dst + CHUNK_BYTES, src + CHUNK_BYTES, 4096);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait_all0();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---- wgrad -------------------------------------------------------------------------------------------------------------
constexpr int WG_STAGES = 5;
constexpr int WG_HALF = 8192;                       // 64 points x 128 B of one chunk
constexpr int WG_STAGE_BYTES = 5 * WG_HALF;         // dZ (2 blocks of 64 features) | X (2 blocks) | encoding chunk
constexpr int WG_THREADS = 256;
constexpr int WG_OFF_BAR = WG_STAGES * WG_STAGE_BYTES;
constexpr int WG_SMEM_BYTES = WG_OFF_BAR + 256 + 1024;
constexpr int WG_MAX_LAYERS = 10;                   // D pts layers + feature + views
constexpr int WG_EMB_COL = 256;                     // accumulator columns [256, 384): dZ^T x [encoding chunk | encoding chunk]

struct WgLayer {
  unsigned long long dz_off, x_off;
  float* gW;
  float* gb;
  int dz_chunks, has_x, ld, x_col0, emb_lo, emb_hi, emb_dst, n_out, cost, pad;
};
struct WgArgs {
  uint8_t* ws;
  unsigned long long emb_off, gs_off;
  long long T;
  int n_layers, pad;
  WgLayer layers[WG_MAX_LAYERS];
};

// tiles [t0, t1) of layer l that belong to cluster c: the work line is the concatenation of the layers, layer l being T tiles of
// `cost` units each; cluster c owns the tiles whose start lies in [total*c/nc, total*(c+1)/nc).
__device__ __forceinline__ void wg_piece(const WgArgs& a, int l, long long c, long long nc, long long* t0, long long* t1) {
  long long total = 0, off = 0;
  for (int i = 0; i < a.n_layers; ++i) {
    if (i == l) off = total;
    total += (long long)a.layers[i].cost * a.T;
  }
  const long long lo = total * c / nc, hi = total * (c + 1) / nc;
  const long long cost = a.layers[l].cost;
  auto first_at_or_after = [&](long long x) {
    long long t = x <= off ? 0 : (x - off + cost - 1) / cost;
    return t > a.T ? a.T : t;
  };
  *t0 = first_at_or_after(lo);
  *t1 = first_at_or_after(hi);
}

__global__ void __launch_bounds__(WG_THREADS, 1) nerf_mlp_tc_wgrad_kernel(const __grid_constant__ WgArgs a,
                                                                          const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const long long cid = cluster_id_x(), ncl = num_clusters_x();

  const uint32_t bar_full = sbase + WG_OFF_BAR, bar_empty = bar_full + 8 * WG_STAGES;
  const uint32_t bar_acc = bar_empty + 8 * WG_STAGES, bar_accfree = bar_acc + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + WG_OFF_BAR + 8 * (2 * WG_STAGES + 2));

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_accfree, 8);                     // 4 local + 4 remote drain warps
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer (both CTAs): its M half of dZ, its N half of X, the encoding chunk; 64 points per stage ----
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int l = 0; l < a.n_layers; ++l) {
        long long t0, t1;
        wg_piece(a, l, cid, ncl, &t0, &t1);
        const WgLayer& Ld = a.layers[l];
        const uint32_t bytes = (uint32_t)(3 + 2 * Ld.has_x) * WG_HALF;
        const int ach0 = Ld.n_out == 256 ? 2 * (int)cta_rank : 0;
        for (long long t = t0; t < t1; ++t) {
          for (int kh = 0; kh < 2; ++kh) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (cta_rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * bytes);
            const uint32_t base = sbase + stage * WG_STAGE_BYTES;
            const uint32_t bar = bar_full + 8 * stage;
#pragma unroll
            for (int i = 0; i < 2; ++i)
              tma_load_2d_pair(base + i * WG_HALF, &tmap, 0,
                               (int32_t)((Ld.dz_off + (unsigned long long)(t * Ld.dz_chunks + ach0 + i) * CHUNK_BYTES + kh * WG_HALF) >> 7), bar);
            if (Ld.has_x) {
#pragma unroll
              for (int i = 0; i < 2; ++i)
                tma_load_2d_pair(base + (2 + i) * WG_HALF, &tmap, 0,
                                 (int32_t)((Ld.x_off + (unsigned long long)(t * 4 + 2 * cta_rank + i) * CHUNK_BYTES + kh * WG_HALF) >> 7), bar);
            }
            tma_load_2d_pair(base + 4 * WG_HALF, &tmap, 0,
                             (int32_t)((a.emb_off + (unsigned long long)t * CHUNK_BYTES + kh * WG_HALF) >> 7), bar);
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA): K = points.  acc[0:256) += dZ^T X ; acc[256:384) += dZ^T [enc | enc] ----
    if (cta_rank == 0) {
      uint32_t stage = 0, phase = 0, free_phase = 0;
      bool first_piece = true;
      constexpr uint32_t idesc_x = make_idesc_mn(2 * TILE_M, 256), idesc_e = make_idesc_mn(2 * TILE_M, 128);
      for (int l = 0; l < a.n_layers; ++l) {
        long long t0, t1;
        wg_piece(a, l, cid, ncl, &t0, &t1);
        if (t0 >= t1) continue;
        const int has_x = a.layers[l].has_x;
        if (!first_piece) {                          // the previous piece's accumulators have been drained in both CTAs
          mbar_wait_cluster(bar_accfree, free_phase);
          free_phase ^= 1;
          tc_fence_after();
        }
        first_piece = false;
        for (long long t = t0; t < t1; ++t) {
          for (int kh = 0; kh < 2; ++kh) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t fresh = (t == t0 && kh == 0) ? 1u : 0u;
            if (elect_one()) {
              const uint32_t base = sbase + stage * WG_STAGE_BYTES;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t accum = (fresh && ks == 0) ? 0u : 1u;
                const uint64_t a_desc = make_smem_desc_mn(base + ks * 2048, WG_HALF);
                if (has_x) mma_f16_ss_pair(tmem_base, a_desc, make_smem_desc_mn(base + 2 * WG_HALF + ks * 2048, WG_HALF), idesc_x, accum);
                mma_f16_ss_pair(tmem_base + WG_EMB_COL, a_desc, make_smem_desc_mn(base + 4 * WG_HALF + ks * 2048, WG_HALF), idesc_e, accum);
              }
              mma_commit_pair(bar_empty + 8 * stage, (uint16_t)3);
            }
            __syncwarp();
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) mma_commit_pair(bar_acc, (uint16_t)3);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ---- drain warps (both CTAs): thread = one output feature (TMEM lane); un-scale and red.add into the fp32 gradients ----
    const int quarter = warp & 3;
    const int lrow = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t free_target = cta_rank != 0 ? map_to_cta(bar_accfree, 0) : bar_accfree;
    float inv;
    grad_scale(*reinterpret_cast<const uint32_t*>(a.ws + a.gs_off), &inv);
    uint32_t acc_phase = 0;
    for (int l = 0; l < a.n_layers; ++l) {
      long long t0, t1;
      wg_piece(a, l, cid, ncl, &t0, &t1);
      if (t0 >= t1) continue;
      const WgLayer& Ld = a.layers[l];
      mbar_wait(bar_acc, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      const bool active = Ld.n_out == 256 || cta_rank == 0;
      const int f = Ld.n_out == 256 ? 128 * (int)cta_rank + lrow : lrow;
      if (active) {
        float* grow = Ld.gW + (size_t)f * Ld.ld;
        if (Ld.has_x) {
#pragma unroll 1
          for (int c8 = 0; c8 < 8; ++c8) {
            uint32_t r[32];
            tmem_ld32(t_lane + c8 * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(grow + Ld.x_col0 + c8 * 32 + j, __uint_as_float(r[j]) * inv);
          }
        }
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t r[32];
          tmem_ld32(t_lane + WG_EMB_COL + c2 * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int e = c2 * 32 + j;
            const float v = __uint_as_float(r[j]) * inv;
            if (e >= Ld.emb_lo && e < Ld.emb_hi) atomicAdd(grow + Ld.emb_dst + e - Ld.emb_lo, v);
            if (e == ONES_COL) atomicAdd(Ld.gb + f, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank != 0) mbar_arrive_remote(free_target);
        else mbar_arrive(bar_accfree);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---- skinny heads: alpha_linear (W -> 1) and rgb_linear (W/2 -> 3) gradients -----------------------------------------
// HBM-bound (768 B per point re-read from the stash).  Block = 8 warps; warp g walks rows g, g+8, ... of a tile, lane j owns the
// 16-byte piece j of the row (8 columns of h_last; lanes < 16 also 8 columns of h_v) -- 512 B coalesced per warp load, the
// swizzle position is fixed per thread because row & 7 == g.  Partials live in registers across tiles; one shared-memory
// reduction over the 8 row groups and one atomicAdd per column at the end.  Not scaled: d_out is used in fp32.
__global__ void __launch_bounds__(256) head_wgrad_tc_kernel(const uint8_t* __restrict__ ws, const __grid_constant__ TrainLayout L,
                                                            const float4* __restrict__ d_out, int64_t P,
                                                            float* __restrict__ g_alpha_w, float* __restrict__ g_alpha_b,
                                                            float* __restrict__ g_rgb_w, float* __restrict__ g_rgb_b) {
  __shared__ float4 sd[TILE_M];
  __shared__ float red[8][32][33];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int chunk = lane >> 3, pj = lane & 7;
  const uint32_t piece_off = (uint32_t)chunk * CHUNK_BYTES + ((uint32_t)(pj ^ g) << 4);      // + row * 128
  float aa[8], ar[8], ag[8], ab[8], bias[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) aa[i] = ar[i] = ag[i] = ab[i] = 0.f;
  for (long long tile = blockIdx.x; tile < L.T; tile += gridDim.x) {
    if (threadIdx.x < TILE_M) {
      const int64_t p = tile * TILE_M + threadIdx.x;
      float4 d = p < P ? __ldg(d_out + p) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float bx = reinterpret_cast<const float*>(ws + L.alpha)[p] * 10.0f;
      d.w *= bx > 20.0f ? 1.0f : sigmoidf_(bx);
      sd[threadIdx.x] = d;
    }
    __syncthreads();
    const uint8_t* hbase = ws + L.h[L.D - 1] + (size_t)tile * 4 * CHUNK_BYTES + piece_off;
    const uint8_t* vbase = ws + L.hv + (size_t)tile * 2 * CHUNK_BYTES + piece_off;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int r = g + 8 * i;
      const float4 d = sd[r];
      const uint4 hq = __ldg(reinterpret_cast<const uint4*>(hbase + r * 128));
      const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[k]));
        aa[2 * k] = fmaf(d.w, f.x, aa[2 * k]);
        aa[2 * k + 1] = fmaf(d.w, f.y, aa[2 * k + 1]);
      }
      if (lane < 16) {
        const uint4 vq = __ldg(reinterpret_cast<const uint4*>(vbase + r * 128));
        const uint32_t vw[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&vw[k]));
          ar[2 * k] = fmaf(d.x, f.x, ar[2 * k]); ar[2 * k + 1] = fmaf(d.x, f.y, ar[2 * k + 1]);
          ag[2 * k] = fmaf(d.y, f.x, ag[2 * k]); ag[2 * k + 1] = fmaf(d.y, f.y, ag[2 * k + 1]);
          ab[2 * k] = fmaf(d.z, f.x, ab[2 * k]); ab[2 * k + 1] = fmaf(d.z, f.y, ab[2 * k + 1]);
        }
      }
      if (lane == 0) { bias[0] += d.x; bias[1] += d.y; bias[2] += d.z; bias[3] += d.w; }
    }
    __syncthreads();
  }
  // reduce over the 8 row groups: red[g][lane][slot]; slots 0..7 alpha, 8..15 r, 16..23 g, 24..31 b, 32 unused
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[g][lane][i] = aa[i];
    red[g][lane][8 + i] = ar[i];
    red[g][lane][16 + i] = ag[i];
    red[g][lane][24 + i] = ab[i];
  }
  __syncthreads();
  {
    // thread t -> alpha column t (lane j = t / 8 owns columns 8j..8j+7)
    const int t = threadIdx.x, j = t >> 3, i = t & 7;
    float s = 0.f;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) s += red[gg][j][i];
    atomicAdd(g_alpha_w + t, s);
    if (t < TILE_M) {
      float sr = 0.f, sg = 0.f, sb = 0.f;
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) { sr += red[gg][j][8 + i]; sg += red[gg][j][16 + i]; sb += red[gg][j][24 + i]; }
      atomicAdd(g_rgb_w + t, sr);
      atomicAdd(g_rgb_w + TILE_M + t, sg);
      atomicAdd(g_rgb_w + 2 * TILE_M + t, sb);
    }
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) red[g][0][k] = bias[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int gg = 0; gg < 8; ++gg) s += red[gg][0][threadIdx.x];
    atomicAdd(threadIdx.x < 3 ? g_rgb_b + threadIdx.x : g_alpha_b, s);
  }
}
