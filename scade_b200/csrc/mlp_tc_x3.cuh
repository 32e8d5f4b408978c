// SCADE_PREC_TC_F16X3: the "tight" tensor-core mode of the field network -- tcgen05 throughput at an fp32-level tolerance.
// Included into namespace scade::tc of mlp_tc.cu.
//
//   reference arithmetic: nn.Linear = true fp32 GEMMs (model/run_nerf_helpers.py:131-139, 227; TF32 is off in torch).
//   tcgen05 has no fp32 MMA kind, so every operand of the wide layers is carried as an fp16 (hi, lo) PAIR,
//       x = hi + lo,  hi = fp16(x),  lo = fp16(x - hi)                         (|x - hi - lo| <= 2^-22 |x|)
//   and every product is three tensor-core passes into the SAME fp32 TMEM accumulator:
//       A W^T  ~=  A_hi W_hi^T  +  A_lo W_hi^T  +  A_hi W_lo^T                  (the dropped A_lo W_lo^T term is ~2^-22 relative)
//   SURVEY App. D measured this class of split at 83-93 dB end to end against fp32 (single-pass fp16: 55-62 dB).
//
// Kernel shape (nerf_mlp_tc_x3_kernel): the SM-pair machinery of the fast kernel (cluster of 2 CTAs, tcgen05.mma.cta_group::2 with
// M = 256, weight stages streamed by 2D TMA through the 4-slot ring, accumulators in TMEM) with ONE 256-point super-tile in flight:
//   * shared memory holds the tile's A operand twice -- activation chunks [0..3] = hi halves, [4..7] = lo halves, encoding
//     chunk 0 = hi, 1 = lo -- in exactly the bytes the fast kernel spends on two tiles;
//   * the packed weight stream interleaves (W_hi stage, W_lo stage) per K stage; a W_hi stage feeds 8 MMAs (A_hi, A_lo), a W_lo
//     stage 4 (A_hi): 48 MMAs per hidden layer instead of 16;
//   * the two 256-column halves of TMEM alternate between layers, and the epilogue is CHUNK-STAGGERED: each of the 16 epilogue
//     warps owns 32 rows x 16 columns of every 64-column K chunk, so chunk j of the next layer's operand is complete after
//     (j+1)/4 of the epilogue and the issuer starts the next layer's MMAs on it (into the other accumulator half) while
//     chunks j+1.. are still being converted -- the epilogue hides under the 3x longer MMA phase without a second tile;
//   * bias: fp32 row through shared memory (layers without the encoding chunk) or hi/lo halves riding on the encoding chunk's
//     two 1.0 columns (W_hi stages only; the W_lo stages carry zeros there); alpha_linear / rgb_linear are fp32 dot products of
//     the un-rounded activations; the positional encoding uses sinf/cosf (<= 2 ulp), not the MUFU approximations.
// Forward / eval only: training at an fp32 tolerance goes through the fp32 FFMA path (functional.py).
#pragma once

constexpr int X3_OFF_SBIAS = OFF_BIAS;              // 256 floats: fp32 bias row of the current layer
constexpr int X3_OFF_SALPHA = OFF_BIAS + W * 4;     // 256 floats: alpha_linear weights (last hidden layer)
constexpr int X3_BAR_ACC = 64, X3_BAR_CHUNK = 72, X3_BAR_EMB = 104, X3_TMEM_SLOT = 128;      // byte offsets inside OFF_BAR
constexpr int X3_ARRIVALS = 2 * PP_EPI_WARPS;       // 16 local + 16 remote epilogue warps per operand barrier
constexpr int X3_LO = 4;                            // activation chunk index of the lo halves; encoding chunk 1 = lo

// MMA issuer of the leader CTA.  Stage order of a layer in the packed stream: for every ring stage i (kpack K chunks of this
// CTA's N half): W_hi stage, W_lo stage.
__device__ __forceinline__ void x3_mma_issuer(const NetPlan& plan, uint32_t sbase, uint32_t tmem_base, uint32_t bar_full,
                                              uint32_t bar_empty, uint32_t bar_acc, uint32_t bar_chunk, uint32_t bar_emb,
                                              int64_t unit0, int64_t n_steps, int64_t n_units) {
  constexpr uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
  uint32_t slot = 0, ph = 0, emb_ph = 0, acc_par = 0;
  uint32_t ch_ph = 0;                                          // bit c = phase of bar_chunk[c]
  for (int64_t step = unit0; step < n_steps; step += n_units) {
    for (int l = 0; l < plan.n_layers; ++l) {
      const int n_k = plan.layers[l].n_k, kpack = plan.layers[l].kpack;
      const int n_kst = (n_k + kpack - 1) / kpack;
      const int has_emb = plan.layers[l].a_src[0] == SRC_EMB;
      const int emb_ks0 = plan.layers[l].emb_ks0;
      const uint32_t idesc = make_idesc(2 * TILE_M, plan.layers[l].n_out);
      const uint32_t d_addr = tmem_base + acc_par * W;
      for (int i = 0; i < n_kst; ++i) {
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {                 // 0: W_hi stage (A_hi and A_lo), 1: W_lo stage (A_hi)
          mbar_wait(bar_full + 8 * slot, ph);
          tc_fence_after();
          for (int j = 0; j < kpack; ++j) {
            const int kc = i * kpack + j;
            if (kc >= n_k) break;
            const bool emb = has_emb && kc == 0;
            const int c = kc - has_emb;                        // activation chunk (forward layers: a_step == 1)
            if (pass == 0) {                                   // first touch of this K chunk: its operand must be complete
              if (emb) {
                if (l == 0) { mbar_wait(bar_emb, emb_ph); emb_ph ^= 1; }
              } else {
                mbar_wait(bar_chunk + 8 * c, (ch_ph >> c) & 1u);
                ch_ph ^= 1u << c;
              }
              tc_fence_after();
            }
            if (elect_one()) {
              const uint32_t a_hi = emb ? sbase + OFF_EMB : sbase + OFF_A + c * CHUNK_BYTES;
              const uint32_t a_lo = emb ? sbase + OFF_EMB + CHUNK_BYTES : sbase + OFF_A + (X3_LO + c) * CHUNK_BYTES;
              const uint64_t ah_desc = desc_hi | (uint64_t)((a_hi & 0x3FFFF) >> 4);
              const uint64_t al_desc = desc_hi | (uint64_t)((a_lo & 0x3FFFF) >> 4);
              const uint64_t b_desc = desc_hi | (uint64_t)(((sbase + OFF_STAGE + slot * STAGE_BYTES + j * (STAGE_BYTES / 2)) & 0x3FFFF) >> 4);
              const int ks0 = emb ? emb_ks0 : 0;
#pragma unroll
              for (int ks = 0; ks < KCHUNK / 16; ++ks) {
                if (ks < ks0) continue;
                if (pass == 0) {
                  mma_f16_ss_pair(d_addr, ah_desc + 2 * ks, b_desc + 2 * ks, idesc, (kc != 0 || ks != ks0) ? 1u : 0u);   // layer's first MMA overwrites
                  mma_f16_ss_pair(d_addr, al_desc + 2 * ks, b_desc + 2 * ks, idesc, 1u);
                } else {
                  mma_f16_ss_pair(d_addr, ah_desc + 2 * ks, b_desc + 2 * ks, idesc, 1u);
                }
              }
            }
            __syncwarp();
          }
          if (elect_one()) mma_commit_pair(bar_empty + 8 * slot, (uint16_t)3);
          __syncwarp();
          if (++slot == NUM_STAGES) { slot = 0; ph ^= 1; }
        }
      }
      if (elect_one()) mma_commit_pair(bar_acc, (uint16_t)3);
      __syncwarp();
      acc_par ^= 1;
    }
  }
}

// fp32 pair -> (hi, lo) fp16 pairs
__device__ __forceinline__ void split_pack_f16x2(float a, float b, uint32_t* hi, uint32_t* lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  *hi = *reinterpret_cast<const uint32_t*>(&h);
  *lo = *reinterpret_cast<const uint32_t*>(&l);
}

// One thread's share of a hidden / feature layer: row `row`, columns 64 j + 16 g + [0, 16) of every chunk j.
template <bool kBias, bool kRelu, bool kAlpha>
__device__ __forceinline__ float x3_hidden_epilogue(uint32_t t_acc, int g, uint8_t* a_row, uint32_t rx, const float* sbias,
                                                    const float* salpha, uint32_t chunk_target, bool remote, int lane) {
  uint32_t rb[2][16];
  float al0 = 0.f, al1 = 0.f;
  tmem_ld16(t_acc + 16 * g, rb[0]);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t* r = rb[j & 1];
    tmem_ld_wait();
    if (j + 1 < 4) tmem_ld16(t_acc + 64 * (j + 1) + 16 * g, rb[(j + 1) & 1]);
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    if (kBias) {
      const float4* p4 = reinterpret_cast<const float4*>(sbias + 64 * j + 16 * g);
      const float4 b0 = p4[0], b1 = p4[1], b2 = p4[2], b3 = p4[3];
      const float b[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += b[i];
    }
    if (kRelu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (kAlpha) {                                              // alpha_linear on the un-rounded activations (H:233)
      const float4* p4 = reinterpret_cast<const float4*>(salpha + 64 * j + 16 * g);
      const float4 w0 = p4[0], w1 = p4[1], w2 = p4[2], w3 = p4[3];
      const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
      for (int i = 0; i < 16; i += 2) { al0 = fmaf(v[i], w[i], al0); al1 = fmaf(v[i + 1], w[i + 1], al1); }
    }
    uint32_t hq[8], lq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_pack_f16x2(v[2 * i], v[2 * i + 1], &hq[i], &lq[i]);
    uint8_t* hi_chunk = a_row + j * CHUNK_BYTES;
    uint8_t* lo_chunk = a_row + (X3_LO + j) * CHUNK_BYTES;
    const uint32_t o0 = ((uint32_t)(2 * g) << 4) ^ rx, o1 = ((uint32_t)(2 * g + 1) << 4) ^ rx;
    *reinterpret_cast<uint4*>(hi_chunk + o0) = make_uint4(hq[0], hq[1], hq[2], hq[3]);
    *reinterpret_cast<uint4*>(hi_chunk + o1) = make_uint4(hq[4], hq[5], hq[6], hq[7]);
    *reinterpret_cast<uint4*>(lo_chunk + o0) = make_uint4(lq[0], lq[1], lq[2], lq[3]);
    *reinterpret_cast<uint4*>(lo_chunk + o1) = make_uint4(lq[4], lq[5], lq[6], lq[7]);
    // chunk j of the next layer's operand: this warp's 32 rows x 16 columns are in place
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (remote) mbar_arrive_remote_relaxed(chunk_target + 8 * j);
      else mbar_arrive(chunk_target + 8 * j);
    }
  }
  return al0 + al1;
}

__global__ void __launch_bounds__(PP_THREADS, 1) nerf_mlp_tc_x3_kernel(const __grid_constant__ FwdArgs a,
                                                                       const __grid_constant__ NetPlan plan,
                                                                       const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem_x3[];
  uint8_t* smem = smem_x3;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();                  // SWIZZLE_128B operand tiles need 1024-byte alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int64_t unit0 = cluster_id_x(), n_units = num_clusters_x();
  const int64_t n_steps = (a.P + 2 * TILE_M - 1) / (2 * TILE_M);          // 256 points per cluster step

  const uint32_t bar_full = sbase + OFF_BAR, bar_empty = bar_full + 8 * NUM_STAGES;
  const uint32_t bar_acc = sbase + OFF_BAR + X3_BAR_ACC, bar_chunk = sbase + OFF_BAR + X3_BAR_CHUNK, bar_emb = sbase + OFF_BAR + X3_BAR_EMB;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + X3_TMEM_SLOT);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NUM_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    for (int c = 0; c < 4; ++c) mbar_init(bar_chunk + 8 * c, X3_ARRIVALS);
    mbar_init(bar_emb, X3_ARRIVALS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PP_EPI_WARP0) {
    regs_control();
    if (warp == 0) {
      if (lane == 0) pp_weight_producer(plan, &tmap, sbase, bar_full, bar_empty, cta_rank, unit0, n_steps, n_units);
    } else if (warp == 1) {
      if (cta_rank == 0) x3_mma_issuer(plan, sbase, tmem_base, bar_full, bar_empty, bar_acc, bar_chunk, bar_emb, unit0, n_steps, n_units);
    }
  } else {
    // ================= prologue / epilogue warps: thread == one row x (16 columns of every K chunk) =================
    regs_epilogue();
    const int ew = warp - PP_EPI_WARP0;
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int g = ew >> 2;                              // 16-column group inside every 64-column chunk
    const int row = quarter * 32 + lane;
    const int tid = ew * 32 + lane;                     // 0..511
    const PackedTail* tail = reinterpret_cast<const PackedTail*>(a.packed + (size_t)plan.stages_per_pass * STAGE_BYTES);
    const bool remote = cta_rank != 0;
    const uint32_t chunk_target = remote ? map_to_cta(bar_chunk, 0) : bar_chunk;
    const uint32_t emb_target = remote ? map_to_cta(bar_emb, 0) : bar_emb;
    float* sbias = reinterpret_cast<float*>(smem + X3_OFF_SBIAS);
    float* salpha = reinterpret_cast<float*>(smem + X3_OFF_SALPHA);
    const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128), rx = (uint32_t)((row & 7) << 4);
    uint8_t* a_row = smem + OFF_A + row_off;
    uint8_t* emb_row = smem + OFF_EMB + row_off;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t acc_phase = 0, acc_par = 0;
    const int all_bar = 9, quarter_bar = 1 + quarter;   // named barriers: the 512 epilogue threads / the 4 warps that share a row quarter

    auto point_of = [&](int64_t step) { return (2 * step + cta_rank) * (int64_t)TILE_M + row; };

    // ---- this thread's 16 encoding-chunk columns [16 g, 16 g + 16) of step `step`, as 8 (hi, lo) fp16 pairs ----
    auto encode = [&](int64_t step, uint32_t (&hq)[8], uint32_t (&lq)[8]) {
      const int64_t p_raw = point_of(step);
      const int64_t p = p_raw < a.P ? p_raw : a.P - 1;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
      const int in_ch = 3 + 6 * a.multires;
      if (a.rays != nullptr) {
        const int64_t r = a.P < (int64_t)0x7fffffff ? (int64_t)((uint32_t)p / (uint32_t)a.S) : p / a.S;
        const float* ray = a.rays + r * a.ray_stride;
        const float zz = a.z[p];
        const float cen[3] = {a.cx, a.cy, a.cz};
        float x[3], xp[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float pt = __fadd_rn(ray[d], __fmul_rn(ray[3 + d], zz));     // RS:657
          x[d] = __fmul_rn(__fsub_rn(pt, cen[d]), a.bb_scale);               // RS:52
          xp[d] = __fmul_rn(x[d], 3.14159274101257324f);                     // H:165
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = 16 * g + i;                                        // warp-uniform
          if (col < 3) {
            v[i] = x[col];
          } else if (col < in_ch) {
            const int k = (col - 3) / 6, rem = (col - 3) - 6 * k;            // octave, position inside [sin3 | cos3]  (H:163-166)
            const int d = rem >= 3 ? rem - 3 : rem;
            const float arg = __fmul_rn(d == 0 ? xp[0] : (d == 1 ? xp[1] : xp[2]), (float)(1 << k));
            v[i] = rem >= 3 ? cosf(arg) : sinf(arg);
          } else if (col < in_ch + 3) {
            v[i] = ray[8 + col - in_ch];                                     // view direction (multires_views == 0)
          }
        }
      } else {
        const float* xin = a.x_embedded + p * a.in_all;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = 16 * g + i;
          if (col < a.in_all) v[i] = xin[col];
        }
      }
      if (g == 3) {
        v[ONES_COL - 48] = 1.0f;
        v[ONES_COL + 1 - 48] = 1.0f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) split_pack_f16x2(v[2 * i], v[2 * i + 1], &hq[i], &lq[i]);
    };
    auto store_emb_and_signal = [&](const uint32_t (&hq)[8], const uint32_t (&lq)[8]) {
      const uint32_t o0 = ((uint32_t)(2 * g) << 4) ^ rx, o1 = ((uint32_t)(2 * g + 1) << 4) ^ rx;
      *reinterpret_cast<uint4*>(emb_row + o0) = make_uint4(hq[0], hq[1], hq[2], hq[3]);
      *reinterpret_cast<uint4*>(emb_row + o1) = make_uint4(hq[4], hq[5], hq[6], hq[7]);
      *reinterpret_cast<uint4*>(emb_row + CHUNK_BYTES + o0) = make_uint4(lq[0], lq[1], lq[2], lq[3]);
      *reinterpret_cast<uint4*>(emb_row + CHUNK_BYTES + o1) = make_uint4(lq[4], lq[5], lq[6], lq[7]);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (remote) mbar_arrive_remote_relaxed(emb_target);
        else mbar_arrive(emb_target);
      }
    };

    if (unit0 < n_steps) {
      uint32_t hq[8], lq[8];
      encode(unit0, hq, lq);
      store_emb_and_signal(hq, lq);
    }

    for (int64_t step = unit0; step < n_steps; step += n_units) {
      const int64_t p_raw = point_of(step);
      const bool live = p_raw < a.P;
      float alpha = 0.f;                                  // this thread's share of alpha_linear (its 64 columns)
      for (int l = 0; l < plan.n_layers; ++l) {
        const int kind = plan.layers[l].kind, bias_epi = plan.layers[l].bias_epi;
        const uint32_t t_acc = t_lane + acc_par * W;
        if (kind != 3) {
          // fetched while the layer's MMAs run: one element of the rows that go through shared memory
          float s_b = 0.f, s_a = 0.f;
          if (bias_epi && tid < W) s_b = __ldg(tail->bias[l] + tid);
          if (kind == 1 && tid >= W) s_a = __ldg(tail->w_alpha + tid - W);
          mbar_wait(bar_acc, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
          if (bias_epi || kind == 1) {
            // no thread of this CTA can still be reading the previous rows: every warp signalled its last chunk of the previous
            // layer before this layer's MMAs could complete
            if (bias_epi && tid < W) sbias[tid] = s_b;
            if (kind == 1 && tid >= W) salpha[tid - W] = s_a;
            named_bar_sync(all_bar, PP_EPI_WARPS * 32);
          }
          if (kind == 1) {
            if (bias_epi) alpha = x3_hidden_epilogue<true, true, true>(t_acc, g, a_row, rx, sbias, salpha, chunk_target, remote, lane);
            else alpha = x3_hidden_epilogue<false, true, true>(t_acc, g, a_row, rx, sbias, salpha, chunk_target, remote, lane);
          } else if (kind == 2) {
            x3_hidden_epilogue<true, false, false>(t_acc, g, a_row, rx, sbias, salpha, chunk_target, remote, lane);
          } else if (bias_epi) {
            x3_hidden_epilogue<true, true, false>(t_acc, g, a_row, rx, sbias, salpha, chunk_target, remote, lane);
          } else {
            x3_hidden_epilogue<false, true, false>(t_acc, g, a_row, rx, sbias, salpha, chunk_target, remote, lane);
          }
        } else {
          // views layer (N = 128) + rgb_linear + output (H:238-242): this thread owns accumulator columns [32 g, 32 g + 32).
          // The next step's encoding is computed BEFORE waiting for this layer's MMAs and stored right after them.
          const int64_t next = step + n_units;
          const bool has_next = next < n_steps;
          uint32_t hq[8], lq[8];
          if (has_next) encode(next, hq, lq);
          const bool stager = (ew == 0);
          float4 wst[3];
          if (stager) {
#pragma unroll
            for (int i = 0; i < 3; ++i) wst[i] = __ldg(reinterpret_cast<const float4*>(&tail->w_rgb_p[0][0]) + i * 32 + lane);
          }
          mbar_wait(bar_acc, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
          uint32_t rr[32];
          tmem_ld32(t_acc + 32 * g, rr);
          tmem_ld_wait();
          tc_fence_before();
          if (has_next) store_emb_and_signal(hq, lq);      // the encoding chunk is free: this layer's MMAs (its last readers) retired
          // rgb_linear weights [3][128] + hand-over rows live in the (dead) lo half of activation chunk 3
          float* srgb = reinterpret_cast<float*>(smem + OFF_A + (X3_LO + 3) * CHUNK_BYTES);
          float4* scratch = reinterpret_cast<float4*>(smem + OFF_A + (X3_LO + 3) * CHUNK_BYTES + 2048);      // [3][128]
          if (stager) {
#pragma unroll
            for (int i = 0; i < 3; ++i) reinterpret_cast<float4*>(srgb)[i * 32 + lane] = wst[i];
          }
          named_bar_sync(all_bar, PP_EPI_WARPS * 32);
          const float4* wr4 = reinterpret_cast<const float4*>(srgb + 32 * g);
          const float4* wg4 = reinterpret_cast<const float4*>(srgb + 128 + 32 * g);
          const float4* wb4 = reinterpret_cast<const float4*>(srgb + 256 + 32 * g);
          float cr[2] = {0.f, 0.f}, cg[2] = {0.f, 0.f}, cb[2] = {0.f, 0.f};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 wr = wr4[q], wg = wg4[q], wb = wb4[q];
            const float h0 = fmaxf(__uint_as_float(rr[4 * q + 0]), 0.f), h1 = fmaxf(__uint_as_float(rr[4 * q + 1]), 0.f);
            const float h2 = fmaxf(__uint_as_float(rr[4 * q + 2]), 0.f), h3 = fmaxf(__uint_as_float(rr[4 * q + 3]), 0.f);
            cr[0] = fmaf(h0, wr.x, cr[0]); cg[0] = fmaf(h0, wg.x, cg[0]); cb[0] = fmaf(h0, wb.x, cb[0]);
            cr[1] = fmaf(h1, wr.y, cr[1]); cg[1] = fmaf(h1, wg.y, cg[1]); cb[1] = fmaf(h1, wb.y, cb[1]);
            cr[0] = fmaf(h2, wr.z, cr[0]); cg[0] = fmaf(h2, wg.z, cg[0]); cb[0] = fmaf(h2, wb.z, cb[0]);
            cr[1] = fmaf(h3, wr.w, cr[1]); cg[1] = fmaf(h3, wg.w, cg[1]); cb[1] = fmaf(h3, wb.w, cb[1]);
          }
          const float pr = cr[0] + cr[1], pg = cg[0] + cg[1], pb = cb[0] + cb[1];
          if (g != 0) {
            scratch[(g - 1) * TILE_M + row] = make_float4(pr, pg, pb, alpha);
            __threadfence_block();
            named_bar_arrive(quarter_bar, 128);
          } else {
            named_bar_sync(quarter_bar, 128);
            const float4 o1 = scratch[row], o2 = scratch[TILE_M + row], o3 = scratch[2 * TILE_M + row];
            const float al = ((alpha + o1.w) + (o2.w + o3.w)) + __ldg(&tail->b_alpha);
            if (live) {
              a.out[p_raw] = make_float4(((pr + o1.x) + (o2.x + o3.x)) + __ldg(&tail->b_rgb[0]),
                                         ((pg + o1.y) + (o2.y + o3.y)) + __ldg(&tail->b_rgb[1]),
                                         ((pb + o1.z) + (o2.z + o3.z)) + __ldg(&tail->b_rgb[2]), softplus_beta10(al));
            }
          }
          // every warp is done with the staged weights / hand-over rows before the next step's layer-0 epilogue overwrites them
          named_bar_sync(all_bar, PP_EPI_WARPS * 32);
        }
        acc_par ^= 1;
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
