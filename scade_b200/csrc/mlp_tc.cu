// SCADE_PREC_TC_F16: the NeRF field network as ONE persistent, warp-specialised tcgen05 kernel.
//
//   reference path replaced:  run_network (run_scade_scannet.py:48-63) -> Embedder.embed
//   (model/run_nerf_helpers.py:142-172) -> NeRF.forward (H:223-247), i.e. 19 elementwise launches
//   + 12 cuBLAS SGEMMs + cats per call, each round-tripping [P,256] fp32 activations through HBM.
//
// Design (B200, sm_100a) -- nerf_mlp_tc_pp_kernel
//   * Clusters of 2 CTAs on an SM pair, persistent over 512-point steps.  A step = 2 "super-tiles" of 256 points
//     (row tile t of both CTAs); MMAs are tcgen05.mma.cta_group::2 with M = 256, N = 256 (128 for the views layer), K = 16.
//   * Activations never leave the SM: fp16 A operand tiles live in shared memory in the canonical K-major SWIZZLE_128B
//     layout (4 chunks of [128 x 64] per tile + one [128 x 64] "encoding chunk" holding gamma(x), the view direction, two
//     constant-1 columns and zero padding); accumulators live in TMEM (2 tiles x 256 fp32 columns = all 512 columns).
//   * Weights are pre-packed (scade_mlp_pack_f16) into the exact shared-memory image of each [128 (N) x 64 (K)] fp16 stage,
//     in consumption order; each CTA streams only its own N half of every weight block, once per layer for both tiles,
//     through a 4-slot mbarrier ring filled by 2D-tiled TMA whose completion lands on the leader CTA's barrier.
//   * The two tiles ping-pong: while tile 1's MMAs of layer l run, tile 0's epilogue warps turn its layer-l accumulators into
//     its layer-(l+1) operand, and vice versa.  Every wide layer is exactly 4 ring stages (no bias stages; the views layer
//     packs two K chunks per stage), so the ring never blocks a tile on a slot the other tile still holds.
//   * warp 0: TMA producer; warp 1: TMEM allocator + single-thread tcgen05.mma issuer; warp 2: the compositor of the kComp
//     variant (alpha compositing of the values the views epilogue parks for it, see CompArgs), else idle; warp 3 idle; the
//     control warpgroup hands its registers to the epilogue warpgroups with setmaxnreg; warps 4..19: prologue / epilogue
//     (thread == one point row x 128 columns): positional encoding straight into the swizzled A tile, then per layer
//     TMEM -> registers -> + fp32 bias (row broadcast from shared memory) -> ReLU -> fp16 -> swizzled A tile of the next layer.
//   * Layers whose K includes the encoding chunk (layer 0, the skip layer, the views layer) get their bias from the tensor core:
//     the packed weights carry fp16 hi/lo halves of the bias in the K positions of the chunk's two constant-1 columns.
//   * alpha_linear (256->1) is fused into the last hidden layer's epilogue and rgb_linear (128->3) into the views epilogue, both
//     as fp32 dot products on the un-rounded fp32 activations; softplus(beta=10) is applied before the single float4 store of
//     (rgb_raw, sigma) per point -- the only HBM write of the kernel (kComp: weights and the per-ray maps instead).
//   * skip connection (H:230) and view concat (H:235) are extra K chunks that re-use the encoding chunk; nothing is
//     concatenated in memory.
//   * kStash = true additionally leaves every layer's fp16 operand tile, ReLU sign masks and the alpha pre-activation in the
//     training stash (TMA bulk stores); mlp_tc_train.cuh holds the backward kernels.
//   * -DSCADE_TC_TRACE=1 (tools/tc_trace.py) compiles in clock stamps and timing ablations.
#include <cuda.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <type_traits>

#include "composite.cuh"
#include "mlp_common.cuh"
#include "tc_ptx.cuh"

namespace scade {

namespace tc {

constexpr int TILE_M = 128;              // rows per tcgen05.mma (one TMEM lane per row)
constexpr int TILES = 2;                 // row tiles per CTA step
constexpr int W = 256;                   // layer width handled by this kernel
constexpr int KCHUNK = 64;               // fp16 elements per 128-byte swizzled row
constexpr int STAGE_N = 128;             // weight rows per stage
constexpr int STAGE_BYTES = STAGE_N * KCHUNK * 2;   // 16 KB
constexpr int CHUNK_BYTES = TILE_M * KCHUNK * 2;    // 16 KB
constexpr int NUM_STAGES = 4;
constexpr int MAX_LAYERS = 12;           // D (<= 8) + feature + views
constexpr int MAX_STAGE_DESCS = 160;         // the hi/lo stream of SCADE_PREC_TC_F16X3 has 148 stages for 8x256

// shared-memory map (relative to a 1024-byte aligned base)
constexpr int OFF_A = 0;                                 // [TILES][4 chunks]
constexpr int OFF_EMB = OFF_A + TILES * 4 * CHUNK_BYTES;  // [TILES]
constexpr int OFF_STAGE = OFF_EMB + TILES * CHUNK_BYTES;  // [NUM_STAGES]
constexpr int OFF_BAR = OFF_STAGE + NUM_STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;         // barriers + alignment slack (dgrad kernel)
// forward kernel: + one 256-float staging row per tile for the current layer's fp32 bias; no slack -- the dynamic shared
// memory base is 1024-byte aligned when the kernel has no static shared memory (checked at kernel start)
constexpr int OFF_BIAS = OFF_BAR + 256;
constexpr int FWD_SMEM_BYTES = OFF_BIAS + TILES * W * 4;
static_assert(FWD_SMEM_BYTES <= 227 * 1024, "forward kernel shared memory exceeds the 227 KB per-CTA limit");
// kComp (compositing fused into the network kernel): the views epilogue parks each point's (rgb_raw, sigma) in a 4 KB
// per-CTA slot of an L2-resident ring in the call's workspace; the CTA's compositor warp picks it up from there
constexpr int COMP_RING_BYTES_PER_CTA = TILES * TILE_M * (16 + 4);   // [256] float4, then [256] float (last step only)

enum ASrc : int { SRC_EMB = 4 };         // 0..3 = activation chunk c

struct LayerDesc {
  int n_k;                // K chunks
  int a_src[5];           // source chunk of each K chunk
  int relu;
  int kind;               // 0 = hidden, 1 = last hidden (also computes alpha), 2 = feature, 3 = views (final)
  int bias_epi;           // 1: the epilogue adds the fp32 bias (layer has no encoding K chunk whose 1.0 columns carry it)
  int n_out;              // output width (256, or 128 for the views layer)
  int a_step;             // issuer: K chunk i (after the encoding chunk) reads activation chunk a_step * i
  int kpack;              // K chunks per ring stage (2 when a CTA's N half is only 64 rows: views layer)
  int emb_ks0;            // first K=16 step of the encoding chunk this layer needs (views layer: 3 -- only the view direction and
                          // the 1.0 columns, K positions 48..63, have non-zero weights)
  int bias_stage;         // 1: one extra ring stage carries the bias as a K=16 MMA step against the encoding chunk's 1.0 columns
                          //    (last hidden layer: its shared-memory row is taken by the alpha_linear weights)
  int a_tmem;             // 1: the activation K chunks of this layer are read from TENSOR MEMORY (columns [0, 128) of the tile's
                          //    region: 256 fp16 per row, two per column) -- the views layer of the render forward, whose N = 128
                          //    MMAs are bound by the shared-memory A fetch, not by the math
  int d_col;              // accumulator column offset inside the tile's 256-column region (128 for that views layer)
  int out_tmem;           // 1: the epilogue leaves this layer's fp16 output in tensor memory (the feature layer feeding it)
};

struct NetPlan {
  int n_layers;
  int stages_per_pass;
  LayerDesc layers[MAX_LAYERS];
  int first_stage[MAX_LAYERS];   // the layer's stages are [first_stage, first_stage + n_stages) of the packed stream
  int n_stages[MAX_LAYERS];
};

// small fp32 table behind the stage images in the packed buffer
struct PackedTail {
  float4 w_rgb[W / 2];     // (r, g, b, 0) weights of rgb_linear per hidden column (dgrad prologue)
  float w_alpha[W];
  float b_alpha, b_rgb[3];
  float w_rgb_p[3][W / 2]; // the same weights, one row per colour channel (forward epilogue: 4 columns per 16-byte load)
  float bias[MAX_LAYERS][W];   // fp32 biases of the layers whose epilogue adds them (bias_epi)
};

// One ring stage = up to two row blocks: block j fills stage rows [dst_row0, dst_row0 + nrows) with K chunk columns.
struct StageBlock {
  int col0; int ncols; int dst_col0; int dst_row0; int nrows;
  int bias_mode;                      // 0 none; 1: cols 60/61 <- hi/lo fp16 halves of bias[row0 + n]; 2: bias stage (cols 12/13)
};
struct StageDesc {
  const float* W; int ld; int row0;   // stage row (dst_row0 + n) of a block <- weight row row0 + n
  const float* bias;
  int trans;                          // 1: element (n, k) = W[(col0 + k) * ld + row0 + n]  (dgrad: B = W^T)
  int n_blocks;
  int lo;                             // 1: the stage holds the fp16 residuals fp16(w - fp16(w)) (SCADE_PREC_TC_F16X3), zeros at bias positions
  StageBlock blk[2];
};
struct PackPlan {
  int n_stages;                       // stage descriptors in st[]
  int n_fwd;                          // the first n_fwd are the forward stream, block n_fwd of the launch writes the fp32 tail
  unsigned long long bwd_off;         // byte offset of the stages behind the tail (dgrad stream)
  StageDesc st[MAX_STAGE_DESCS];
  const float* w_alpha; const float* b_alpha; const float* w_rgb; const float* b_rgb;
  const float* epi_bias[MAX_LAYERS];  // bias vector of layer l if its epilogue adds it, else nullptr
};
constexpr int ONES_COL = 60;      // encoding-chunk columns 60 and 61 hold 1.0 (bias hi / lo ride on them)

// ---- timeline tracing (tools/tc_trace.py; compiled in only with -DSCADE_TC_TRACE=1, a separate library) ----------------
// Lane 0 of every warp of CTA 0 appends (tag << 48 | clock) words to its own 8192-entry region of a global buffer.
#ifndef SCADE_TC_TRACE
#define SCADE_TC_TRACE 0
#endif
#if SCADE_TC_TRACE
constexpr int TRACE_CAP = 8192;
__device__ unsigned long long* g_trace_buf = nullptr;
struct Tracer {
  unsigned long long* buf; int n;
  __device__ __forceinline__ Tracer() : buf(nullptr), n(1) {
    if (g_trace_buf != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0) buf = g_trace_buf + (threadIdx.x >> 5) * TRACE_CAP;
  }
  __device__ __forceinline__ void operator()(int tag) {
    if (buf != nullptr && n < TRACE_CAP) { buf[n++] = ((unsigned long long)tag << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull); }
  }
  __device__ __forceinline__ ~Tracer() { if (buf != nullptr) buf[0] = (unsigned long long)n; }
};
#define TRACE(tr, tag) (tr)(tag)
#else
struct Tracer { };
#define TRACE(tr, tag) do { } while (0)
#endif

// sin/cos of arguments up to ~pi*2^8 for fp16 consumers: two-term Cody-Waite reduction by 2*pi, then the
// MUFU approximations on [-pi, pi] (abs error ~5e-7, three orders below the fp16 rounding that follows).
__device__ __forceinline__ void sincos_reduced(float arg, float* s, float* c) {
  float q = rintf(arg * 0.15915494309189535f);
  float r = fmaf(q, -6.2831854820251465f, arg);
  r = fmaf(q, 1.7484555e-07f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}

// ---- the kernel ------------------------------------------------------------------------------------
// Alpha compositing (compute_weights + raw2outputs, run_scade_scannet.py:511-562) fused into the network kernel (kComp).
// Rays mode, S a multiple of 32.  The views epilogue (the critical loop of the kernel: it co-limits the step with the tensor
// pipe) only parks the point's four raw values; warp 2 of the CTA -- otherwise idle -- is the COMPOSITOR: it walks the CTA's
// points in order, 32 samples of one ray at a time, with the running transmittance and the per-lane partial sums of the ray
// in registers, exactly like the stand-alone composite_ray_fwd (composite.cuh): same device functions, same operation order,
// bit-identical results.  It has a whole step (~58k cycles) for 8 chunks of ~3k cycles.
//   S in {32, 64, 128, 256}: steps are strided over the clusters; a CTA's 256 points per step hold whole rays.
//   any other S ("chain" mode, e.g. the 64 + 128 = 192 samples of the reference's config file): each CTA walks a CONTIGUOUS
//   range of chain_iters * 256 points that starts and ends on a ray boundary, so a ray that straddles tiles or steps stays
//   with one compositor.
struct CompArgs {
  float* weights;                 // [N,S]
  float* rgb_map;                 // [N,3]  nullable
  float* disp_map;                // [N]    nullable
  float* acc_map;                 // [N]    nullable
  float* depth_map;               // [N]    nullable
  int write_raw;                  // also store raw [N,S,4] (retraw)
  int chain_iters;                // 0: steps strided over the clusters; > 0: chain mode, steps per CTA
  int tail_mode;                  // last step of a CTA: the epilogue threads evaluate the per-sample terms (pp_compositor)
  int pad;
  uint8_t* ring;                  // [CTAs][COMP_RING_BYTES_PER_CTA] hand-over slots (workspace)
};

struct FwdArgs {
  const uint8_t* packed;          // stage images, consumption order
  const float* rays; int ray_stride; const float* z; int S;     // rays mode
  const float* x_embedded; int in_all;                          // embedded mode (rays == nullptr)
  int64_t P;
  float cx, cy, cz, bb_scale;
  int multires, multires_views;
  float4* out;
  int64_t n_pairs;
  int dbg;              // trace build only -- timing ablations (results are WRONG when set): 1 = ignore weight arrival,
                        // 2 = ignore operand readiness, 32 = no epilogue work, 64 = compositor: hand-shake only,
                        // 128 = no compositor and no parking at all
  CompArgs comp;        // kComp only
};

// Activation / gradient stash of the tensor-core training path (workspace of a forward call with save_for_backward=1).
// Every activation region is an array of 16 KB "chunks": the exact shared-memory image of a [128 points x 64 features]
// fp16 tile in the K-major SWIZZLE_128B layout the forward MMAs consume -- which, read with points as the K dimension,
// is also the canonical MN-major SWIZZLE_128B operand tile of the weight-gradient MMAs.  T = number of 128-point tiles.
struct TrainLayout {
  long long T;
  int D, pad;
  unsigned long long emb;        // [T][1]   encoding chunk (gamma(x), view dir, two 1.0 columns)
  unsigned long long feat;       // [T][4]   feature_linear output (no activation)
  unsigned long long hv;         // [T][2]   views layer output (post-ReLU)
  unsigned long long dzv;        // [T][2]   d loss / d (views pre-activation), scaled
  unsigned long long dzf;        // [T][4]   d loss / d feature
  unsigned long long maskv;      // [T][128] uint4: sign bits of the views pre-activations (sign_mask32 order)
  unsigned long long alpha;      // [T*128]  fp32 alpha pre-activation (H:233)
  unsigned long long gs;         // 64 uint32: [0] = bit pattern of max(|d_rgb_raw|, |d_alpha|) over the call's points
  unsigned long long total;
  unsigned long long h[8];       // [T][4]   output of pts_linears[l] (post-ReLU)
  unsigned long long dz[8];      // [T][4]   d loss / d (pre-activation of pts_linears[l]), scaled
  unsigned long long maskh[8];   // [T][2 column halves][128 rows] uint4: sign bits of pts_linears[l] pre-activations (a warp stores 512 B)
};
struct StashArgs {
  uint8_t* ws;
  TrainLayout L;
};

// =====================================================================================================
// v6: SM-pair kernel with the two row tiles ping-ponging on ONE shared weight ring
// =====================================================================================================
// Cluster of 2 CTAs.  "Super-tile" t = row tile t of both CTAs (256 points); the leader CTA issues
// tcgen05.mma.cta_group::2 (M=256, N=256/128).  Each CTA streams only its own N half of every weight block, ONCE per
// layer, through a 4-slot ring.  Super-tile 0 runs up to 4 stages ahead of super-tile 1; a slot is released by
// super-tile 1's MMAs.  So while tile 1's MMAs of layer l occupy the tensor pipe, tile 0's epilogue warps turn its
// layer-l accumulators into the layer-(l+1) operand, and vice versa: epilogue, barrier round trips and pipeline
// fill/drain of one tile hide under the other tile's MMAs, with no extra weight traffic and no extra shared memory.
// 16 epilogue warps (8 per tile: thread = one row x 128 columns) keep a tile's epilogue shorter than a layer of MMAs.
constexpr int PP_EPI_WARPS = 16;
constexpr int PP_EPI_WARP0 = 4;                        // warpgroup 0: TMA producer, MMA issuer, 2 idle warps; warps 4..19: prologue / epilogue
constexpr int PP_THREADS = (PP_EPI_WARP0 + PP_EPI_WARPS) * 32;
// register hand-over: the control warpgroup shrinks to 32 registers per thread, the four epilogue warpgroups grow from the
// launch allocation (96) to 112  (128 x 32 + 512 x 112 = 61,440 = the 640 x 96 the CTA owns; asking for more would block forever)
__device__ __forceinline__ void regs_control() { asm volatile("setmaxnreg.dec.sync.aligned.u32 32;"); }
__device__ __forceinline__ void regs_epilogue() { asm volatile("setmaxnreg.inc.sync.aligned.u32 112;"); }

// ---- roles shared by the ping-pong forward kernel and the dgrad kernel ---------------------------------------------
// TMA producer of this CTA's half stages (every stage once per layer).  The packed stream is addressed through a 2D
// tensor map ([rows of 128 B] x 128 rows per stage); completion of BOTH CTAs' copies lands on the leader's barrier.
__device__ __forceinline__ void pp_weight_producer(const NetPlan& plan, const CUtensorMap* tmap, uint32_t sbase,
                                                   uint32_t bar_full, uint32_t bar_empty, uint32_t cta_rank, int64_t unit0,
                                                   int64_t n_steps, int64_t n_units) {
  uint32_t stage = 0, phase = 0;
  for (int64_t step = unit0; step < n_steps; step += n_units) {
    for (int l = 0; l < plan.n_layers; ++l) {
      const int first = plan.first_stage[l], last = first + plan.n_stages[l];
      for (int s = first + (int)cta_rank; s < last; s += 2) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (cta_rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * STAGE_BYTES);
        tma_load_2d_pair(sbase + OFF_STAGE + stage * STAGE_BYTES, tmap, 0, s * STAGE_N, bar_full + 8 * stage);
        if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }
}

// MMA issuer of the leader CTA (whole warp walks the loops, one elected lane issues).  Ring stage i of a layer holds
// `kpack` consecutive K chunks of this CTA's N half (kpack = 2 for the views layer, whose half is only 64 rows).
// Layers of <= NUM_STAGES stages: super-tile 0 consumes the whole layer, then super-tile 1 does, releasing the slots as it
// goes -- by then the next layer's stages are already in flight, so the ring never blocks.  The one 5-stage layer (skip
// layer: encoding chunk + 4 activation chunks) interleaves the tiles so that the 5th stage is requested two stages before
// it is needed (trace-measured commit -> reload -> full latency is ~1600 cycles = 12 MMAs).
__device__ __forceinline__ void pp_mma_issuer(const NetPlan& plan, uint32_t sbase, uint32_t tmem_base, uint32_t bar_full,
                                              uint32_t bar_empty, uint32_t bar_acc, uint32_t bar_aready, int64_t unit0,
                                              int64_t n_steps, int64_t n_units, int dbg = 0) {
  constexpr uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
  // ring position of each super-tile (tile 1 trails tile 0)
  uint32_t slot0 = 0, slot1 = 0, ph0 = 0, ph1 = 0, a_ph0 = 0, a_ph1 = 0;
  [[maybe_unused]] Tracer tr;
  for (int64_t step = unit0; step < n_steps; step += n_units) {
    for (int l = 0; l < plan.n_layers; ++l) {
      const int n_k = plan.layers[l].n_k, kpack = plan.layers[l].kpack;
      const int n_kst = (n_k + kpack - 1) / kpack;             // stages holding K chunks
      const int n_own = n_kst + plan.layers[l].bias_stage;     // this layer's stages per CTA (<= 5)
      const int has_emb = plan.layers[l].a_src[0] == SRC_EMB;
      const int a_step = plan.layers[l].a_step;                // A chunk of K chunk i (after the encoding chunk) = a_step * i
      const int emb_ks0 = plan.layers[l].emb_ks0;
      const uint32_t idesc = make_idesc(2 * TILE_M, plan.layers[l].n_out);
      const int a_tmem = plan.layers[l].a_tmem, d_col = plan.layers[l].d_col;
      // one (stage, tile): wait for the slot in both CTAs, issue its MMAs for super-tile t, optionally release it
      auto consume_at = [&](int t, uint32_t& slot_ref, uint32_t& ph_ref, int i, bool release) {
        const uint32_t sl = slot_ref, p = ph_ref;
#if SCADE_TC_TRACE
        if (!(dbg & 1))                                  // timing ablation: ignore weight arrival (results are wrong)
#endif
        mbar_wait(bar_full + 8 * sl, p);                 // both CTAs' halves of the stage have landed
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_addr = tmem_base + t * W + d_col;
          if (i == n_kst) {
            // bias stage: A = K-step 3 of the encoding chunk (the two 1.0 columns), B = K-step 0 (bias hi/lo at k = 12, 13)
            const uint64_t a_desc = desc_hi | (uint64_t)(((sbase + OFF_EMB + t * CHUNK_BYTES + 3 * 32) & 0x3FFFF) >> 4);
            const uint64_t b_desc = desc_hi | (uint64_t)(((sbase + OFF_STAGE + sl * STAGE_BYTES) & 0x3FFFF) >> 4);
            mma_f16_ss_pair(d_addr, a_desc, b_desc, idesc, 1u);
          }
          for (int j = 0; j < kpack && i < n_kst; ++j) {
            const int kc = i * kpack + j;
            if (kc >= n_k) break;
            const bool emb = has_emb && kc == 0;
            const uint32_t a_addr = emb ? sbase + OFF_EMB + t * CHUNK_BYTES
                                        : sbase + OFF_A + (t * 4 + (kc - has_emb) * a_step) * CHUNK_BYTES;
            const uint64_t a_desc = desc_hi | (uint64_t)((a_addr & 0x3FFFF) >> 4);
            const uint64_t b_desc = desc_hi | (uint64_t)(((sbase + OFF_STAGE + sl * STAGE_BYTES + j * (STAGE_BYTES / 2)) & 0x3FFFF) >> 4);
            const int ks0 = emb ? emb_ks0 : 0;                 // the layer's first MMA (K chunk 0, step ks0) overwrites the accumulator
            if (a_tmem && !emb) {
              // A from tensor memory: activation chunk c = 64 K values = 32 columns of the tile's region, 8 columns per K step
              const uint32_t a_col = tmem_base + t * W + (kc - has_emb) * a_step * (KCHUNK / 2);
#pragma unroll
              for (int ks = 0; ks < KCHUNK / 16; ++ks)
                mma_f16_ts_pair(d_addr, a_col + 8 * ks, b_desc + 2 * ks, idesc, (kc != 0 || ks != 0) ? 1u : 0u);
            } else {
#pragma unroll
              for (int ks = 0; ks < KCHUNK / 16; ++ks)
                if (ks >= ks0) mma_f16_ss_pair(d_addr, a_desc + 2 * ks, b_desc + 2 * ks, idesc, (kc != 0 || ks != ks0) ? 1u : 0u);
            }
          }
          if (release) mma_commit_pair(bar_empty + 8 * sl, (uint16_t)3);
        }
        __syncwarp();
        if (++slot_ref == NUM_STAGES) { slot_ref = 0; ph_ref ^= 1; }
      };
      auto consume = [&](int t, int i, bool release) {
        if (t == 0) consume_at(0, slot0, ph0, i, release);
        else consume_at(1, slot1, ph1, i, release);
      };
      auto wait_a = [&](int t) {
        TRACE(tr, 0x100 + 2 * l + t);                    // issuer: begins waiting for tile t's operand of layer l
#if SCADE_TC_TRACE
        if (dbg & 2) { TRACE(tr, 0x200 + 2 * l + t); return; }     // timing ablation: ignore operand readiness
#endif
        if (t == 0) { mbar_wait(bar_aready, a_ph0); a_ph0 ^= 1; }
        else { mbar_wait(bar_aready + 8, a_ph1); a_ph1 ^= 1; }
        tc_fence_after();
        TRACE(tr, 0x200 + 2 * l + t);                    // issuer: operand ready
      };
      auto publish = [&](int t) {
        if (elect_one()) mma_commit_pair(bar_acc + 8 * t, (uint16_t)3);
        __syncwarp();
        TRACE(tr, 0x300 + 2 * l + t);                    // issuer: all MMAs of (layer l, tile t) issued + commit
      };
      wait_a(0);
      if (n_own <= NUM_STAGES) {
        for (int i = 0; i < n_own; ++i) consume(0, i, false);
        publish(0);
        wait_a(1);
        for (int i = 0; i < n_own; ++i) consume(1, i, true);
        publish(1);
      } else {
        // 5 stages on a 4-slot ring: a0 a1 a2 | b0 | a3 | b1 | a4 | b2 b3 b4   (a = tile 0, b = tile 1)
        consume(0, 0, false); consume(0, 1, false); consume(0, 2, false);
        wait_a(1);
        consume(1, 0, true);
        consume(0, 3, false);
        consume(1, 1, true);
        consume(0, 4, false);
        publish(0);
        consume(1, 2, true); consume(1, 3, true); consume(1, 4, true);
        publish(1);
      }
    }
  }
}

// One thread's share of a hidden / feature layer epilogue: 128 fp32 accumulator columns (TMEM, 32 at a time, double
// buffered) -> (+ bias row from shared memory) -> (ReLU) -> fp16 -> the two swizzled A chunks at `dst` (row offset applied;
// rx4 = (row & 7) << 4 is the row's swizzle term).  kAlpha: also the dot product of the un-rounded ReLU'd activations with
// the 128 alpha_linear weights at `srow` (H:233).  kMask: sign masks of the pre-activations for the training stash.
// Branch-free inside; the caller dispatches on the layer's flags.
template <bool kBias, bool kRelu, bool kAlpha, bool kMask>
__device__ __forceinline__ float hidden_epilogue(uint32_t t_col, uint8_t* dst, uint32_t rx4, const float* srow, float4 wa,
                                                 uint32_t* sgn) {
  uint32_t rb[2][32];
  float al0 = 0.f, al1 = 0.f, al2 = 0.f, al3 = 0.f;
  tmem_ld32(t_col, rb[0]);
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    uint32_t* r = rb[c4 & 1];
    tmem_ld_wait();
    if (c4 + 1 < 4) tmem_ld32(t_col + (c4 + 1) * 32, rb[(c4 + 1) & 1]);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (kBias) {                                         // 4 broadcast LDS.128 in flight, then their 16 columns
        const float4* p4 = reinterpret_cast<const float4*>(srow + c4 * 32 + g * 16);
        const float4 v0 = p4[0], v1 = p4[1], v2 = p4[2], v3 = p4[3];
        const float v[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) r[g * 16 + i] = __float_as_uint(__uint_as_float(r[g * 16 + i]) + v[i]);
      }
      if (kAlpha) {
        // lane j of the warp holds the alpha_linear weights of columns 4j..4j+3 (wa): 16 shuffles fetch this group's 16
        float v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int src = c4 * 8 + g * 4 + j;
          v[4 * j + 0] = __shfl_sync(0xffffffffu, wa.x, src);
          v[4 * j + 1] = __shfl_sync(0xffffffffu, wa.y, src);
          v[4 * j + 2] = __shfl_sync(0xffffffffu, wa.z, src);
          v[4 * j + 3] = __shfl_sync(0xffffffffu, wa.w, src);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float hx = fmaxf(__uint_as_float(r[g * 16 + i]), 0.f);
          if ((i & 3) == 0) al0 = fmaf(hx, v[i], al0);
          else if ((i & 3) == 1) al1 = fmaf(hx, v[i], al1);
          else if ((i & 3) == 2) al2 = fmaf(hx, v[i], al2);
          else al3 = fmaf(hx, v[i], al3);
        }
      }
    }
    if (kMask) sgn[c4] = sign_mask32(r);
    uint8_t* chunk = dst + (c4 >> 1) * CHUNK_BYTES;
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
      uint4 q;
      if (kRelu) {
        q.x = pack_relu_f16x2(r[pc * 8 + 0], r[pc * 8 + 1]);
        q.y = pack_relu_f16x2(r[pc * 8 + 2], r[pc * 8 + 3]);
        q.z = pack_relu_f16x2(r[pc * 8 + 4], r[pc * 8 + 5]);
        q.w = pack_relu_f16x2(r[pc * 8 + 6], r[pc * 8 + 7]);
      } else {
        q.x = pack_plain_f16x2(r[pc * 8 + 0], r[pc * 8 + 1]);
        q.y = pack_plain_f16x2(r[pc * 8 + 2], r[pc * 8 + 3]);
        q.z = pack_plain_f16x2(r[pc * 8 + 4], r[pc * 8 + 5]);
        q.w = pack_plain_f16x2(r[pc * 8 + 6], r[pc * 8 + 7]);
      }
      *reinterpret_cast<uint4*>(chunk + ((uint32_t)((((c4 & 1) * 4 + pc) << 4)) ^ rx4)) = q;
    }
  }
  return (al0 + al1) + (al2 + al3);
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// feature_linear's epilogue when the views layer reads its A operand from tensor memory: fp32 accumulators (+ bias, no
// activation, H:234) -> fp16 pairs -> columns [0, 128) of the tile's own TMEM region (row = lane, column j = features 2j, 2j+1).
// The two column-half threads of a row interleave 32-column chunks (thread `half` takes chunks half, 2 + half, 4 + half, 6 + half)
// and write the 16 packed columns of chunk i at [16 i, 16 i + 16).  The packed image overwrites accumulator columns that the
// OTHER thread of the pair reads (chunks 0..3), so the stores start only after a 64-thread barrier that both pass once their
// loads of chunks 0..3 have completed; the first chunk's result waits in registers until then.
__device__ __forceinline__ void feature_epilogue_tmem(uint32_t t_tile, int half, const float* sbias_tile, int pair_bar) {
  uint32_t r[32], first[16];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int i = 2 * s + half;                                 // 32-column chunk
    tmem_ld32(t_tile + 32 * i, r);
    tmem_ld_wait();
    if (s == 1) named_bar_sync(pair_bar, 64);
    const float4* p4 = reinterpret_cast<const float4*>(sbias_tile + 32 * i);
    uint32_t pk[16];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = p4[q];
      pk[2 * q] = pack_plain_f16x2(__float_as_uint(__uint_as_float(r[4 * q]) + b.x), __float_as_uint(__uint_as_float(r[4 * q + 1]) + b.y));
      pk[2 * q + 1] = pack_plain_f16x2(__float_as_uint(__uint_as_float(r[4 * q + 2]) + b.z), __float_as_uint(__uint_as_float(r[4 * q + 3]) + b.w));
    }
    if (s == 0) {
#pragma unroll
      for (int q = 0; q < 16; ++q) first[q] = pk[q];
    } else {
      if (s == 1) tmem_st16(t_tile + 16 * half, first);
      tmem_st16(t_tile + 16 * i, pk);
    }
  }
  tmem_st_wait();
}

// kStash: additionally write every layer's fp16 output chunk (TMA bulk store of the shared-memory image the next layer's
// MMAs read anyway), the ReLU sign masks and the alpha pre-activation to the training stash.
// The compositor warp of the kComp kernel (see CompArgs).  Runs with the control warps' 32 registers: its loop is latency-
// tolerant (one 32-sample chunk at a time, the next chunk's loads in flight).  In the CTA's LAST step nothing is left to hide
// its ~3k cycles per chunk behind, so there the epilogue threads evaluate the per-sample terms themselves (in parallel, ~300
// cycles once per launch) and park (sigmoid(rgb), alpha | 1 - alpha + 1e-10); the compositor is left with the scan and the sums.
__device__ __noinline__ void pp_compositor(const FwdArgs& a, uint32_t bar_raw, uint32_t cta_rank, int64_t unit0, int64_t n_steps,
                                           int64_t n_units, int lane) {
  const int S = a.S;
  const int64_t cta = 2 * unit0 + cta_rank;
  const float4* slot = reinterpret_cast<const float4*>(a.comp.ring + cta * COMP_RING_BYTES_PER_CTA) + lane;
  const float* slot_tf = reinterpret_cast<const float*>(a.comp.ring + cta * COMP_RING_BYTES_PER_CTA + TILES * TILE_M * 16) + lane;
  float carry = 1.0f, norm = 0.f, sr = 0.f, sg = 0.f, sb = 0.f, sdepth = 0.f, sacc = 0.f;
  int it = 0;
  [[maybe_unused]] Tracer tr;
  for (int64_t step = unit0; step < n_steps; step += n_units, ++it) {
    const int64_t base = (a.comp.chain_iters > 0 ? cta * (int64_t)a.comp.chain_iters + it : 2 * step + cta_rank) * (int64_t)(TILES * TILE_M);
    const bool light = a.comp.tail_mode && step + n_units >= n_steps;   // last step: the terms arrive evaluated
    for (int tile = 0; tile < TILES; ++tile) {
      int64_t p = base + tile * TILE_M + lane;                          // a chunk is wholly live or wholly dead: 32 | S | P
      const bool tile_live = p < a.P;
      int64_t r = 0;
      int i = 0;                                                        // sample index within the ray
      float zi = 0.f, zn = 0.f;
      if (tile_live) {                                                  // (z does not depend on the network: fetched before the wait)
        r = a.P < (int64_t)0x7fffffff ? (int64_t)((uint32_t)p / (uint32_t)S) : p / S;
        i = (int)(p - r * S);
        zi = a.z[p];
        zn = (i + 1 < S) ? a.z[p + 1] : zi;
      }
      TRACE(tr, 0x800 + tile);                                          // compositor: waits for the tile's values
      mbar_wait(bar_raw + 8 * tile, (uint32_t)it & 1u);
      TRACE(tr, 0x900 + tile);                                          // ... parked
#if SCADE_TC_TRACE
      if (a.dbg & 64) {                                                 // timing ablation: hand-shake only
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_raw + 16 + 8 * tile);
        continue;
      }
#endif
      if (tile_live) {
        const float4* src = slot + tile * TILE_M;
        const float* src_tf = slot_tf + tile * TILE_M;
        float4 rw = __ldcg(src);
        float tf = light ? __ldcg(src_tf) : 0.f;
        for (int q = 0; q < 4; ++q) {
          // the next chunk's operands are requested before this chunk's arithmetic
          const bool more = q < 3 && p + 32 < a.P;
          const int i_n = (i + 32 - lane == S) ? lane : i + 32;
          float4 rw_n = rw;
          float zi_n = 0.f, zn_n = 0.f, tf_n = 0.f;
          if (more) {
            rw_n = __ldcg(src + (q + 1) * 32);
            zi_n = a.z[p + 32];
            if (light) tf_n = __ldcg(src_tf + (q + 1) * 32);
            else zn_n = (i_n + 1 < S) ? a.z[p + 33] : zi_n;
          }
          if (i == lane) {                                              // first chunk of a ray
            carry = 1.0f; sr = 0.f; sg = 0.f; sb = 0.f; sdepth = 0.f; sacc = 0.f;
            if (!light) {                                               // RS:516
              const float* rd = a.rays + r * a.ray_stride + 3;
              const float dx = rd[0], dy = rd[1], dz = rd[2];
              norm = sqrtf(dx * dx + dy * dy + dz * dz);
            }
          }
          float al = rw.w, cr = rw.x, cg = rw.y, cb = rw.z;
          if (!light) {
            const SampleTerms t = sample_terms(rw.w, 0.f, zi, zn, i == S - 1, norm);
            al = t.alpha; tf = t.tfac;
            cr = sigmoidf_(rw.x); cg = sigmoidf_(rw.y); cb = sigmoidf_(rw.z);   // RS:543
          }
          const float incl = warp_scan_prod(tf, lane);
          float excl = __shfl_up_sync(FULL, incl, 1);
          if (lane == 0) excl = 1.0f;
          const float w = al * (carry * excl);
          carry *= __shfl_sync(FULL, incl, 31);
          a.comp.weights[p] = w;
          sr = fmaf(w, cr, sr);                                         // RS:556
          sg = fmaf(w, cg, sg);
          sb = fmaf(w, cb, sb);
          sdepth = fmaf(w, zi, sdepth);                                 // RS:558
          sacc += w;                                                    // RS:560
          if (i_n == lane) {                                            // that was the last chunk of the ray
            const float tr = warp_sum(sr), tg = warp_sum(sg), tb = warp_sum(sb), td = warp_sum(sdepth), ta = warp_sum(sacc);
            if (lane == 0) {
              if (a.comp.rgb_map) { a.comp.rgb_map[r * 3] = tr; a.comp.rgb_map[r * 3 + 1] = tg; a.comp.rgb_map[r * 3 + 2] = tb; }
              if (a.comp.depth_map) a.comp.depth_map[r] = td;
              if (a.comp.acc_map) a.comp.acc_map[r] = ta;
              if (a.comp.disp_map) {
                const float qd = td / ta;                               // RS:559; torch.max propagates the nan of 0/0
                a.comp.disp_map[r] = (qd != qd) ? qd : 1.0f / fmaxf(1e-10f, qd);
              }
            }
            ++r;
          }
          if (!more) break;
          rw = rw_n; zi = zi_n; zn = zn_n; tf = tf_n; i = i_n; p += 32;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_raw + 16 + 8 * tile);              // the slot may be overwritten
      TRACE(tr, 0xA00 + tile);                                          // ... composited
    }
  }
}

template <bool kStash, bool kComp = false>
__global__ void __launch_bounds__(PP_THREADS, 1) nerf_mlp_tc_pp_kernel(const __grid_constant__ FwdArgs a,
                                                                       const __grid_constant__ NetPlan plan,
                                                                       const __grid_constant__ CUtensorMap tmap,
                                                                       const __grid_constant__ StashArgs sa) {
  extern __shared__ __align__(1024) uint8_t smem_fwd[];
  uint8_t* smem = smem_fwd;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();                  // SWIZZLE_128B operand tiles need 1024-byte alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int64_t unit0 = cluster_id_x(), n_units = num_clusters_x();
  const int64_t n_steps = (a.n_pairs + 1) / 2;

  const uint32_t bar_full = sbase + OFF_BAR, bar_empty = bar_full + 8 * NUM_STAGES;
  const uint32_t bar_acc = bar_empty + 16 * NUM_STAGES, bar_aready = bar_acc + 16;
  const uint32_t bar_raw = bar_empty + 8 * NUM_STAGES;   // kComp: [tile] raw values parked, [2 + tile] slot read (4 x 8 B, else unused)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (3 * NUM_STAGES + 4));

  if (threadIdx.x == 0) {
    for (int s = 0; s < NUM_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_acc + 8 * t, 1);
      mbar_init(bar_aready + 8 * t, PP_EPI_WARPS);      // 8 local + 8 remote epilogue warps per super-tile
      if (kComp) {
        mbar_init(bar_raw + 8 * t, 4);                  // the tile's four column-half-0 warps
        mbar_init(bar_raw + 16 + 8 * t, 1);             // the compositor
      }
    }
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
      if (cta_rank == 0) pp_mma_issuer(plan, sbase, tmem_base, bar_full, bar_empty, bar_acc, bar_aready, unit0, n_steps, n_units, a.dbg);
    } else if (kComp && warp == 2) {
#if SCADE_TC_TRACE
      if (!(a.dbg & 128))
#endif
      pp_compositor(a, bar_raw, cta_rank, unit0, n_steps, n_units, lane);
    }
  } else {
    // ================= prologue / epilogue warps: thread == one row x 128 columns =================
    regs_epilogue();
    const int ew = warp - PP_EPI_WARP0;
    const int tile = ew >> 3;
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (ew >> 2) & 1;                     // which 128 accumulator columns
    const int row = quarter * 32 + lane;
    uint8_t* a_tile = smem + OFF_A + tile * 4 * CHUNK_BYTES;
    uint8_t* emb_tile = smem + OFF_EMB + tile * CHUNK_BYTES;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + tile * W;
    const PackedTail* tail = reinterpret_cast<const PackedTail*>(a.packed + (size_t)plan.stages_per_pass * STAGE_BYTES);
    const uint32_t my_acc = bar_acc + 8 * tile, my_aready = bar_aready + 8 * tile;
    const uint32_t aready_target = cta_rank != 0 ? map_to_cta(my_aready, 0) : my_aready;
    const int pair_bar = 1 + tile * 4 + quarter;         // named barrier shared by the two column-half warps of this row quarter
    const int tile_bar = 9 + tile;                       // named barrier of the tile's 8 epilogue warps
    const int tid_tile = half * 128 + row;               // 0..255 within the tile's epilogue threads
    float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS) + tile * W;
    const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128), rx4 = (uint32_t)((row & 7) << 4);
    uint32_t acc_phase = 0;
    [[maybe_unused]] Tracer tr;
    auto signal_a_ready = [&]() {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank != 0) mbar_arrive_remote_relaxed(aready_target);
        else mbar_arrive(my_aready);
      }
    };
    // global point index of this thread's row in step `step`
    // (`it` = how many steps this cluster has done before `step`)
    auto point_of = [&](int64_t step, int it) {
      if (kComp && a.comp.chain_iters > 0)               // chain mode: this CTA's own contiguous range, one step after the other
        return ((2 * unit0 + cta_rank) * (int64_t)a.comp.chain_iters + it) * (int64_t)(TILES * TILE_M) + tile * TILE_M + row;
      return (2 * step + cta_rank) * (int64_t)(TILES * TILE_M) + tile * TILE_M + row;
    };

    // ---- positional encoding of this thread's 32 encoding-chunk columns [32*half, 32*half + 32) for step `step`, as 16 packed
    //      fp16 pairs.  Computed one layer early (while the views layer's MMAs run) and stored once those MMAs have retired.
    auto encode = [&](int64_t step, int it, uint32_t (&pk)[16], float (&vd)[3]) {
      const int64_t p_raw = point_of(step, it);
      const int64_t p = p_raw < a.P ? p_raw : a.P - 1;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
      if (a.rays != nullptr) {
        const int64_t r = a.P < (int64_t)0x7fffffff ? (int64_t)((uint32_t)p / (uint32_t)a.S) : p / a.S;
        const float* ray = a.rays + r * a.ray_stride;
        const float zz = a.z[p];
        const float cen[3] = {a.cx, a.cy, a.cz};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          float pt = __fadd_rn(ray[d], __fmul_rn(ray[3 + d], zz));            // RS:657
          float x = __fmul_rn(__fsub_rn(pt, cen[d]), a.bb_scale);             // RS:52
          float xp = __fmul_rn(x, 3.14159274101257324f);                      // H:165
          if (half == 0) {
            v[d] = x;                                                         // columns 0..2
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              if (k < a.multires) {
                float s, c;
                sincos_reduced(__fmul_rn(xp, (float)(1 << k)), &s, &c);
                if (3 + 6 * k + d < 32) v[3 + 6 * k + d] = s;                 // sin block of octave k
                if (6 + 6 * k + d < 32) v[6 + 6 * k + d] = c;                 // cos block
              }
            }
          } else {
#pragma unroll
            for (int k = 4; k < 9; ++k) {
              if (k < a.multires) {
                float s, c;
                sincos_reduced(__fmul_rn(xp, (float)(1 << k)), &s, &c);
                if (3 + 6 * k + d >= 32) v[3 + 6 * k + d - 32] = s;
                if (6 + 6 * k + d >= 32) v[6 + 6 * k + d - 32] = c;
              }
            }
          }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) vd[d] = ray[8 + d];
      } else {
        vd[0] = vd[1] = vd[2] = 0.f;
        const float* x = a.x_embedded + p * a.in_all;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = 32 * half + i;
          if (col < a.in_all) v[i] = x[col];
        }
      }
      if (half == 1) {
        v[ONES_COL - 32] = 1.0f;
        v[ONES_COL + 1 - 32] = 1.0f;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
    };
    auto store_emb = [&](const uint32_t (&pk)[16], const float (&vd)[3]) {
#pragma unroll
      for (int pc = 0; pc < 4; ++pc)
        *reinterpret_cast<uint4*>(emb_tile + sw128_offset(row, half * 4 + pc)) = make_uint4(pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]);
      if (a.rays != nullptr) {
        // view direction (multires_views == 0): columns in_ch .. in_ch+2, written by the thread that owns them
        const int in_ch = 3 + 6 * a.multires;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int col = in_ch + d;
          if ((col >> 5) == half)
            *reinterpret_cast<__half*>(emb_tile + sw128_offset(row, col >> 3) + (col & 7) * 2) = __float2half_rn(vd[d]);
        }
      }
    };

    if (unit0 < n_steps) {
      TRACE(tr, 0x400);                                  // epilogue warp: first prologue begins
      uint32_t pk[16];
      float vd[3];
      encode(unit0, 0, pk, vd);
      store_emb(pk, vd);
      signal_a_ready();
      TRACE(tr, 0x401);                                  // prologue done, operand signalled
    }

    int it = 0;
    for (int64_t step = unit0; step < n_steps; step += n_units, ++it) {
      const int64_t p_raw = point_of(step, it);
      const bool live = p_raw < a.P;
      const int64_t tile_g = 2 * (2 * step + cta_rank) + tile;            // 128-point tile index in the stash

      float alpha = 0.f;                                  // this thread's share of alpha_linear (its 128 columns)
      for (int l = 0; l < plan.n_layers; ++l) {
        const int kind = plan.layers[l].kind, relu = plan.layers[l].relu, bias_epi = plan.layers[l].bias_epi;
        if (kind != 3) {
          // fetched while the layer's MMAs run: this thread's element of the bias row that goes through shared memory
          float s_mine = 0.f;
          float4 wa = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias_epi) s_mine = __ldg(tail->bias[l] + tid_tile);
          if (kind == 1) wa = __ldg(reinterpret_cast<const float4*>(tail->w_alpha + half * 128) + lane);
          mbar_wait(my_acc, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
          TRACE(tr, 0x500 + l);                            // accumulators of layer l visible
#if SCADE_TC_TRACE
          if (a.dbg & 32) {                                // timing ablation: no epilogue work at all
            signal_a_ready();
            TRACE(tr, 0x700 + l);
            continue;
          }
#endif
          // hidden / feature layer: this thread's 128 accumulator columns -> A chunks 2*half, 2*half+1
          uint32_t sgn[4];
          const uint32_t t_col = t_lane + half * 128;
          if (kStash) {
            if (l == 0 && half == 0 && lane == 0) {       // the encoding chunk of this tile is complete: stash this warp's 32 rows
              bulk_s2g(sa.ws + sa.L.emb + (size_t)tile_g * CHUNK_BYTES + quarter * 4096, smem_u32(emb_tile) + quarter * 4096, 4096);
              bulk_commit();
            }
            if (lane == 0) bulk_wait_read0();             // earlier stash stores no longer read the rows overwritten below
            __syncwarp();
          }
          if (bias_epi) {
            // the tile's row goes through shared memory (broadcast LDS.128 in the epilogue).  No thread of the tile can still
            // be reading the previous layer's row: every warp signalled its operand before these MMAs were issued.
            sbias[tid_tile] = s_mine;
            named_bar_sync(tile_bar, 256);
          }
          uint8_t* dst = a_tile + 2 * half * CHUNK_BYTES + row_off;
          const float* srow = sbias + half * 128;
          if (kind == 1) {
            if (bias_epi) alpha = hidden_epilogue<true, true, true, kStash>(t_col, dst, rx4, srow, wa, sgn);
            else alpha = hidden_epilogue<false, true, true, kStash>(t_col, dst, rx4, srow, wa, sgn);
          } else if (kind == 2) {
            if (!kStash && plan.layers[l].out_tmem) feature_epilogue_tmem(t_lane, half, sbias, pair_bar);
            else hidden_epilogue<true, false, false, false>(t_col, dst, rx4, srow, wa, sgn);
          }
          else if (bias_epi) hidden_epilogue<true, true, false, kStash>(t_col, dst, rx4, srow, wa, sgn);
          else hidden_epilogue<false, true, false, kStash>(t_col, dst, rx4, srow, wa, sgn);
          TRACE(tr, 0x600 + l);                          // operand chunks written
          signal_a_ready();
          TRACE(tr, 0x700 + l);                          // signalled
          if (kStash) {
            if (relu)
              *reinterpret_cast<uint4*>(sa.ws + sa.L.maskh[l] + ((size_t)(tile_g * 2 + half) * TILE_M + row) * 16) =
                  make_uint4(sgn[0], sgn[1], sgn[2], sgn[3]);
            if (lane == 0) {
              uint8_t* dst = sa.ws + (kind == 2 ? sa.L.feat : sa.L.h[l]) + (size_t)(tile_g * 4 + 2 * half) * CHUNK_BYTES + quarter * 4096;
              const uint32_t src = smem_u32(a_tile) + 2 * half * CHUNK_BYTES + quarter * 4096;
              bulk_s2g(dst, src, 4096);
              bulk_s2g(dst + CHUNK_BYTES, src + CHUNK_BYTES, 4096);
              bulk_commit();
            }
          }
        } else {
          // views layer (N = 128) + rgb_linear + output (H:238-242).  The two column-half threads of a row take 64 hidden
          // columns each; half 1 hands its partial sums (and its share of alpha) to half 0 through shared memory.
          // The next step's encoding is computed BEFORE waiting for this layer's MMAs and stored right after them.
          const int64_t next = step + n_units;
          const bool has_next = next < n_steps;
          uint32_t pk[16];
          float vd[3];
          if (has_next) encode(next, it + 1, pk, vd);
          // rgb_linear's weights (3 x 128 fp32) are staged by one warp of the tile in activation chunk 3, which is dead once
          // this layer's MMAs have retired
          const bool stager = (half == 1 && quarter == 0);
          float4 wst[3];
          if (stager) {
#pragma unroll
            for (int i = 0; i < 3; ++i) wst[i] = __ldg(reinterpret_cast<const float4*>(&tail->w_rgb_p[0][0]) + i * 32 + lane);
          }
          // kComp, last step (see pp_compositor): this sample's z, its successor's and the ray's direction norm, fetched under
          // the MMAs (RS:514-516)
          float c_z = 0.f, c_zn = 0.f, c_norm = 0.f;
          bool c_last = false;
          if (kComp && half == 0 && !has_next && a.comp.tail_mode) {
            const int64_t pc = live ? p_raw : 0;
            const int64_t rc = a.P < (int64_t)0x7fffffff ? (int64_t)((uint32_t)pc / (uint32_t)a.S) : pc / a.S;
            const int si = (int)(pc - rc * a.S);
            c_last = si == a.S - 1;
            c_z = a.z[pc];
            c_zn = c_last ? c_z : a.z[pc + 1];
            const float* rd = a.rays + rc * a.ray_stride + 3;
            const float dx = rd[0], dy = rd[1], dz = rd[2];
            c_norm = sqrtf(dx * dx + dy * dy + dz * dz);
          }
          mbar_wait(my_acc, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
          TRACE(tr, 0x500 + l);
#if SCADE_TC_TRACE
          if (a.dbg & 32) {
            tc_fence_before();
            if (has_next) { store_emb(pk, vd); signal_a_ready(); }
            TRACE(tr, 0x700 + l);
            continue;
          }
#endif
          // 1. this thread's 64 accumulator columns -> registers; the tile's TMEM and encoding chunk are then free, so the
          //    next step's layer 0 is released BEFORE the rgb arithmetic (which then runs under that layer's MMAs)
          uint32_t rr[2][32];
          const uint32_t t_views = t_lane + plan.layers[l].d_col + half * 64;
          tmem_ld32(t_views, rr[0]);
          tmem_ld32(t_views + 32, rr[1]);
          tmem_ld_wait();
          tc_fence_before();
          if (has_next) {
            store_emb(pk, vd);
            signal_a_ready();
          }
          TRACE(tr, 0x600 + l);
          // 2. rgb_linear weights -> shared memory (chunk 3 is dead: this layer's MMAs have retired)
          float cr[2] = {0.f, 0.f}, cg[2] = {0.f, 0.f}, cb[2] = {0.f, 0.f};
          uint32_t sgn[2];
          uint8_t* hv_chunk = a_tile + 2 * half * CHUNK_BYTES;      // stash staging of this half's 64 h_v columns (own chunk)
          float4* scratch = reinterpret_cast<float4*>(a_tile + 3 * CHUNK_BYTES + quarter * 4096) + lane;   // half 1's own, dead chunk
          if (kStash) {
            if (lane == 0) bulk_wait_read0();              // this warp's feature-chunk stash stores have read their source
            __syncwarp();
          }
          float* srgb = reinterpret_cast<float*>(a_tile + 3 * CHUNK_BYTES + 1024);     // [3][128], clear of the scratch rows
          if (stager) {
#pragma unroll
            for (int i = 0; i < 3; ++i) reinterpret_cast<float4*>(srgb)[i * 32 + lane] = wst[i];
          }
          named_bar_sync(tile_bar, 256);
          const float4* wr4 = reinterpret_cast<const float4*>(srgb + half * 64);
          const float4* wg4 = reinterpret_cast<const float4*>(srgb + 128 + half * 64);
          const float4* wb4 = reinterpret_cast<const float4*>(srgb + 256 + half * 64);
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const uint32_t* r = rr[c2];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 wr = wr4[c2 * 8 + q], wg = wg4[c2 * 8 + q], wb = wb4[c2 * 8 + q];
              const float h0 = fmaxf(__uint_as_float(r[4 * q + 0]), 0.f), h1 = fmaxf(__uint_as_float(r[4 * q + 1]), 0.f);
              const float h2 = fmaxf(__uint_as_float(r[4 * q + 2]), 0.f), h3 = fmaxf(__uint_as_float(r[4 * q + 3]), 0.f);
              cr[0] = fmaf(h0, wr.x, cr[0]); cg[0] = fmaf(h0, wg.x, cg[0]); cb[0] = fmaf(h0, wb.x, cb[0]);
              cr[1] = fmaf(h1, wr.y, cr[1]); cg[1] = fmaf(h1, wg.y, cg[1]); cb[1] = fmaf(h1, wb.y, cb[1]);
              cr[0] = fmaf(h2, wr.z, cr[0]); cg[0] = fmaf(h2, wg.z, cg[0]); cb[0] = fmaf(h2, wb.z, cb[0]);
              cr[1] = fmaf(h3, wr.w, cr[1]); cg[1] = fmaf(h3, wg.w, cg[1]); cb[1] = fmaf(h3, wb.w, cb[1]);
            }
            if (kStash) {
              sgn[c2] = sign_mask32(r);
#pragma unroll
              for (int pc = 0; pc < 4; ++pc) {
                uint4 q;
                q.x = pack_relu_f16x2(r[pc * 8 + 0], r[pc * 8 + 1]);
                q.y = pack_relu_f16x2(r[pc * 8 + 2], r[pc * 8 + 3]);
                q.z = pack_relu_f16x2(r[pc * 8 + 4], r[pc * 8 + 5]);
                q.w = pack_relu_f16x2(r[pc * 8 + 6], r[pc * 8 + 7]);
                *reinterpret_cast<uint4*>(hv_chunk + sw128_offset(row, c2 * 4 + pc)) = q;
              }
            }
          }
          // 3. half 1 hands its partial sums (and its share of alpha) to half 0
          const float pr = cr[0] + cr[1], pg = cg[0] + cg[1], pb = cb[0] + cb[1];
          if (half == 1) {
            *scratch = make_float4(pr, pg, pb, alpha);
            __threadfence_block();
            named_bar_arrive(pair_bar, 64);
          } else {
            named_bar_sync(pair_bar, 64);
            const float4 o = *scratch;
            const float al = (alpha + o.w) + __ldg(&tail->b_alpha);
            const float4 rawv = make_float4((pr + o.x) + __ldg(&tail->b_rgb[0]), (pg + o.y) + __ldg(&tail->b_rgb[1]),
                                            (pb + o.z) + __ldg(&tail->b_rgb[2]), softplus_beta10(al));
            if (live && (!kComp || a.comp.write_raw)) a.out[p_raw] = rawv;
            if (kStash) reinterpret_cast<float*>(sa.ws + sa.L.alpha)[tile_g * TILE_M + row] = al;
#if SCADE_TC_TRACE
            if (kComp && !(a.dbg & 128)) {
#else
            if (kComp) {
#endif
              // hand the point to the compositor warp: slot free (it has read the previous step's values) -> store -> signal
              if (it > 0) mbar_wait(bar_raw + 16 + 8 * tile, (uint32_t)(it - 1) & 1u);
              uint8_t* slot = a.comp.ring + (2 * unit0 + cta_rank) * (int64_t)COMP_RING_BYTES_PER_CTA;
              float4 park = rawv;
              if (!has_next && a.comp.tail_mode) {          // last step: the per-sample terms, evaluated here (RS:512-520, 543)
                const SampleTerms t = sample_terms(rawv.w, 0.f, c_z, c_zn, c_last, c_norm);
                park = make_float4(sigmoidf_(rawv.x), sigmoidf_(rawv.y), sigmoidf_(rawv.z), t.alpha);
                reinterpret_cast<float*>(slot + TILES * TILE_M * 16)[tile * TILE_M + row] = t.tfac;
              }
              reinterpret_cast<float4*>(slot)[tile * TILE_M + row] = park;
              __syncwarp();
#if SCADE_TC_TRACE
              if (a.dbg & 256) {                            // timing ablation: arrive without the release
                if (lane == 0) asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar_raw + 8 * tile) : "memory");
              } else
#endif
              if (lane == 0) mbar_arrive(bar_raw + 8 * tile);
            }
          }
          if (kStash) {
            *reinterpret_cast<uint2*>(sa.ws + sa.L.maskv + (size_t)(tile_g * TILE_M + row) * 16 + half * 8) = make_uint2(sgn[0], sgn[1]);
            fence_proxy_async();                           // the h_v staging rows were written through the generic proxy
            __syncwarp();
            if (lane == 0) {
              bulk_s2g(sa.ws + sa.L.hv + (size_t)(tile_g * 2 + half) * CHUNK_BYTES + quarter * 4096, smem_u32(hv_chunk) + quarter * 4096, 4096);
              bulk_commit();
            }
          }
          // 4. every warp of the tile is done with the staged weights / scratch rows before the next step's layer-0 epilogue
          //    (which overwrites chunk 3) can begin
          named_bar_sync(tile_bar, 256);
          TRACE(tr, 0x700 + l);                          // views epilogue done
        }
      }
    }
  }

  if (kStash && warp >= PP_EPI_WARP0 && lane == 0) bulk_wait_all0();     // stash stores are complete before the CTA retires
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---- weight packing ---------------------------------------------------------------------------------
__device__ __forceinline__ void split_f16(float b, __half* hi, __half* lo) {
  *hi = __float2half_rn(b);
  *lo = __float2half_rn(b - __half2float(*hi));
}

// Source of element (n, k) of a stage: at most one fp32 value (a weight, or a bias half riding on the 1.0 columns).
// mode 0: zero; 1: weight -> hi (or lo for a W_lo stage); 2: bias -> hi half; 3: bias -> lo half (both zero in a W_lo stage).
__device__ __forceinline__ const float* pack_source(const StageDesc& sd, int n, int k, int* mode) {
  *mode = 0;
  const float* p = nullptr;
  for (int j = 0; j < sd.n_blocks; ++j) {
    const StageBlock& b = sd.blk[j];
    const int bn = n - b.dst_row0;
    if (bn < 0 || bn >= b.nrows) continue;
    if (b.bias_mode == 2) {
      // bias stage: K-step 0 only; positions 12/13 meet the encoding chunk's 1.0 columns (60/61 = 48 + 12/13)
      if (k == 12 || k == 13) { p = sd.bias + sd.row0 + bn; *mode = (k == 12) ? 2 : 3; }
      continue;
    }
    const int sc = k - b.dst_col0;
    if (sc >= 0 && sc < b.ncols) {
      p = sd.trans ? sd.W + (int64_t)(b.col0 + sc) * sd.ld + sd.row0 + bn : sd.W + (int64_t)(sd.row0 + bn) * sd.ld + b.col0 + sc;
      *mode = 1;
    }
    if (b.bias_mode == 1 && (k == ONES_COL || k == ONES_COL + 1)) {      // both bias halves ride in the W_hi stage
      p = sd.bias + sd.row0 + bn;
      *mode = (k == ONES_COL) ? 2 : 3;
    }
  }
  return p;
}

// One launch packs a network's whole stream: blocks [0, n_fwd) the forward stages, block n_fwd the fp32 tail, the rest the
// dgrad stages behind it (at bwd_off).  A block is one 16 KB stage; a thread's 32 elements are fetched 8 at a time with
// independent loads (the re-pack sits on the train step's critical path after every optimizer step: ~25 us -> a few us).
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackPlan plan, uint8_t* __restrict__ out) {
  if ((int)blockIdx.x == plan.n_fwd) {
    // fp32 tail: rgb_linear (two layouts), alpha_linear, head biases, epilogue biases
    if (plan.w_alpha == nullptr) return;
    PackedTail* tail = reinterpret_cast<PackedTail*>(out + (size_t)plan.n_fwd * STAGE_BYTES);
    for (int k = threadIdx.x; k < W / 2; k += blockDim.x) {
      tail->w_rgb[k] = make_float4(plan.w_rgb[k], plan.w_rgb[W / 2 + k], plan.w_rgb[W + k], 0.f);
      tail->w_rgb_p[0][k] = plan.w_rgb[k];
      tail->w_rgb_p[1][k] = plan.w_rgb[W / 2 + k];
      tail->w_rgb_p[2][k] = plan.w_rgb[W + k];
    }
    for (int k = threadIdx.x; k < W; k += blockDim.x) tail->w_alpha[k] = plan.w_alpha[k];
    for (int l = 0; l < MAX_LAYERS; ++l)
      for (int k = threadIdx.x; k < W; k += blockDim.x) tail->bias[l][k] = plan.epi_bias[l] ? plan.epi_bias[l][k] : 0.f;
    if (threadIdx.x == 0) {
      tail->b_alpha = plan.b_alpha[0];
      tail->b_rgb[0] = plan.b_rgb[0]; tail->b_rgb[1] = plan.b_rgb[1]; tail->b_rgb[2] = plan.b_rgb[2];
    }
    return;
  }
  const bool bwd = (int)blockIdx.x > plan.n_fwd;
  const StageDesc& sd = plan.st[bwd ? blockIdx.x - 1 : blockIdx.x];
  __half* dst = reinterpret_cast<__half*>(bwd ? out + plan.bwd_off + (size_t)(blockIdx.x - plan.n_fwd - 1) * STAGE_BYTES
                                              : out + (size_t)blockIdx.x * STAGE_BYTES);
  constexpr int PER_THREAD = STAGE_N * KCHUNK / 256, BATCH = 8;
#pragma unroll 1
  for (int it0 = 0; it0 < PER_THREAD; it0 += BATCH) {
    float v[BATCH];
    int mode[BATCH], nn[BATCH], kk[BATCH];
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int idx = threadIdx.x + (it0 + i) * 256;
      // consecutive threads walk the source's contiguous axis: k for (out, in) weights, n for the transposed dgrad stages
      nn[i] = sd.trans ? idx % STAGE_N : idx / KCHUNK;
      kk[i] = sd.trans ? idx / STAGE_N : idx % KCHUNK;
      const float* p = pack_source(sd, nn[i], kk[i], &mode[i]);
      v[i] = p != nullptr ? __ldg(p) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      __half hi, lo;
      split_f16(v[i], &hi, &lo);
      __half hv = __float2half_rn(0.f);
      if (mode[i] == 1) hv = sd.lo ? lo : hi;
      else if (mode[i] == 2) hv = sd.lo ? hv : hi;
      else if (mode[i] == 3) hv = sd.lo ? hv : lo;
      const uint32_t off = sw128_offset(nn[i], kk[i] >> 3) + (kk[i] & 7) * 2;
      dst[off >> 1] = hv;
    }
  }
}

// Build the layer table and the stage list for a network description.  Stage order = consumption order:
// layer -> ring stage -> CTA (CTA 0's N half, CTA 1's N half).  A ring stage holds `kpack` K chunks of one CTA's N half.
static void build_plans(const scade_net& net, NetPlan* np, PackPlan* pp, bool x3 = false, bool views_tmem = false) {
  const scade_net_desc& d = net.desc;
  NetDims nd(d);
  NetPlan P{};
  PackPlan Q{};
  // K chunk kc of a layer: the encoding chunk (columns emb_col0.. of the weight matrix land at stage columns emb_dst..,
  // bias on the 1.0 columns) or activation chunk c (weight columns h_col0 + 64 c ..)
  auto add_layer = [&](const float* Wt, int fan_in, bool with_emb, int emb_col0, int emb_ncols, int emb_dst, int h_col0,
                       bool with_h, int n_out, int relu, int kind, const float* bias) {
    LayerDesc L{};
    const int li = P.n_layers;
    P.first_stage[li] = Q.n_stages;
    L.relu = relu; L.kind = kind; L.n_out = n_out; L.a_step = 1;
    L.bias_stage = 0;                                          // (kept in the format: a K=16 bias MMA step for layers that want it)
    L.bias_epi = (with_emb || L.bias_stage) ? 0 : 1;
    Q.epi_bias[li] = L.bias_epi ? bias : nullptr;
    int nk = 0;
    if (with_emb) L.a_src[nk++] = SRC_EMB;
    if (with_h) for (int c = 0; c < 4; ++c) L.a_src[nk++] = c;
    L.n_k = nk;
    const int rows = n_out / 2;                                // N rows per CTA
    L.kpack = STAGE_N / rows;                                  // 1 (N = 256) or 2 (N = 128)
    L.emb_ks0 = with_emb ? emb_dst / 16 : 0;                   // K steps below the first mapped column hold only zero weights
    const int n_own = (nk + L.kpack - 1) / L.kpack;
    for (int i = 0; i < n_own; ++i)
      for (int rep = 0; rep < (x3 ? 2 : 1); ++rep)              // x3: (W_hi stage, W_lo stage) per ring stage
      for (int h = 0; h < 2; ++h) {
        StageDesc s{};
        s.W = Wt; s.ld = fan_in; s.row0 = h * rows; s.bias = bias; s.trans = 0; s.lo = rep;
        for (int j = 0; j < L.kpack; ++j) {
          const int kc = i * L.kpack + j;
          if (kc >= nk) break;
          StageBlock& b = s.blk[s.n_blocks++];
          b.dst_row0 = j * rows; b.nrows = rows;
          if (L.a_src[kc] == SRC_EMB) { b.col0 = emb_col0; b.ncols = emb_ncols; b.dst_col0 = emb_dst; b.bias_mode = 1; }
          else { b.col0 = h_col0 + 64 * L.a_src[kc]; b.ncols = 64; b.dst_col0 = 0; b.bias_mode = 0; }
        }
        Q.st[Q.n_stages++] = s;
      }
    if (L.bias_stage)
      for (int h = 0; h < 2; ++h) {
        StageDesc s{};
        s.W = nullptr; s.ld = 0; s.row0 = h * rows; s.bias = bias; s.trans = 0; s.n_blocks = 1;
        s.blk[0].dst_row0 = 0; s.blk[0].nrows = rows; s.blk[0].bias_mode = 2;
        Q.st[Q.n_stages++] = s;
      }
    P.n_stages[li] = Q.n_stages - P.first_stage[li];
    P.layers[P.n_layers++] = L;
  };
  for (int i = 0; i < d.D; ++i) {
    const float* Wt = net.params[2 * i];
    const float* b = net.params[2 * i + 1];
    int kind = (i == d.D - 1) ? 1 : 0;
    if (i == 0) add_layer(Wt, nd.in_ch, true, 0, nd.in_ch, 0, 0, false, W, 1, kind, b);
    else if (i - 1 == d.skip) add_layer(Wt, nd.in_ch + W, true, 0, nd.in_ch, 0, nd.in_ch, true, W, 1, kind, b);
    else add_layer(Wt, W, false, 0, 0, 0, 0, true, W, 1, kind, b);
  }
  const int pv = 2 * d.D;
  add_layer(net.params[pv + 2], W, false, 0, 0, 0, 0, true, W, 0, 2, net.params[pv + 3]);                        // feature_linear
  add_layer(net.params[pv], W + nd.in_views, true, W, nd.in_views, nd.in_ch, 0, true, W / 2, 1, 3, net.params[pv + 1]);   // views
  if (views_tmem) {                                          // feature layer -> TMEM, views layer reads it from there
    P.layers[P.n_layers - 2].out_tmem = 1;
    P.layers[P.n_layers - 1].a_tmem = 1;
    P.layers[P.n_layers - 1].d_col = W / 2;
  }
  P.stages_per_pass = Q.n_stages;
  Q.n_fwd = Q.n_stages;
  Q.w_alpha = net.params[pv + 4]; Q.b_alpha = net.params[pv + 5];
  Q.w_rgb = net.params[pv + 6]; Q.b_rgb = net.params[pv + 7];
  if (np) *np = P;
  if (pp) *pp = Q;
}

static int count_stages(const scade_net_desc& d) {
  int n = 2;                                   // layer 0: the encoding chunk x 2 CTAs (bias inside)
  for (int i = 1; i < d.D; ++i) n += (i - 1 == d.skip) ? 10 : 8;      // 4 activation chunks (+ encoding chunk after the skip) x 2 CTAs
  n += 8;                                      // feature_linear
  n += 6;                                      // views (N = 128): 5 K chunks, two per stage, x 2 CTAs
  return n;
}


#include "mlp_tc_train.cuh"
#include "mlp_tc_x3.cuh"

// ---- training: backward stream, stash layout ---------------------------------------------------------------------------
// Backward weight stream (dgrad): layer j of the chain multiplies by W^T, so stage (K chunk kc, N half h) holds
// B[n][k] = W[64 kc + k][col0 + 128 h + n].  Order: views, feature, pts_{D-1} .. pts_1 (pts_0 needs no input gradient).
static int count_bwd_stages(const scade_net_desc& d) { return 4 + 8 + 8 * (d.D - 1); }

static void build_bwd_plans(const scade_net& net, NetPlan* np, PackPlan* pp) {
  const scade_net_desc& d = net.desc;
  NetDims nd(d);
  NetPlan P{};
  PackPlan Q{};
  auto add_layer = [&](const float* Wt, int ld, int n_out_fwd, int col0, int a_step) {
    LayerDesc L{};
    P.first_stage[P.n_layers] = Q.n_stages;
    L.relu = 0; L.kind = 0; L.n_out = W; L.a_step = a_step; L.bias_epi = 0; L.kpack = 1; L.bias_stage = 0; L.emb_ks0 = 0;
    L.n_k = n_out_fwd / KCHUNK;
    for (int kc = 0; kc < L.n_k; ++kc) {
      L.a_src[kc] = kc * a_step;
      for (int h = 0; h < 2; ++h) {
        StageDesc sd{};
        sd.W = Wt; sd.ld = ld; sd.row0 = col0 + STAGE_N * h; sd.bias = nullptr; sd.trans = 1; sd.n_blocks = 1;
        sd.blk[0].col0 = 64 * kc; sd.blk[0].ncols = 64; sd.blk[0].dst_col0 = 0; sd.blk[0].dst_row0 = 0; sd.blk[0].nrows = STAGE_N;
        Q.st[Q.n_stages++] = sd;
      }
    }
    P.n_stages[P.n_layers] = Q.n_stages - P.first_stage[P.n_layers];
    P.layers[P.n_layers++] = L;
  };
  const int pv = 2 * d.D;
  add_layer(net.params[pv], W + nd.in_views, W / 2, 0, 2);          // views_linears.0^T (feature columns)        H:235-238
  add_layer(net.params[pv + 2], W, W, 0, 1);                        // feature_linear^T                           H:234
  for (int i = d.D - 1; i >= 1; --i) {
    const bool after_skip = (i - 1 == d.skip);                      // input = [input_pts, h]: h starts at column in_ch (H:230)
    add_layer(net.params[2 * i], after_skip ? nd.in_ch + W : W, W, after_skip ? nd.in_ch : 0, 1);
  }
  P.stages_per_pass = Q.n_stages;
  if (np) *np = P;
  if (pp) *pp = Q;
}

static size_t fwd_stream_bytes(const scade_net_desc& d, bool x3 = false) {
  return (size_t)count_stages(d) * (x3 ? 2 : 1) * STAGE_BYTES + align_up(sizeof(PackedTail), 1024);
}

static TrainLayout train_layout(const scade_net_desc& d, int64_t P) {
  TrainLayout L{};
  const int64_t n_pairs = ceil_div<int64_t>(P, TILES * TILE_M);
  L.T = 4 * ((n_pairs + 1) / 2);
  L.D = d.D;
  unsigned long long off = 0;
  auto chunks = [&](int per_tile) { unsigned long long o = off; off += (unsigned long long)L.T * per_tile * CHUNK_BYTES; return o; };
  L.emb = chunks(1);
  for (int l = 0; l < d.D; ++l) L.h[l] = chunks(4);
  L.feat = chunks(4);
  L.hv = chunks(2);
  L.dzv = chunks(2);
  L.dzf = chunks(4);
  for (int l = 0; l < d.D; ++l) L.dz[l] = chunks(4);
  for (int l = 0; l < d.D; ++l) { L.maskh[l] = off; off += (unsigned long long)L.T * TILE_M * 32; }
  L.maskv = off; off += (unsigned long long)L.T * TILE_M * 16;
  L.alpha = off; off += (unsigned long long)L.T * TILE_M * 4;
  L.gs = off; off += 256;
  L.total = off;
  return L;
}

// the [rows x 128 B] view of a buffer of swizzled 128-byte rows, loaded in boxes of `box_rows` rows
static int encode_rows_tmap(CUtensorMap* tmap, const void* base, size_t bytes, int box_rows) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SCADE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return SCADE_ERR_CUDA;
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {128, (cuuint64_t)(bytes / 128)};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SCADE_ERR_CUDA;
  }
  return SCADE_OK;
}

// cudaFuncSetAttribute is per device: remember which devices of this process have been configured
static int set_kernel_attributes() {
  static bool done[64] = {};
  int dev = 0;
  SCADE_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && done[dev]) return SCADE_OK;
  SCADE_CUDA(cudaFuncSetAttribute(nerf_mlp_tc_pp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES));
  SCADE_CUDA(cudaFuncSetAttribute(nerf_mlp_tc_pp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES));
  SCADE_CUDA(cudaFuncSetAttribute(nerf_mlp_tc_pp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES));
  SCADE_CUDA(cudaFuncSetAttribute(nerf_mlp_tc_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES));
  SCADE_CUDA(cudaFuncSetAttribute(nerf_mlp_tc_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  SCADE_CUDA(cudaFuncSetAttribute(nerf_mlp_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
  if (dev >= 0 && dev < 64) done[dev] = true;
  return SCADE_OK;
}

static int launch_pair(const void* kern, int clusters, int threads, int smem, cudaStream_t st, void** args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  SCADE_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
  return SCADE_OK;
}

}  // namespace tc

bool mlp_tc_supported(const scade_net_desc& d) {
  NetDims nd(d);
  return d.W == tc::W && d.D >= 2 && d.D <= 8 && nd.in_all <= tc::ONES_COL && d.multires <= 9 && d.multires_views == 0 &&
         d.skip != d.D - 1 && 2 * tc::count_stages(d) <= tc::MAX_STAGE_DESCS;
}

size_t mlp_tc_packed_bytes(const scade_net_desc& d, bool x3) {
  if (x3) return tc::fwd_stream_bytes(d, true);                                              // (W_hi, W_lo) forward stream + tail
  return tc::fwd_stream_bytes(d) + (size_t)tc::count_bwd_stages(d) * tc::STAGE_BYTES;      // forward stream + tail | dgrad stream
}

int mlp_tc_pack(const scade_net& net, void* packed_out, cudaStream_t st, bool x3) {
  tc::PackPlan pp;
  tc::build_plans(net, nullptr, &pp, x3);
  if (!x3) {
    // the dgrad stream rides in the same launch, behind the forward stream and its tail
    tc::PackPlan pb;
    tc::build_bwd_plans(net, nullptr, &pb);
    if (pp.n_stages + pb.n_stages > tc::MAX_STAGE_DESCS) {
      set_error("mlp_pack: %d stages exceed the pack plan", pp.n_stages + pb.n_stages);
      return SCADE_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < pb.n_stages; ++i) pp.st[pp.n_stages + i] = pb.st[i];
    pp.n_stages += pb.n_stages;
    pp.bwd_off = tc::fwd_stream_bytes(net.desc);
  }
  tc::pack_kernel<<<pp.n_stages + 1, 256, 0, st>>>(pp, reinterpret_cast<uint8_t*>(packed_out));
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

MlpCompositePlan mlp_tc_composite_plan(int S, int64_t P, int n_sms) {
  const int64_t cta_steps = ceil_div<int64_t>(P, tc::TILES * tc::TILE_M);           // 256-point steps in all
  const int pairs = n_sms / 2;
  MlpCompositePlan cp{};
  if (mlp_tc_composite_strided(S)) {
    cp.clusters = (int)std::min<int64_t>((cta_steps + 1) / 2, pairs);
    return cp;
  }
  const int64_t period = S / std::gcd(S, tc::TILES * tc::TILE_M);                    // steps after which a range is on a ray boundary again
  const int64_t iters = ceil_div<int64_t>(ceil_div<int64_t>(cta_steps, 2 * pairs), period) * period;
  cp.clusters = (int)ceil_div<int64_t>(cta_steps, 2 * iters);
  cp.chain_iters = (int)iters;
  return cp;
}

size_t mlp_tc_workspace_bytes(const scade_net_desc& d, int64_t P, int save) {
  // without a stash: the hand-over ring of the fused compositing (one 5 KB slot per CTA; sized for any device)
  return save ? (size_t)tc::train_layout(d, P).total : (size_t)256 * tc::COMP_RING_BYTES_PER_CTA;
}

int mlp_tc_stash_layout(const scade_net_desc& d, int64_t P, int64_t* out, int n) {
  const tc::TrainLayout L = tc::train_layout(d, P);
  int64_t v[64];
  int k = 0;
  v[k++] = L.T; v[k++] = L.D; v[k++] = (int64_t)L.total; v[k++] = (int64_t)L.emb; v[k++] = (int64_t)L.feat; v[k++] = (int64_t)L.hv;
  v[k++] = (int64_t)L.dzv; v[k++] = (int64_t)L.dzf; v[k++] = (int64_t)L.maskv; v[k++] = (int64_t)L.alpha; v[k++] = (int64_t)L.gs;
  for (int l = 0; l < 8; ++l) v[k++] = (int64_t)L.h[l];
  for (int l = 0; l < 8; ++l) v[k++] = (int64_t)L.dz[l];
  for (int l = 0; l < 8; ++l) v[k++] = (int64_t)L.maskh[l];
  for (int i = 0; i < n && i < k; ++i) out[i] = v[i];
  return k;
}

int mlp_tc_forward(const scade_net& net, const float* rays, int ray_stride, const float* z, const float* x_embedded,
                   int64_t N, int S, const float* bb_center, float bb_scale, float* raw_out, void* workspace,
                   size_t ws_bytes, int save, cudaStream_t st, bool x3, const MlpCompositeOut* comp) {
  if (comp != nullptr && (x3 || save || rays == nullptr || !mlp_tc_composite_supported(S) || comp->weights == nullptr)) {
    set_error("mlp_forward: fused compositing needs SCADE_PREC_TC_F16 without save_for_backward, rays mode and S a multiple of 32 (S=%d)", S);
    return SCADE_ERR_UNSUPPORTED;
  }
  if (x3 && save) {
    set_error("mlp_forward (tc_f16x3): the tight tensor-core mode is forward-only; train through SCADE_PREC_FP32 or SCADE_PREC_TC_F16");
    return SCADE_ERR_UNSUPPORTED;
  }
  SCADE_TRY(tc::set_kernel_attributes());
  tc::NetPlan plan;
  // render forward (no stash, fast mode): the views layer takes its A operand from tensor memory (SCADE_TC_VIEWS_TMEM=0: from
  // shared memory, for A/B)
  static const bool views_tmem_on = [] { const char* e = getenv("SCADE_TC_VIEWS_TMEM"); return e == nullptr || atoi(e) != 0; }();
  tc::build_plans(net, &plan, nullptr, x3, !x3 && !save && views_tmem_on);
  NetDims nd(net.desc);
  tc::FwdArgs a{};
  a.packed = reinterpret_cast<const uint8_t*>(net.packed_f16);
  a.rays = rays; a.ray_stride = ray_stride; a.z = z; a.S = S;
  a.x_embedded = x_embedded; a.in_all = nd.in_all;
  a.P = N * (int64_t)S;
  if (bb_center) { a.cx = bb_center[0]; a.cy = bb_center[1]; a.cz = bb_center[2]; }
  a.bb_scale = bb_scale;
  a.multires = net.desc.multires; a.multires_views = net.desc.multires_views;
  a.out = reinterpret_cast<float4*>(raw_out);
  a.n_pairs = ceil_div<int64_t>(a.P, tc::TILES * tc::TILE_M);
  if (comp != nullptr) {
    a.comp.weights = comp->weights; a.comp.rgb_map = comp->rgb_map; a.comp.disp_map = comp->disp_map;
    a.comp.acc_map = comp->acc_map; a.comp.depth_map = comp->depth_map;
    a.comp.write_raw = raw_out != nullptr;
    const size_t ring_bytes = (size_t)num_sms() * tc::COMP_RING_BYTES_PER_CTA;
    if (workspace == nullptr || ws_bytes < ring_bytes || (reinterpret_cast<uintptr_t>(workspace) & 15)) {
      set_error("mlp_forward (fused compositing): needs a 16-byte aligned workspace of %zu bytes (scade_mlp_workspace_bytes), got %zu",
                ring_bytes, ws_bytes);
      return SCADE_ERR_WORKSPACE;
    }
    a.comp.ring = reinterpret_cast<uint8_t*>(workspace);
    static const bool tail_on = [] { const char* e = getenv("SCADE_TC_COMP_TAIL"); return e == nullptr || atoi(e) != 0; }();   // 0: for A/B
    a.comp.tail_mode = tail_on;
  }
#if SCADE_TC_TRACE
  { const char* e = getenv("SCADE_TC_DBG"); a.dbg = e ? atoi(e) : 0; }
#endif
  // the packed stream as a 2D byte tensor: rows of 128 B (one swizzled K-major row), 128 rows per 16 KB stage
  CUtensorMap tmap;
  SCADE_TRY(tc::encode_rows_tmap(&tmap, net.packed_f16, (size_t)plan.stages_per_pass * tc::STAGE_BYTES, tc::STAGE_N));
  tc::StashArgs sa{};
  if (save) {
    sa.L = tc::train_layout(net.desc, a.P);
    if (workspace == nullptr || ws_bytes < sa.L.total) {
      set_error("mlp_forward (tc_f16, save_for_backward): workspace %zu < %llu bytes", ws_bytes, sa.L.total);
      return SCADE_ERR_WORKSPACE;
    }
    if (reinterpret_cast<uintptr_t>(workspace) & 127) {
      set_error("mlp_forward (tc_f16, save_for_backward): workspace must be 128-byte aligned");
      return SCADE_ERR_INVALID_ARGUMENT;
    }
    sa.ws = reinterpret_cast<uint8_t*>(workspace);
  }
  if (x3) {
    const int64_t x3_steps = ceil_div<int64_t>(a.P, 2 * tc::TILE_M);
    const int x3_clusters = (int)std::min<int64_t>(x3_steps, num_sms() / 2);
    void* x3_args[] = {&a, &plan, &tmap};
    SCADE_TRY(tc::launch_pair((const void*)tc::nerf_mlp_tc_x3_kernel, x3_clusters, tc::PP_THREADS, tc::FWD_SMEM_BYTES, st, x3_args));
    SCADE_LAUNCH_CHECK();
    return SCADE_OK;
  }
  int clusters = (int)std::min<int64_t>((a.n_pairs + 1) / 2, num_sms() / 2);
  if (comp != nullptr) {
    const MlpCompositePlan cp = mlp_tc_composite_plan(S, a.P, num_sms());
    clusters = cp.clusters;
    a.comp.chain_iters = cp.chain_iters;
    if (cp.chain_iters > 0) a.n_pairs = 2 * (int64_t)cp.chain_iters * clusters;      // every cluster runs exactly chain_iters steps
  }
  void* args[] = {&a, &plan, &tmap, &sa};
  if (comp != nullptr)
    SCADE_TRY(tc::launch_pair((const void*)tc::nerf_mlp_tc_pp_kernel<false, true>, clusters, tc::PP_THREADS, tc::FWD_SMEM_BYTES, st, args));
  else
    SCADE_TRY(tc::launch_pair(save ? (const void*)tc::nerf_mlp_tc_pp_kernel<true> : (const void*)tc::nerf_mlp_tc_pp_kernel<false>,
                              clusters, tc::PP_THREADS, tc::FWD_SMEM_BYTES, st, args));
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

// Backward of a forward call that stashed (save_for_backward = 1): gradients of all parameter tensors, accumulated.
int mlp_tc_backward(const scade_net& net, const float* d_out, int64_t P, float* const* grads, void* workspace, size_t ws_bytes,
                    cudaStream_t st) {
  const scade_net_desc& d = net.desc;
  NetDims nd(d);
  const tc::TrainLayout L = tc::train_layout(d, P);
  if (workspace == nullptr || ws_bytes < L.total) {
    set_error("mlp_backward (tc_f16): workspace %zu < %llu bytes", ws_bytes, L.total);
    return SCADE_ERR_WORKSPACE;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  SCADE_TRY(tc::set_kernel_attributes());
  // 1. gradient scale from max |d_out|
  SCADE_CUDA(cudaMemsetAsync(ws + L.gs, 0, 256, st));
  {
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(P, 256), 4 * num_sms());
    tc::absmax_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(d_out), reinterpret_cast<const float*>(ws + L.alpha), P,
                                                reinterpret_cast<uint32_t*>(ws + L.gs));
    SCADE_LAUNCH_CHECK();
  }
  const int64_t n_pairs = ceil_div<int64_t>(P, tc::TILES * tc::TILE_M);
  const int64_t n_steps = (n_pairs + 1) / 2;
  const int clusters = (int)std::min<int64_t>(n_steps, num_sms() / 2);
  const uint8_t* packed = reinterpret_cast<const uint8_t*>(net.packed_f16);
  const size_t bwd_off = tc::fwd_stream_bytes(d);
  // 2. dgrad chain
  {
    tc::NetPlan plan;
    tc::build_bwd_plans(net, &plan, nullptr);
    CUtensorMap tmap;
    SCADE_TRY(tc::encode_rows_tmap(&tmap, packed + bwd_off, (size_t)plan.stages_per_pass * tc::STAGE_BYTES, tc::STAGE_N));
    tc::BwdArgs a{};
    a.packed = packed;
    a.tail_off = (unsigned long long)tc::count_stages(d) * tc::STAGE_BYTES;
    a.ws = ws; a.L = L;
    a.d_out = reinterpret_cast<const float4*>(d_out);
    a.P = P; a.n_pairs = n_pairs;
    void* args[] = {&a, &plan, &tmap};
    SCADE_TRY(tc::launch_pair((const void*)tc::nerf_mlp_tc_dgrad_kernel, clusters, tc::PP_THREADS, tc::SMEM_BYTES, st, args));
    SCADE_LAUNCH_CHECK();
  }
  // 3. weight + bias gradients of the wide layers
  {
    tc::WgArgs a{};
    a.ws = ws; a.emb_off = L.emb; a.gs_off = L.gs; a.T = L.T;
    const int pv = 2 * d.D;
    for (int i = 0; i < d.D; ++i) {
      tc::WgLayer& Ld = a.layers[a.n_layers++];
      const bool after_skip = i >= 1 && (i - 1 == d.skip);
      Ld.dz_off = L.dz[i]; Ld.dz_chunks = 4; Ld.n_out = tc::W;
      Ld.has_x = i >= 1; Ld.x_off = i >= 1 ? L.h[i - 1] : 0;
      Ld.gW = grads[2 * i]; Ld.gb = grads[2 * i + 1];
      Ld.ld = i == 0 ? nd.in_ch : (after_skip ? nd.in_ch + tc::W : tc::W);
      Ld.x_col0 = after_skip ? nd.in_ch : 0;
      Ld.emb_lo = 0; Ld.emb_hi = (i == 0 || after_skip) ? nd.in_ch : 0; Ld.emb_dst = 0;
      Ld.cost = 3 + 2 * Ld.has_x;
    }
    {
      tc::WgLayer& Ld = a.layers[a.n_layers++];                    // feature_linear
      Ld.dz_off = L.dzf; Ld.dz_chunks = 4; Ld.n_out = tc::W; Ld.has_x = 1; Ld.x_off = L.h[d.D - 1];
      Ld.gW = grads[pv + 2]; Ld.gb = grads[pv + 3]; Ld.ld = tc::W; Ld.cost = 5;
    }
    {
      tc::WgLayer& Ld = a.layers[a.n_layers++];                    // views_linears.0: input = [feature, input_views]  (H:235)
      Ld.dz_off = L.dzv; Ld.dz_chunks = 2; Ld.n_out = tc::W / 2; Ld.has_x = 1; Ld.x_off = L.feat;
      Ld.gW = grads[pv]; Ld.gb = grads[pv + 1]; Ld.ld = tc::W + nd.in_views;
      Ld.emb_lo = nd.in_ch; Ld.emb_hi = nd.in_ch + nd.in_views; Ld.emb_dst = tc::W; Ld.cost = 5;
    }
    CUtensorMap tmap;
    SCADE_TRY(tc::encode_rows_tmap(&tmap, ws, (size_t)L.maskh[0], 64));          // all chunk regions precede the masks
    const int wclusters = (int)std::min<int64_t>(num_sms() / 2, std::max<int64_t>(1, L.T));
    void* args[] = {&a, &tmap};
    SCADE_TRY(tc::launch_pair((const void*)tc::nerf_mlp_tc_wgrad_kernel, wclusters, tc::WG_THREADS, tc::WG_SMEM_BYTES, st, args));
    SCADE_LAUNCH_CHECK();
  }
  // 4. alpha_linear / rgb_linear
  {
    const int pv = 2 * d.D;
    // a block walks its tiles one after the other at load latency (~12-19 us per 128-point tile): as many blocks as tiles,
    // up to 4 per SM (fewer, larger blocks were 3x slower at 512 rays)
    const int blocks = (int)std::min<int64_t>(L.T, 4 * num_sms());
    tc::head_wgrad_tc_kernel<<<blocks, 256, 0, st>>>(ws, L, reinterpret_cast<const float4*>(d_out), P, grads[pv + 4], grads[pv + 5],
                                                     grads[pv + 6], grads[pv + 7]);
    SCADE_LAUNCH_CHECK();
  }
  return SCADE_OK;
}

}  // namespace scade

#if SCADE_TC_TRACE
// trace library only (python -m scade_b200.build --trace): register the device buffer the traced warps append to
extern "C" int scade_debug_tc_trace(void* device_buf) {
  return (int)cudaMemcpyToSymbol(scade::tc::g_trace_buf, &device_buf, sizeof(device_buf));
}
#endif
