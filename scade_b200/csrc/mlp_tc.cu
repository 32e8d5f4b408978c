// SCADE_PREC_TC_F16: the NeRF field network as ONE persistent, warp-specialised tcgen05 kernel.
//
//   reference path replaced:  run_network (run_scade_scannet.py:48-63) -> Embedder.embed
//   (model/run_nerf_helpers.py:142-172) -> NeRF.forward (H:223-247), i.e. 19 elementwise launches
//   + 12 cuBLAS SGEMMs + cats per call, each round-tripping [P,256] fp32 activations through HBM.
//
// Design (B200, sm_100a)
//   * CTA = 2 row tiles of 128 points (256 points per step), persistent over tile pairs.
//   * Activations never leave the SM: fp16 A operand tiles live in shared memory in the canonical
//     K-major SWIZZLE_128B layout (4 chunks of [128 x 64] per tile + one [128 x 64] "encoding chunk"
//     holding gamma(x), the view direction and zero padding); accumulators live in TMEM
//     (2 tiles x 256 fp32 columns = all 512 columns).
//   * Weights are pre-packed (scade_mlp_pack_f16) into the exact shared-memory image of each
//     [128 (N) x 64 (K)] fp16 stage, in consumption order, so the producer warp streams them with plain
//     16 KB cp.async.bulk (TMA) copies through a 4-stage mbarrier ring.  Both row tiles consume every
//     stage, halving L2->SMEM weight traffic per point.
//   * warp 0: TMA producer; warp 1: TMEM allocator + single-thread tcgen05.mma issuer;
//     warps 2..9: prologue/epilogue (thread == point row): positional encoding straight into the
//     swizzled A tile, then per layer TMEM -> registers -> bias + ReLU -> fp16 -> swizzled A tile of the
//     next layer.  alpha_linear (256->1) and rgb_linear (128->3) are fp32 dot products done in the
//     epilogues on the un-rounded fp32 activations; softplus(beta=10) is applied before the single
//     float4 store of (rgb_raw, sigma) per point -- the only HBM write of the kernel.
//   * skip connection (H:230) and view concat (H:235) are extra K chunks that re-use the encoding chunk;
//     nothing is concatenated in memory.
#include <cuda_fp16.h>

#include <algorithm>

#include "mlp_common.cuh"

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
constexpr int MAX_STAGE_DESCS = 96;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;

// shared-memory map (relative to a 1024-byte aligned base)
constexpr int OFF_A = 0;                                 // [TILES][4 chunks]
constexpr int OFF_EMB = OFF_A + TILES * 4 * CHUNK_BYTES;  // [TILES]
constexpr int OFF_STAGE = OFF_EMB + TILES * CHUNK_BYTES;  // [NUM_STAGES]
constexpr int OFF_BAR = OFF_STAGE + NUM_STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;         // barriers + alignment slack

enum ASrc : int { SRC_EMB = 4 };         // 0..3 = activation chunk c

struct LayerDesc {
  int n_k;                // K chunks
  int a_src[5];           // source chunk of each K chunk
  int n_halves;           // N / 128
  int relu;
  int kind;               // 0 = hidden, 1 = last hidden (also computes alpha), 2 = feature, 3 = views (final)
  int bias_idx;           // parameter index of the bias
};

struct NetPlan {
  int n_layers;
  int stages_per_pass;
  LayerDesc layers[MAX_LAYERS];
  const float* bias[MAX_LAYERS];
  const float* w_alpha; const float* b_alpha;
  const float* w_rgb; const float* b_rgb;
};

struct StageDesc {
  const float* W; int ld; int col0; int ncols; int dst_col0; int row0; int nrows;
};
struct PackPlan {
  int n_stages;
  StageDesc st[MAX_STAGE_DESCS];
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 operands, fp32 accumulate
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <- TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);       // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (ignored for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset = 1024 B, bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B
  return d;
}
// instruction descriptor for kind::f16: A = B = F16, D = F32, both K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of the 16-byte piece `piece` (8 fp16 along K) of row `row` inside a [rows x 64] SW128 chunk
__device__ __host__ __forceinline__ uint32_t sw128_offset(int row, int piece) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((piece ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

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
struct FwdArgs {
  const uint8_t* packed;          // stage images, consumption order
  const float* rays; int ray_stride; const float* z; int S;     // rays mode
  const float* x_embedded; int in_all;                          // embedded mode (rays == nullptr)
  int64_t P;
  float cx, cy, cz, bb_scale;
  int multires, multires_views;
  float4* out;
  int64_t n_pairs;
};

__global__ void __launch_bounds__(THREADS, 1) nerf_mlp_tc_kernel(const __grid_constant__ FwdArgs a,
                                                                 const __grid_constant__ NetPlan plan) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // barriers: full[4], empty[4], acc_full, a_ready, then the TMEM base address slot
  const uint32_t bar_full = sbase + OFF_BAR, bar_empty = bar_full + 8 * NUM_STAGES;
  const uint32_t bar_acc = bar_empty + 8 * NUM_STAGES, bar_aready = bar_acc + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (2 * NUM_STAGES + 2));

  if (threadIdx.x == 0) {
    for (int s = 0; s < NUM_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_aready, EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: stream the packed weight stages =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const uint8_t* src = a.packed;
        for (int s = 0; s < plan.stages_per_pass; ++s) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          mbar_expect_tx(bar_full + 8 * stage, STAGE_BYTES);
          bulk_g2s(sbase + OFF_STAGE + stage * STAGE_BYTES, src, STAGE_BYTES, bar_full + 8 * stage);
          src += STAGE_BYTES;
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      constexpr uint32_t idesc = make_idesc(TILE_M, STAGE_N);
      for (int64_t pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        for (int l = 0; l < plan.n_layers; ++l) {
          const LayerDesc& L = plan.layers[l];
          mbar_wait(bar_aready, a_phase);
          a_phase ^= 1;
          tc_fence_after();
          for (int kc = 0; kc < L.n_k; ++kc) {
            const int src = L.a_src[kc];
            for (int nh = 0; nh < L.n_halves; ++nh) {
              mbar_wait(bar_full + 8 * stage, phase);
              tc_fence_after();
              const uint32_t b_addr = sbase + OFF_STAGE + stage * STAGE_BYTES;
#pragma unroll
              for (int t = 0; t < TILES; ++t) {
                const uint32_t a_addr = (src == SRC_EMB) ? sbase + OFF_EMB + t * CHUNK_BYTES
                                                         : sbase + OFF_A + (t * 4 + src) * CHUNK_BYTES;
                const uint32_t d_addr = tmem_base + t * W + nh * STAGE_N;
#pragma unroll
                for (int ks = 0; ks < KCHUNK / 16; ++ks) {
                  mma_f16_ss(d_addr, make_smem_desc(a_addr + ks * 32), make_smem_desc(b_addr + ks * 32), idesc,
                             (kc | ks) != 0 ? 1u : 0u);
                }
              }
              mma_commit(bar_empty + 8 * stage);       // frees the weight stage when these MMAs retire
              if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
            }
          }
          mma_commit(bar_acc);                          // accumulators of this layer complete
        }
      }
    }
  } else {
    // ================= prologue / epilogue warps: thread == point row =================
    const int ew = warp - 2;
    const int tile = ew >> 2;
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    uint8_t* a_tile = smem + OFF_A + tile * 4 * CHUNK_BYTES;
    uint8_t* emb_tile = smem + OFF_EMB + tile * CHUNK_BYTES;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + tile * W;
    uint32_t acc_phase = 0;

    for (int64_t pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
      const int64_t p_raw = pair * (TILES * TILE_M) + tile * TILE_M + row;
      const bool live = p_raw < a.P;
      const int64_t p = live ? p_raw : a.P - 1;

      // ---- positional encoding -> fp16 encoding chunk (columns: gamma(x), view dir, zeros) ----
      {
        float v[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = 0.f;
        float vd[3] = {0.f, 0.f, 0.f};
        if (a.rays != nullptr) {
          const int64_t r = p / a.S;
          const float* ray = a.rays + r * a.ray_stride;
          const float zz = a.z[p];
          const float cen[3] = {a.cx, a.cy, a.cz};
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float pt = __fadd_rn(ray[d], __fmul_rn(ray[3 + d], zz));            // RS:657
            float x = __fmul_rn(__fsub_rn(pt, cen[d]), a.bb_scale);             // RS:52
            v[d] = x;
            float xp = __fmul_rn(x, 3.14159274101257324f);                      // H:165
#pragma unroll
            for (int k = 0; k < 9; ++k) {
              if (k < a.multires) {
                float s, c;
                sincos_reduced(__fmul_rn(xp, (float)(1 << k)), &s, &c);
                v[3 + 6 * k + d] = s;
                v[6 + 6 * k + d] = c;
              }
            }
            vd[d] = ray[8 + d];                                                 // RS:632
          }
        } else {
          const float* x = a.x_embedded + p * a.in_all;
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i < a.in_all) v[i] = x[i];
        }
#pragma unroll
        for (int piece = 0; piece < 8; ++piece) {
          uint4 q;
          q.x = pack_f16x2(v[piece * 8 + 0], v[piece * 8 + 1]);
          q.y = pack_f16x2(v[piece * 8 + 2], v[piece * 8 + 3]);
          q.z = pack_f16x2(v[piece * 8 + 4], v[piece * 8 + 5]);
          q.w = pack_f16x2(v[piece * 8 + 6], v[piece * 8 + 7]);
          *reinterpret_cast<uint4*>(emb_tile + sw128_offset(row, piece)) = q;
        }
        if (a.rays != nullptr) {
          // the (un-encoded, multires_views == 0) view direction sits right after gamma(x): columns in_ch..in_ch+2
          const int in_ch = 3 + 6 * a.multires;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const int col = in_ch + d;
            *reinterpret_cast<__half*>(emb_tile + sw128_offset(row, col >> 3) + (col & 7) * 2) = __float2half_rn(vd[d]);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_aready);

      float alpha = 0.f;
      for (int l = 0; l < plan.n_layers; ++l) {
        const LayerDesc& L = plan.layers[l];
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        const float* bias = plan.bias[l];
        if (L.kind != 3) {
          // hidden / feature layer: 256 columns -> next layer's A chunks
#pragma unroll 1
          for (int c8 = 0; c8 < W / 32; ++c8) {
            uint32_t r[32];
            tmem_ld32(t_lane + c8 * 32, r);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = __uint_as_float(r[j]) + __ldg(bias + c8 * 32 + j);
              f[j] = L.relu ? fmaxf(x, 0.f) : x;
            }
            if (L.kind == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) alpha = fmaf(f[j], __ldg(plan.w_alpha + c8 * 32 + j), alpha);   // H:233
            }
            uint8_t* chunk = a_tile + (c8 >> 1) * CHUNK_BYTES;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
              uint4 q;
              q.x = pack_f16x2(f[pc * 8 + 0], f[pc * 8 + 1]);
              q.y = pack_f16x2(f[pc * 8 + 2], f[pc * 8 + 3]);
              q.z = pack_f16x2(f[pc * 8 + 4], f[pc * 8 + 5]);
              q.w = pack_f16x2(f[pc * 8 + 6], f[pc * 8 + 7]);
              *reinterpret_cast<uint4*>(chunk + sw128_offset(row, (c8 & 1) * 4 + pc)) = q;
            }
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_aready);
        } else {
          // views layer (N = 128) + rgb_linear + output                          H:238-242
          float cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll 1
          for (int c8 = 0; c8 < (W / 2) / 32; ++c8) {
            uint32_t r[32];
            tmem_ld32(t_lane + c8 * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = c8 * 32 + j;
              float h = fmaxf(__uint_as_float(r[j]) + __ldg(bias + col), 0.f);
              cr = fmaf(h, __ldg(plan.w_rgb + col), cr);
              cg = fmaf(h, __ldg(plan.w_rgb + (W / 2) + col), cg);
              cb = fmaf(h, __ldg(plan.w_rgb + W + col), cb);
            }
          }
          tc_fence_before();
          if (live) {
            float al = alpha + __ldg(plan.b_alpha);
            a.out[p_raw] = make_float4(cr + __ldg(plan.b_rgb), cg + __ldg(plan.b_rgb + 1), cb + __ldg(plan.b_rgb + 2),
                                       softplus_beta10(al));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- weight packing ---------------------------------------------------------------------------------
__global__ void pack_kernel(const __grid_constant__ PackPlan plan, uint8_t* __restrict__ out) {
  const StageDesc& sd = plan.st[blockIdx.x];
  __half* dst = reinterpret_cast<__half*>(out + (size_t)blockIdx.x * STAGE_BYTES);
  for (int idx = threadIdx.x; idx < STAGE_N * KCHUNK; idx += blockDim.x) {
    int n = idx / KCHUNK, k = idx % KCHUNK;
    float v = 0.f;
    int sc = k - sd.dst_col0;
    if (n < sd.nrows && sc >= 0 && sc < sd.ncols) v = sd.W[(int64_t)(sd.row0 + n) * sd.ld + sd.col0 + sc];
    uint32_t off = sw128_offset(n, k >> 3) + (k & 7) * 2;
    dst[off >> 1] = __float2half_rn(v);
  }
}

// Build the layer table and the stage list for a network description.
static void build_plans(const scade_net& net, NetPlan* np, PackPlan* pp) {
  const scade_net_desc& d = net.desc;
  NetDims nd(d);
  NetPlan P{};
  PackPlan Q{};
  auto add_stage = [&](const float* Wt, int ld, int col0, int ncols, int dst_col0, int row0, int nrows) {
    StageDesc s{Wt, ld, col0, ncols, dst_col0, row0, nrows};
    Q.st[Q.n_stages++] = s;
  };
  auto add_layer = [&](const float* Wt, int fan_in, bool with_emb, int emb_col0, int emb_ncols, int emb_dst, int h_col0,
                       bool with_h, int n_out, int relu, int kind, int bias_idx) {
    LayerDesc L{};
    L.n_halves = n_out / STAGE_N;
    L.relu = relu; L.kind = kind; L.bias_idx = bias_idx;
    int nk = 0;
    if (with_emb) L.a_src[nk++] = SRC_EMB;
    if (with_h) for (int c = 0; c < 4; ++c) L.a_src[nk++] = c;
    L.n_k = nk;
    for (int kc = 0; kc < nk; ++kc)
      for (int nh = 0; nh < L.n_halves; ++nh) {
        if (L.a_src[kc] == SRC_EMB) add_stage(Wt, fan_in, emb_col0, emb_ncols, emb_dst, nh * STAGE_N, STAGE_N);
        else add_stage(Wt, fan_in, h_col0 + 64 * L.a_src[kc], 64, 0, nh * STAGE_N, STAGE_N);
      }
    P.bias[P.n_layers] = net.params[bias_idx];
    P.layers[P.n_layers++] = L;
  };
  for (int i = 0; i < d.D; ++i) {
    const float* Wt = net.params[2 * i];
    int kind = (i == d.D - 1) ? 1 : 0;
    if (i == 0) add_layer(Wt, nd.in_ch, true, 0, nd.in_ch, 0, 0, false, W, 1, kind, 1);
    else if (i - 1 == d.skip) add_layer(Wt, nd.in_ch + W, true, 0, nd.in_ch, 0, nd.in_ch, true, W, 1, kind, 2 * i + 1);
    else add_layer(Wt, W, false, 0, 0, 0, 0, true, W, 1, kind, 2 * i + 1);
  }
  const int pv = 2 * d.D;
  add_layer(net.params[pv + 2], W, false, 0, 0, 0, 0, true, W, 0, 2, pv + 3);                          // feature_linear
  add_layer(net.params[pv], W + nd.in_views, true, W, nd.in_views, nd.in_ch, 0, true, W / 2, 1, 3, pv + 1);   // views
  P.stages_per_pass = Q.n_stages;
  P.w_alpha = net.params[pv + 4]; P.b_alpha = net.params[pv + 5];
  P.w_rgb = net.params[pv + 6]; P.b_rgb = net.params[pv + 7];
  if (np) *np = P;
  if (pp) *pp = Q;
}

static int count_stages(const scade_net_desc& d) {
  int n = 2;                                   // layer 0: 1 K chunk x 2 halves
  for (int i = 1; i < d.D; ++i) n += ((i - 1 == d.skip) ? 5 : 4) * 2;
  n += 8;                                      // feature
  n += 5;                                      // views (N = 128)
  return n;
}

}  // namespace tc

bool mlp_tc_supported(const scade_net_desc& d) {
  NetDims nd(d);
  return d.W == tc::W && d.D >= 2 && d.D <= 8 && nd.in_all <= 64 && d.multires <= 9 && d.multires_views == 0 && d.skip != d.D - 1 &&
         tc::count_stages(d) <= tc::MAX_STAGE_DESCS;
}

size_t mlp_tc_packed_bytes(const scade_net_desc& d) { return (size_t)tc::count_stages(d) * tc::STAGE_BYTES; }

int mlp_tc_pack(const scade_net& net, void* packed_out, cudaStream_t st) {
  tc::PackPlan pp;
  tc::build_plans(net, nullptr, &pp);
  tc::pack_kernel<<<pp.n_stages, 256, 0, st>>>(pp, reinterpret_cast<uint8_t*>(packed_out));
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

size_t mlp_tc_workspace_bytes(const scade_net_desc&, int64_t, int) { return 256; }

int mlp_tc_forward(const scade_net& net, const float* rays, int ray_stride, const float* z, const float* x_embedded,
                   int64_t N, int S, const float* bb_center, float bb_scale, float* raw_out, void* workspace,
                   size_t ws_bytes, int save, cudaStream_t st) {
  (void)workspace; (void)ws_bytes;
  if (save) {
    set_error("SCADE_PREC_TC_F16 forward does not stash activations for backward in this version");
    return SCADE_ERR_UNSUPPORTED;
  }
  static bool attr_set = false;
  if (!attr_set) {
    SCADE_CUDA(cudaFuncSetAttribute(tc::nerf_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    attr_set = true;
  }
  tc::NetPlan plan;
  tc::build_plans(net, &plan, nullptr);
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
  int grid = (int)std::min<int64_t>(a.n_pairs, num_sms());
  tc::nerf_mlp_tc_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(a, plan);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

}  // namespace scade
