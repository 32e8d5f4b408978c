// Pieces shared by the fp32 and the tensor-core field-network paths.
#pragma once
#include "common.cuh"

namespace scade {

struct NetDims {
  int in_ch, in_views, in_all;
  explicit NetDims(const scade_net_desc& d) {
    in_ch = 3 + 6 * d.multires;           // get_embedder out_dim, H:151-166
    in_views = 3 + 6 * d.multires_views;
    in_all = in_ch + in_views;
  }
};

inline int num_param_tensors(const scade_net_desc& d) { return 2 * d.D + 8; }

// Channel group g of Embedder.embed (H:163-166): 0 -> x, 1+2k -> sin((x*pi)*2^k), 2+2k -> cos(...).
// pi is rounded to fp32 and the products are formed in the reference's order; sinf/cosf are the
// accurate (<= 2 ulp) versions -- the argument reaches pi*2^8, far outside __sinf's useful range.
__device__ __forceinline__ float embed_channel(float v, int g) {
  if (g == 0) return v;
  int k = (g - 1) >> 1;
  float arg = __fmul_rn(__fmul_rn(v, 3.14159274101257324f), (float)(1 << k));
  return (g & 1) ? sinf(arg) : cosf(arg);
}

// entry points implemented in mlp_fp32.cu
size_t mlp_fp32_workspace_bytes(const scade_net_desc& d, int64_t P, int save);
int mlp_fp32_forward_rays(const scade_net& net, const float* rays, int ray_stride, const float* z, int64_t N, int S,
                          const float* bb_center, float bb_scale, float* raw_out, void* workspace, size_t ws_bytes,
                          int save, cudaStream_t st);
int mlp_fp32_forward_embedded(const scade_net& net, const float* x, int64_t P, float* out, void* workspace,
                              size_t ws_bytes, int save, cudaStream_t st);
int mlp_fp32_backward(const scade_net& net, const float* d_out, int64_t P, float* const* grads, void* workspace,
                      size_t ws_bytes, cudaStream_t st);
int embed_launch(const float* x, int64_t P, int multires, float* out, cudaStream_t st);

// Outputs of the alpha compositing fused into the tensor-core forward (compute_weights + raw2outputs, RS:511-562)
struct MlpCompositeOut {
  float* weights;     // [N,S] required
  float* rgb_map;     // [N,3] nullable
  float* disp_map;    // [N]   nullable
  float* acc_map;     // [N]   nullable
  float* depth_map;   // [N]   nullable
};
// sample counts the fused epilogue handles: a ray is a whole number of 32-row warps
inline bool mlp_tc_composite_supported(int S) { return S >= 32 && S % 32 == 0 && S <= 8192; }
// whole rays inside a CTA's 256 points per step; any other supported S runs the kernel's chain mode
inline bool mlp_tc_composite_strided(int S) { return S == 32 || S == 64 || S == 128 || S == 256; }
// How the fused kernel splits P = N*S points over the SM pairs (host arithmetic only).  clusters = SM pairs launched;
// chain_iters = 0: 512-point cluster steps strided over the clusters; > 0 (chain mode): CTA c of the 2*clusters CTAs walks
// the contiguous points [c * chain_iters * 256, (c + 1) * chain_iters * 256), a range that starts and ends on a ray boundary.
struct MlpCompositePlan { int clusters; int chain_iters; };
MlpCompositePlan mlp_tc_composite_plan(int S, int64_t P, int n_sms);

// entry points implemented in mlp_tc.cu (tcgen05 path)
bool mlp_tc_supported(const scade_net_desc& d);
size_t mlp_tc_packed_bytes(const scade_net_desc& d, bool x3 = false);
int mlp_tc_pack(const scade_net& net, void* packed_out, cudaStream_t st, bool x3 = false);
size_t mlp_tc_workspace_bytes(const scade_net_desc& d, int64_t P, int save);
int mlp_tc_forward(const scade_net& net, const float* rays, int ray_stride, const float* z, const float* x_embedded,
                   int64_t N, int S, const float* bb_center, float bb_scale, float* raw_out, void* workspace,
                   size_t ws_bytes, int save, cudaStream_t st, bool x3 = false, const MlpCompositeOut* comp = nullptr);

int mlp_tc_backward(const scade_net& net, const float* d_out, int64_t P, float* const* grads, void* workspace, size_t ws_bytes,
                    cudaStream_t st);
int mlp_tc_stash_layout(const scade_net_desc& d, int64_t P, int64_t* out, int n);

}  // namespace scade
