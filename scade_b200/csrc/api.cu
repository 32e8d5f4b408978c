// C-ABI glue: error reporting, precision dispatch of the field network, and the one-call
// render_rays orchestration (run_scade_scannet.py:581-751) on a single stream.
#include <stdarg.h>
#include <string.h>

#include "mlp_common.cuh"

namespace scade {

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int check_desc(const scade_net_desc& d) {
  if (d.D < 1 || d.D > 16 || d.W < 2 || (d.W & 1) || d.multires < 0 || d.multires > 16 || d.multires_views < 0 ||
      d.multires_views > 16 || 2 * d.D + 8 > SCADE_MAX_PARAM_TENSORS) {
    set_error("unsupported network description D=%d W=%d multires=%d multires_views=%d skip=%d", d.D, d.W, d.multires,
              d.multires_views, d.skip);
    return SCADE_ERR_INVALID_ARGUMENT;
  }
  return SCADE_OK;
}

static int check_net(const scade_net* net, int precision) {
  SCADE_CHECK_ARG(net != nullptr, "null network");
  SCADE_TRY(check_desc(net->desc));
  for (int i = 0; i < num_param_tensors(net->desc); ++i) SCADE_CHECK_ARG(net->params[i] != nullptr, "null parameter tensor %d", i);
  if (precision == SCADE_PREC_TC_F16 || precision == SCADE_PREC_TC_F16X3) {
    if (!mlp_tc_supported(net->desc)) {
      set_error("the tensor-core precisions support W=256, 2<=D<=8, 3+6*multires+3+6*multires_views<=60; got D=%d W=%d", net->desc.D,
                net->desc.W);
      return SCADE_ERR_UNSUPPORTED;
    }
    SCADE_CHECK_ARG(net->packed_f16 != nullptr, "the tensor-core precisions need packed_f16 (call scade_mlp_pack)");
  } else {
    SCADE_CHECK_ARG(precision == SCADE_PREC_FP32, "unknown precision %d", precision);
  }
  return SCADE_OK;
}

}  // namespace scade

using namespace scade;

extern "C" int scade_version(void) { return SCADE_B200_VERSION; }
extern "C" const char* scade_last_error_string(void) { return g_err; }
extern "C" uint64_t scade_kernel_launch_count(void) { return g_launch_count; }

extern "C" size_t scade_mlp_packed_bytes(const scade_net_desc* desc) {
  if (!desc || check_desc(*desc) != SCADE_OK || !mlp_tc_supported(*desc)) return 0;
  return mlp_tc_packed_bytes(*desc);
}

extern "C" int scade_mlp_pack_f16(const scade_net* net, void* packed_out, void* stream) {
  SCADE_CHECK_ARG(net && packed_out, "mlp_pack_f16: null argument");
  SCADE_TRY(check_desc(net->desc));
  if (!mlp_tc_supported(net->desc)) {
    set_error("mlp_pack_f16: shape not supported by the tensor-core path");
    return SCADE_ERR_UNSUPPORTED;
  }
  for (int i = 0; i < num_param_tensors(net->desc); ++i) SCADE_CHECK_ARG(net->params[i] != nullptr, "null parameter tensor %d", i);
  return mlp_tc_pack(*net, packed_out, as_stream(stream));
}

extern "C" size_t scade_mlp_packed_bytes_for(const scade_net_desc* desc, int precision) {
  if (!desc || check_desc(*desc) != SCADE_OK || !mlp_tc_supported(*desc)) return 0;
  if (precision != SCADE_PREC_TC_F16 && precision != SCADE_PREC_TC_F16X3) return 0;
  return mlp_tc_packed_bytes(*desc, precision == SCADE_PREC_TC_F16X3);
}

extern "C" int scade_mlp_pack(const scade_net* net, int precision, void* packed_out, void* stream) {
  SCADE_CHECK_ARG(net && packed_out, "mlp_pack: null argument");
  SCADE_CHECK_ARG(precision == SCADE_PREC_TC_F16 || precision == SCADE_PREC_TC_F16X3, "mlp_pack: precision %d has no packed stream", precision);
  SCADE_TRY(check_desc(net->desc));
  if (!mlp_tc_supported(net->desc)) {
    set_error("mlp_pack: shape not supported by the tensor-core path");
    return SCADE_ERR_UNSUPPORTED;
  }
  for (int i = 0; i < num_param_tensors(net->desc); ++i) SCADE_CHECK_ARG(net->params[i] != nullptr, "null parameter tensor %d", i);
  return mlp_tc_pack(*net, packed_out, as_stream(stream), precision == SCADE_PREC_TC_F16X3);
}

extern "C" size_t scade_mlp_workspace_bytes(const scade_net_desc* desc, int64_t P, int precision, int save_for_backward) {
  if (!desc || check_desc(*desc) != SCADE_OK || P < 0) return 0;
  if (P == 0) return 256;
  if (precision == SCADE_PREC_TC_F16) return mlp_tc_workspace_bytes(*desc, P, save_for_backward);
  if (precision == SCADE_PREC_TC_F16X3) return 256;
  return mlp_fp32_workspace_bytes(*desc, P, save_for_backward);
}

extern "C" int scade_mlp_forward_rays(const scade_net* net, int precision, const float* rays, int ray_stride,
                                      const float* z_vals, int64_t N, int S, const float* bb_center_host,
                                      float bb_scale, float* raw_out, void* workspace, size_t workspace_bytes,
                                      int save_for_backward, void* stream) {
  SCADE_TRY(check_net(net, precision));
  if (N == 0) return SCADE_OK;
  SCADE_CHECK_ARG(rays && z_vals && bb_center_host && raw_out && N > 0 && S > 0 && ray_stride >= 11,
                  "mlp_forward_rays: bad arguments");
  SCADE_CHECK_ARG((reinterpret_cast<uintptr_t>(raw_out) & 15) == 0, "mlp_forward_rays: raw_out must be 16-byte aligned");
  if (N == 0) return SCADE_OK;
  SCADE_CHECK_ARG(workspace != nullptr, "mlp_forward_rays: null workspace");
  if (precision == SCADE_PREC_TC_F16 || precision == SCADE_PREC_TC_F16X3)
    return mlp_tc_forward(*net, rays, ray_stride, z_vals, nullptr, N, S, bb_center_host, bb_scale, raw_out, workspace,
                          workspace_bytes, save_for_backward, as_stream(stream), precision == SCADE_PREC_TC_F16X3);
  return mlp_fp32_forward_rays(*net, rays, ray_stride, z_vals, N, S, bb_center_host, bb_scale, raw_out, workspace,
                               workspace_bytes, save_for_backward, as_stream(stream));
}

extern "C" int scade_mlp_forward_rays_composite_supported(const scade_net_desc* desc, int precision, int S) {
  return desc != nullptr && precision == SCADE_PREC_TC_F16 && check_desc(*desc) == SCADE_OK && mlp_tc_supported(*desc) &&
         mlp_tc_composite_supported(S);
}

extern "C" int scade_mlp_forward_rays_composite(const scade_net* net, int precision, const float* rays, int ray_stride,
                                                const float* z_vals, int64_t N, int S, const float* bb_center_host,
                                                float bb_scale, float* raw_out, float* weights, float* rgb_map,
                                                float* disp_map, float* acc_map, float* depth_map, void* workspace,
                                                size_t workspace_bytes, void* stream) {
  SCADE_TRY(check_net(net, precision));
  if (N == 0) return SCADE_OK;
  SCADE_CHECK_ARG(rays && z_vals && bb_center_host && weights && N > 0 && S > 0 && ray_stride >= 11,
                  "mlp_forward_rays_composite: bad arguments");
  SCADE_CHECK_ARG((reinterpret_cast<uintptr_t>(raw_out) & 15) == 0, "mlp_forward_rays_composite: raw_out must be 16-byte aligned");
  if (!scade_mlp_forward_rays_composite_supported(&net->desc, precision, S)) {
    set_error("mlp_forward_rays_composite: needs SCADE_PREC_TC_F16 and S a multiple of 32 (precision %d, S=%d)", precision, S);
    return SCADE_ERR_UNSUPPORTED;
  }
  MlpCompositeOut co{weights, rgb_map, disp_map, acc_map, depth_map};
  return mlp_tc_forward(*net, rays, ray_stride, z_vals, nullptr, N, S, bb_center_host, bb_scale, raw_out, workspace,
                        workspace_bytes, 0, as_stream(stream), false, &co);
}

extern "C" int scade_mlp_forward_embedded(const scade_net* net, int precision, const float* x, int64_t P, float* out,
                                          void* workspace, size_t workspace_bytes, int save_for_backward, void* stream) {
  SCADE_TRY(check_net(net, precision));
  if (P == 0) return SCADE_OK;
  SCADE_CHECK_ARG(x && out && P > 0, "mlp_forward_embedded: bad arguments");
  SCADE_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "mlp_forward_embedded: out must be 16-byte aligned");
  if (P == 0) return SCADE_OK;
  SCADE_CHECK_ARG(workspace != nullptr, "mlp_forward_embedded: null workspace");
  if (precision == SCADE_PREC_TC_F16 || precision == SCADE_PREC_TC_F16X3)
    return mlp_tc_forward(*net, nullptr, 0, nullptr, x, P, 1, nullptr, 1.0f, out, workspace, workspace_bytes,
                          save_for_backward, as_stream(stream), precision == SCADE_PREC_TC_F16X3);
  return mlp_fp32_forward_embedded(*net, x, P, out, workspace, workspace_bytes, save_for_backward, as_stream(stream));
}

extern "C" int scade_mlp_backward(const scade_net* net, int precision, const float* d_out, int64_t P,
                                  float* const* grads_host, void* workspace, size_t workspace_bytes, void* stream) {
  SCADE_TRY(check_net(net, precision));
  SCADE_CHECK_ARG(d_out && grads_host && workspace && P >= 0, "mlp_backward: bad arguments");
  SCADE_CHECK_ARG((reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "mlp_backward: d_out must be 16-byte aligned");
  for (int i = 0; i < num_param_tensors(net->desc); ++i) SCADE_CHECK_ARG(grads_host[i] != nullptr, "null gradient tensor %d", i);
  if (P == 0) return SCADE_OK;
  if (precision == SCADE_PREC_TC_F16X3) {
    set_error("mlp_backward: SCADE_PREC_TC_F16X3 is forward-only");
    return SCADE_ERR_UNSUPPORTED;
  }
  if (precision == SCADE_PREC_TC_F16) return mlp_tc_backward(*net, d_out, P, grads_host, workspace, workspace_bytes, as_stream(stream));
  return mlp_fp32_backward(*net, d_out, P, grads_host, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int scade_mlp_tc_stash_layout(const scade_net_desc* desc, int64_t P, int64_t* out, int n) {
  if (!desc || check_desc(*desc) != SCADE_OK || !mlp_tc_supported(*desc) || P < 0 || (n > 0 && !out)) return 0;
  return mlp_tc_stash_layout(*desc, P, out, n);
}

extern "C" int scade_mlp_composite_plan(int S, int64_t N, int n_sms, int* clusters, int* chain_iters) {
  SCADE_CHECK_ARG(clusters && chain_iters && N >= 0 && n_sms >= 2 && mlp_tc_composite_supported(S),
                  "mlp_composite_plan: bad arguments (S must be a multiple of 32)");
  const MlpCompositePlan cp = mlp_tc_composite_plan(S, N * (int64_t)S, n_sms);
  *clusters = cp.clusters;
  *chain_iters = cp.chain_iters;
  return SCADE_OK;
}

extern "C" int scade_embed(const float* x, int64_t P, int multires, float* out, void* stream) {
  SCADE_CHECK_ARG(x && out && P >= 0 && multires >= 0 && multires <= 16, "embed: bad arguments");
  if (P == 0) return SCADE_OK;
  return embed_launch(x, P, multires, out, as_stream(stream));
}

// ---- render_rays -----------------------------------------------------------------------------------
namespace {
struct RenderLayout {
  size_t z0, raw0, w0, zs, zf, rawf, wf, hyp, mlp, mlp_bytes, total;
};
RenderLayout render_layout(const scade_render_cfg& c, const scade_net_desc& dc, const scade_net_desc& df, int64_t N) {
  RenderLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  const int Nc = c.N_samples, Nf = c.N_importance, S = Nc + Nf;
  L.z0 = take((size_t)N * Nc * 4);
  L.raw0 = take((size_t)N * Nc * 16);
  L.w0 = take((size_t)N * Nc * 4);
  L.zs = take((size_t)N * Nf * 4);
  L.zf = take((size_t)N * S * 4);
  L.rawf = take((size_t)N * S * 16);
  L.wf = take((size_t)N * S * 4);
  L.hyp = take((size_t)N * Nf * 4);
  size_t a = scade_mlp_workspace_bytes(&dc, N * Nc, c.precision, 0);
  size_t b = scade_mlp_workspace_bytes(&df, N * S, c.precision, 0);
  L.mlp_bytes = a > b ? a : b;
  L.mlp = take(L.mlp_bytes);
  L.total = off;
  return L;
}
}  // namespace

extern "C" size_t scade_render_rays_workspace_bytes(const scade_render_cfg* cfg, const scade_net_desc* coarse,
                                                    const scade_net_desc* fine, int64_t N) {
  if (!cfg || !coarse || N < 0) return 0;
  return render_layout(*cfg, *coarse, fine ? *fine : *coarse, N).total;
}

extern "C" int scade_render_rays_forward(const scade_render_cfg* cfg, const float* ray_batch, int64_t N,
                                         const scade_net* coarse, const scade_net* fine, const float* t_rand,
                                         const float* u_coarse, const float* u_fine, const scade_render_out* out,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  SCADE_CHECK_ARG(cfg && ray_batch && coarse && out && N >= 0, "render_rays_forward: null argument");
  SCADE_CHECK_ARG(cfg->N_samples >= 3 && cfg->N_importance > 0,
                  "render_rays_forward: needs N_samples >= 3 and N_importance > 0 (the reference's N_importance == 0 branch "
                  "is dead code, RS:664-695,733)");
  SCADE_CHECK_ARG(cfg->ray_stride >= 11, "render_rays_forward: ray rows need >= 11 floats (use_viewdirs=True)");
  if (fine == nullptr) fine = coarse;                                              // RS:716
  SCADE_TRY(check_net(coarse, cfg->precision));
  SCADE_TRY(check_net(fine, cfg->precision));
  const bool perturbed = t_rand != nullptr;
  SCADE_CHECK_ARG(!perturbed || (u_coarse && u_fine), "render_rays_forward: perturb > 0 needs explicit u_coarse and u_fine");
  if (N == 0) return SCADE_OK;
  RenderLayout L = render_layout(*cfg, coarse->desc, fine->desc, N);
  if (workspace == nullptr || workspace_bytes < L.total) {
    set_error("render_rays_forward: workspace %zu < %zu bytes", workspace_bytes, L.total);
    return SCADE_ERR_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  const int Nc = cfg->N_samples, Nf = cfg->N_importance, S = Nc + Nf, rs = cfg->ray_stride;
  auto f = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  float* z0 = out->z_vals0 ? out->z_vals0 : f(L.z0);
  float* w0 = out->weights0 ? out->weights0 : f(L.w0);
  float* zf = out->z_vals ? out->z_vals : f(L.zf);
  float* wf = out->weights ? out->weights : f(L.wf);
  float* rawf = out->raw ? out->raw : f(L.rawf);
  float* hyp = out->pred_hyp ? out->pred_hyp : f(L.hyp);
  // coarse pass                                                                   RS:640-660
  SCADE_TRY(scade_coarse_z_vals(ray_batch, rs, N, Nc, cfg->lindisp, t_rand, z0, stream));
  if (scade_mlp_forward_rays_composite_supported(&coarse->desc, cfg->precision, Nc)) {
    // network + compositing in one kernel (raw never reaches memory), then importance sampling + sort-merge (RS:702-713)
    SCADE_TRY(scade_mlp_forward_rays_composite(coarse, cfg->precision, ray_batch, rs, z0, N, Nc, cfg->bb_center, cfg->bb_scale,
                                               nullptr, w0, out->rgb0, out->disp0, out->acc0, out->depth0, ws + L.mlp, L.mlp_bytes,
                                               stream));
    SCADE_TRY(scade_resample_from_z(z0, w0, N, Nc, Nf, perturbed ? u_coarse : nullptr, cfg->is_joint, f(L.zs), nullptr, zf, nullptr,
                                    stream));
  } else {
    SCADE_TRY(scade_mlp_forward_rays(coarse, cfg->precision, ray_batch, rs, z0, N, Nc, cfg->bb_center, cfg->bb_scale,
                                     f(L.raw0), ws + L.mlp, L.mlp_bytes, 0, stream));
    // compositing (RS:660) + importance sampling + sort-merge (RS:702-713) in one launch
    SCADE_TRY(scade_composite_resample(f(L.raw0), z0, ray_batch + 3, rs, N, Nc, out->rgb0, out->disp0, out->acc0, w0, out->depth0,
                                       Nf, perturbed ? u_coarse : nullptr, cfg->is_joint, f(L.zs), nullptr, zf, nullptr, stream));
  }
  // fine pass                                                                     RS:714-720
  if (scade_mlp_forward_rays_composite_supported(&fine->desc, cfg->precision, S)) {
    SCADE_TRY(scade_mlp_forward_rays_composite(fine, cfg->precision, ray_batch, rs, zf, N, S, cfg->bb_center, cfg->bb_scale,
                                               out->raw, wf, out->rgb_map, out->disp_map, out->acc_map, out->depth_map, ws + L.mlp,
                                               L.mlp_bytes, stream));
    // depth hypotheses from the fine distribution (RS:723-730, 744)
    SCADE_TRY(scade_resample_from_z(zf, wf, N, S, Nf, perturbed ? u_fine : nullptr, cfg->is_joint, hyp, out->u, nullptr, out->z_std,
                                    stream));
  } else {
    SCADE_TRY(scade_mlp_forward_rays(fine, cfg->precision, ray_batch, rs, zf, N, S, cfg->bb_center, cfg->bb_scale, rawf,
                                     ws + L.mlp, L.mlp_bytes, 0, stream));
    // compositing (RS:720) + depth hypotheses from the fine distribution (RS:723-730, 744) in one launch
    SCADE_TRY(scade_composite_resample(rawf, zf, ray_batch + 3, rs, N, S, out->rgb_map, out->disp_map, out->acc_map, wf,
                                       out->depth_map, Nf, perturbed ? u_fine : nullptr, cfg->is_joint, hyp, out->u, nullptr,
                                       out->z_std, stream));
  }
  return SCADE_OK;
}
