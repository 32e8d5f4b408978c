// Fused Adam step on a flat fp32 parameter range.
//
//   reference: torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999)) (run_scade_scannet.py:469) stepped once per
//   iteration (RS:993) -- 48 parameter tensors of the two field networks -> with foreach kernels still ~30 launches per step.
//   Here all parameters, gradients and both moment buffers live in flat fp32 buffers (scade_b200/optim.py re-homes the
//   nn.Parameters as views), so one step is ONE streaming kernel: 16 B read + 12 B written per parameter, HBM-bound.
//
// Arithmetic = torch's single-tensor Adam (no amsgrad, no weight decay, maximize=False):
//   m <- m + (g - m)(1 - b1);  v <- v b2 + (1 - b2) g g;  p <- p - step_size * m / (sqrt(v) / sqrt(1 - b2^t) + eps),
//   step_size = lr / (1 - b1^t).  The two bias corrections are evaluated on the host in double, as torch does.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace scade {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float one_minus_b1, float b2,
                                                   float one_minus_b2, float step_size, float bc2_sqrt, float eps) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    mm = __fadd_rn(mm, __fmul_rn(__fsub_rn(gg, mm), one_minus_b1));                       // lerp_(grad, 1 - beta1)
    vv = __fadd_rn(__fmul_rn(vv, b2), __fmul_rn(__fmul_rn(one_minus_b2, gg), gg));        // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), bc2_sqrt), eps);
    pp = __fsub_rn(pp, __fmul_rn(step_size, __fdiv_rn(mm, denom)));                       // addcdiv_(m, denom, -step_size)
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) upd(p[i], g[i], m[i], v[i]);
}

// Capturable form: the step count and the learning rate live in device memory, so the launch can sit in a CUDA graph that is
// replayed every training step.  The bias corrections are formed per thread in double (the same expressions the host forms).
__global__ void __launch_bounds__(256) adam_graph_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, int64_t n, double beta1, double beta2, float eps,
                                                         const double* __restrict__ lr_dev, const int64_t* __restrict__ step_dev) {
  const double t = (double)(*step_dev + 1);                  // torch's state['step'] after this step
  const double bc1 = 1.0 - pow(beta1, t), bc2 = 1.0 - pow(beta2, t);
  const float step_size = (float)(*lr_dev / bc1), bc2_sqrt = (float)sqrt(bc2);
  const float one_minus_b1 = (float)(1.0 - beta1), b2 = (float)beta2, one_minus_b2 = (float)(1.0 - beta2);
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    mm = __fadd_rn(mm, __fmul_rn(__fsub_rn(gg, mm), one_minus_b1));
    vv = __fadd_rn(__fmul_rn(vv, b2), __fmul_rn(__fmul_rn(one_minus_b2, gg), gg));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), bc2_sqrt), eps);
    pp = __fsub_rn(pp, __fmul_rn(step_size, __fdiv_rn(mm, denom)));
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) upd(p[i], g[i], m[i], v[i]);
}

__global__ void bump_step_kernel(int64_t* step_dev) { *step_dev += 1; }

}  // namespace scade

using namespace scade;

extern "C" int scade_adam_step_graph(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                     const double* lr_dev, double beta1, double beta2, double eps, int64_t* step_dev, void* stream) {
  SCADE_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && lr_dev && step_dev && n >= 0, "adam_step_graph: bad arguments");
  SCADE_CHECK_ARG(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                    reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "adam_step_graph: buffers must be 16-byte aligned");
  if (n > 0) {
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(std::max<int64_t>(n >> 2, 1), 256), 8 * (int64_t)num_sms());
    adam_graph_kernel<<<blocks, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, (float)eps, lr_dev,
                                                             step_dev);
    SCADE_LAUNCH_CHECK();
  }
  bump_step_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

extern "C" int scade_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                               double beta1, double beta2, double eps, int64_t step, void* stream) {
  SCADE_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "adam_step: bad arguments");
  SCADE_CHECK_ARG(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                    reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
  if (n == 0) return SCADE_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(std::max<int64_t>(n >> 2, 1), 256), 8 * (int64_t)num_sms());
  adam_kernel<<<blocks, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2,
                                                     (float)(1.0 - beta2), step_size, bc2_sqrt, (float)eps);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}
