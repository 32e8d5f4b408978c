// Shared helpers for the scade_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/scade_b200.h"

namespace scade {

// ---- error plumbing (no exceptions cross the C ABI) -------------------------------------------
void set_error(const char* fmt, ...);

#define SCADE_CHECK_ARG(cond, ...)                    \
  do {                                                \
    if (!(cond)) {                                    \
      ::scade::set_error(__VA_ARGS__);                \
      return SCADE_ERR_INVALID_ARGUMENT;              \
    }                                                 \
  } while (0)

#define SCADE_CUDA(call)                                                                     \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ::scade::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(e__), \
                         cudaGetErrorString(e__));                                           \
      return SCADE_ERR_CUDA;                                                                 \
    }                                                                                        \
  } while (0)

// every kernel launch of the library goes through this macro; the counter backs bench.py's gpu_launches claim
extern unsigned long long g_launch_count;
#define SCADE_LAUNCH_CHECK()            \
  do {                                  \
    ++::scade::g_launch_count;          \
    SCADE_CUDA(cudaGetLastError());     \
  } while (0)

#define SCADE_TRY(call)          \
  do {                           \
    int s__ = (call);            \
    if (s__ != SCADE_OK) return s__; \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T>
inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- warp primitives -------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
// inclusive scans across the 32 lanes
__device__ __forceinline__ float warp_scan_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float warp_scan_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
// inclusive suffix sum (lane i gets sum over lanes >= i)
__device__ __forceinline__ float warp_rscan_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(FULL, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

// ---- math shared between the fp32 and the tensor-core paths -------------------------------------
// F.softplus(x, beta=10, threshold=20)  (H:242)
__device__ __forceinline__ float softplus_beta10(float x) {
  float bx = x * 10.0f;
  return bx > 20.0f ? x : log1pf(expf(bx)) / 10.0f;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// torch.linspace's scalar formula (ATen RangeFactories), no FMA contraction so that it is
// bit-identical to oracle.linspace.
__device__ __forceinline__ float torch_linspace(float start, float end, int steps, int i) {
  if (steps == 1) return start;
  float step = __fdiv_rn(__fsub_rn(end, start), (float)(steps - 1));
  if (i < steps / 2) return __fadd_rn(start, __fmul_rn(step, (float)i));
  return __fsub_rn(end, __fmul_rn(step, (float)(steps - 1 - i)));
}

}  // namespace scade
