// Ray generation and coarse sample placement.
//   get_rays / get_ray_dirs      model/run_nerf_helpers.py:285-305
//   render() batch assembly      run_scade_scannet.py:123-141
//   t_vals / z_vals / perturb    run_scade_scannet.py:640-655, 564-579
// Pure streaming kernels (HBM-bound, a few bytes per ray); arithmetic is written with explicit
// round-to-nearest intrinsics (no FMA contraction) so that it reproduces the reference's fp32
// operation order bit for bit.
#include "common.cuh"

namespace scade {

struct Camera {
  float fx, fy, cx, cy;
  float r[3][3];
  float t[3];
};

__device__ __forceinline__ void pixel_ray(const Camera& cam, int H, int row, int col, float* o, float* d) {
  // H:295  dirs = [((i+.5)-cx)/fx, (H-(j+.5)-cy)/fy, -1]
  float i = (float)col, j = (float)row;
  float d0 = __fdiv_rn(__fsub_rn(__fadd_rn(i, 0.5f), cam.cx), cam.fx);
  float d1 = __fdiv_rn(__fsub_rn(__fsub_rn((float)H, __fadd_rn(j, 0.5f)), cam.cy), cam.fy);
  float d2 = -1.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // H:297  sum(dirs * c2w[c, :3])
    d[c] = __fadd_rn(__fadd_rn(__fmul_rn(d0, cam.r[c][0]), __fmul_rn(d1, cam.r[c][1])), __fmul_rn(d2, cam.r[c][2]));
    o[c] = cam.t[c];  // H:304
  }
}

__global__ void get_rays_kernel(Camera cam, int H, int col0, int ncols, float* __restrict__ rays_o,
                                float* __restrict__ rays_d) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)H * ncols) return;
  int row = (int)(idx / ncols), col = col0 + (int)(idx % ncols);
  float o[3], d[3];
  pixel_ray(cam, H, row, col, o, d);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rays_o[idx * 3 + c] = o[c];
    rays_d[idx * 3 + c] = d[c];
  }
}

__device__ __forceinline__ void write_batch_row(float* __restrict__ out, const float* o, const float* d, float near,
                                                float far) {
  // RS:129  viewdirs = rays_d / ||rays_d||
  float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
  out[3] = d[0]; out[4] = d[1]; out[5] = d[2];
  out[6] = near; out[7] = far;
  out[8] = __fdiv_rn(d[0], n); out[9] = __fdiv_rn(d[1], n); out[10] = __fdiv_rn(d[2], n);
}

__global__ void make_ray_batch_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t N,
                                      float near, float far, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  float o[3] = {rays_o[idx * 3], rays_o[idx * 3 + 1], rays_o[idx * 3 + 2]};
  float d[3] = {rays_d[idx * 3], rays_d[idx * 3 + 1], rays_d[idx * 3 + 2]};
  write_batch_row(out + idx * 11, o, d, near, far);
}

__global__ void camera_ray_batch_kernel(Camera cam, int H, int col0, int ncols, int64_t pix0, int64_t N, float near,
                                        float far, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  int64_t pix = pix0 + idx;
  int row = (int)(pix / ncols), col = col0 + (int)(pix % ncols);
  float o[3], d[3];
  pixel_ray(cam, H, row, col, o, d);
  write_batch_row(out + idx * 11, o, d, near, far);
}

// One thread per (ray, sample).  z = near*(1-t) + far*t  (RS:648) or 1/(1/near*(1-t) + 1/far*t) (RS:651),
// then the stratified jitter of RS:566-578 when t_rand is given.
__device__ __forceinline__ float coarse_z(float near, float far, int Nc, int i, int lindisp) {
  float t = torch_linspace(0.0f, 1.0f, Nc, i);
  float omt = __fsub_rn(1.0f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
  return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(__fdiv_rn(1.0f, near), omt), __fmul_rn(__fdiv_rn(1.0f, far), t)));
}

__device__ __forceinline__ float jitter(float zm1, float z0, float zp1, bool first, bool last, float t) {
  float lower = first ? z0 : __fmul_rn(0.5f, __fadd_rn(z0, zm1));   // RS:566-568 (mids = .5*(z[1:]+z[:-1]))
  float upper = last ? z0 : __fmul_rn(0.5f, __fadd_rn(zp1, z0));
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t));   // RS:578
}

__global__ void coarse_z_kernel(const float* __restrict__ rays, int ray_stride, int64_t N, int Nc, int lindisp,
                                const float* __restrict__ t_rand, float* __restrict__ z_out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * Nc) return;
  int64_t r = idx / Nc;
  int i = (int)(idx % Nc);
  float near = rays[r * ray_stride + 6], far = rays[r * ray_stride + 7];
  float z0 = coarse_z(near, far, Nc, i, lindisp);
  if (t_rand != nullptr) {
    float zm1 = i > 0 ? coarse_z(near, far, Nc, i - 1, lindisp) : z0;
    float zp1 = i + 1 < Nc ? coarse_z(near, far, Nc, i + 1, lindisp) : z0;
    z0 = jitter(zm1, z0, zp1, i == 0, i == Nc - 1, t_rand[idx]);
  }
  z_out[idx] = z0;
}

__global__ void perturb_kernel(const float* __restrict__ z_in, const float* __restrict__ t_rand, int64_t N, int S,
                               float* __restrict__ z_out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * S) return;
  int i = (int)(idx % S);
  float z0 = z_in[idx];
  float zm1 = i > 0 ? z_in[idx - 1] : z0;
  float zp1 = i + 1 < S ? z_in[idx + 1] : z0;
  z_out[idx] = jitter(zm1, z0, zp1, i == 0, i == S - 1, t_rand[idx]);
}

static Camera make_camera(const float* intr, const float* c2w) {
  Camera cam;
  cam.fx = intr[0]; cam.fy = intr[1]; cam.cx = intr[2]; cam.cy = intr[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) cam.r[r][c] = c2w[r * 4 + c];
    cam.t[r] = c2w[r * 4 + 3];
  }
  return cam;
}

}  // namespace scade

using namespace scade;

extern "C" int scade_get_rays(int H, int W, const float* intrinsic_host, const float* c2w_host, int col0, int ncols,
                              float* rays_o, float* rays_d, void* stream) {
  SCADE_CHECK_ARG(H > 0 && W > 0 && col0 >= 0 && ncols > 0 && col0 + ncols <= W, "get_rays: bad image window");
  SCADE_CHECK_ARG(intrinsic_host && c2w_host && rays_o && rays_d, "get_rays: null pointer");
  int64_t n = (int64_t)H * ncols;
  get_rays_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(
      make_camera(intrinsic_host, c2w_host), H, col0, ncols, rays_o, rays_d);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

extern "C" int scade_make_ray_batch(const float* rays_o, const float* rays_d, int64_t N, float near, float far,
                                    float* ray_batch, void* stream) {
  SCADE_CHECK_ARG(N >= 0 && rays_o && rays_d && ray_batch, "make_ray_batch: bad arguments");
  if (N == 0) return SCADE_OK;
  make_ray_batch_kernel<<<(unsigned)ceil_div<int64_t>(N, 256), 256, 0, as_stream(stream)>>>(rays_o, rays_d, N, near,
                                                                                           far, ray_batch);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

extern "C" int scade_camera_ray_batch(int H, int W, const float* intrinsic_host, const float* c2w_host, int col0,
                                      int ncols, int64_t pix0, int64_t N, float near, float far, float* ray_batch,
                                      void* stream) {
  SCADE_CHECK_ARG(H > 0 && W > 0 && col0 >= 0 && ncols > 0 && col0 + ncols <= W, "camera_ray_batch: bad window");
  SCADE_CHECK_ARG(pix0 >= 0 && N >= 0 && pix0 + N <= (int64_t)H * ncols, "camera_ray_batch: pixel range outside image");
  SCADE_CHECK_ARG(intrinsic_host && c2w_host && ray_batch, "camera_ray_batch: null pointer");
  if (N == 0) return SCADE_OK;
  camera_ray_batch_kernel<<<(unsigned)ceil_div<int64_t>(N, 256), 256, 0, as_stream(stream)>>>(
      make_camera(intrinsic_host, c2w_host), H, col0, ncols, pix0, N, near, far, ray_batch);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

extern "C" int scade_coarse_z_vals(const float* rays, int ray_stride, int64_t N, int Nc, int lindisp,
                                   const float* t_rand, float* z_vals, void* stream) {
  SCADE_CHECK_ARG(rays && z_vals && N >= 0 && Nc > 0 && ray_stride >= 8, "coarse_z_vals: bad arguments");
  if (N == 0) return SCADE_OK;
  coarse_z_kernel<<<(unsigned)ceil_div<int64_t>(N * Nc, 256), 256, 0, as_stream(stream)>>>(rays, ray_stride, N, Nc,
                                                                                          lindisp, t_rand, z_vals);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

extern "C" int scade_perturb_z_vals(const float* z_in, const float* t_rand, int64_t N, int S, float* z_out,
                                    void* stream) {
  SCADE_CHECK_ARG(z_in && t_rand && z_out && N >= 0 && S > 0 && z_in != z_out, "perturb_z_vals: bad arguments");
  if (N == 0) return SCADE_OK;
  perturb_kernel<<<(unsigned)ceil_div<int64_t>(N * S, 256), 256, 0, as_stream(stream)>>>(z_in, t_rand, N, S, z_out);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}
