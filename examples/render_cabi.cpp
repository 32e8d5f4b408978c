// A host with no Python and no torch: renders a 640x480 frame through the C ABI of libscade_b200.so.
//
//   g++ -O2 -std=c++17 examples/render_cabi.cpp -Iinclude -I/usr/local/cuda/include -Lscade_b200/_lib -lscade_b200 \
//       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../scade_b200/_lib' -o examples/render_cabi
//   ./examples/render_cabi [H W]
//
// What a C/C++ renderer service does with the boundary of include/scade_b200.h: it owns its CUDA allocations, fills the two
// networks' fp32 parameter tensors in the reference's state_dict order (model/run_nerf_helpers.py:206-219; here: Xavier-uniform
// random values as in DenseLayer, H:136-139), packs them once (scade_mlp_pack_f16), generates the rays of a camera on the device
// (scade_camera_ray_batch) and calls scade_render_rays_forward chunk by chunk (the reference's batchify_rays loop, RS:66-78).
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "scade_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(2); } } while (0)
#define SK(x) do { int s_ = (x); if (s_ != SCADE_OK) { std::fprintf(stderr, "%s -> %d: %s\n", #x, s_, scade_last_error_string()); std::exit(3); } } while (0)

struct Net {
  scade_net net{};
  std::vector<float*> dev;
  void* packed = nullptr;
};

static float* upload(const std::vector<float>& h) {
  float* d = nullptr;
  CK(cudaMalloc(&d, h.size() * sizeof(float)));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}

// NeRF(D=8, W=256, input_ch=57, input_ch_views=3, skips=[4]) with DenseLayer's Xavier-uniform init (H:136-139)
static Net make_net(unsigned seed, float alpha_bias) {
  Net n;
  n.net.desc = scade_net_desc{8, 256, 9, 0, 4};
  std::mt19937 rng(seed);
  auto xavier = [&](int out, int in, float gain) {
    const float a = gain * std::sqrt(6.0f / (float)(in + out));
    std::uniform_real_distribution<float> U(-a, a);
    std::vector<float> w((size_t)out * in);
    for (auto& v : w) v = U(rng);
    return w;
  };
  const float relu_gain = std::sqrt(2.0f);
  int k = 0;
  auto add = [&](int out, int in, float gain, float bias) {
    n.dev.push_back(upload(xavier(out, in, gain)));
    n.net.params[k++] = n.dev.back();
    n.dev.push_back(upload(std::vector<float>((size_t)out, bias)));
    n.net.params[k++] = n.dev.back();
  };
  for (int i = 0; i < 8; ++i) add(256, i == 0 ? 57 : (i == 5 ? 256 + 57 : 256), relu_gain, 0.f);     // pts_linears (H:208-210)
  add(128, 256 + 3, relu_gain, 0.f);                                                                  // views_linears.0 (H:211)
  add(256, 256, 1.f, 0.f);                                                                            // feature_linear (H:216)
  add(1, 256, 1.f, alpha_bias);                                                                       // alpha_linear (H:217)
  add(3, 128, 1.f, 0.f);                                                                              // rgb_linear (H:218)
  const size_t bytes = scade_mlp_packed_bytes(&n.net.desc);
  if (bytes == 0) { std::fprintf(stderr, "network shape not supported by the tensor-core path\n"); std::exit(4); }
  CK(cudaMalloc(&n.packed, bytes));
  SK(scade_mlp_pack_f16(&n.net, n.packed, nullptr));
  n.net.packed_f16 = n.packed;
  return n;
}

int main(int argc, char** argv) {
  const int H = argc > 2 ? std::atoi(argv[1]) : 480, W = argc > 2 ? std::atoi(argv[2]) : 640;
  const int Nc = 64, Nf = 128, S = Nc + Nf, chunk = 32768;
  std::printf("scade_b200 C ABI version %d\n", scade_version());
  Net coarse = make_net(1, 0.5f), fine = make_net(2, 0.5f);

  scade_render_cfg cfg{};
  cfg.N_samples = Nc; cfg.N_importance = Nf; cfg.lindisp = 0; cfg.precision = SCADE_PREC_TC_F16; cfg.is_joint = 0; cfg.ray_stride = 11;
  cfg.bb_center[0] = cfg.bb_center[1] = cfg.bb_center[2] = 0.f;
  cfg.bb_scale = 0.2f;                                           // 2 / (2 * far), RS:1243-1244 with far = 5
  const float intrinsic[4] = {585.f * W / 640.f, 585.f * H / 480.f, W / 2.f, H / 2.f};
  const float c2w[12] = {1, 0, 0, 0.1f, 0, 1, 0, 0.f, 0, 0, 1, 0.2f};
  const int64_t n_pix = (int64_t)H * W;

  float *rays, *rgb, *depth, *acc, *disp;
  CK(cudaMalloc(&rays, (size_t)chunk * 11 * 4));
  CK(cudaMalloc(&rgb, (size_t)n_pix * 3 * 4));
  CK(cudaMalloc(&depth, (size_t)n_pix * 4));
  CK(cudaMalloc(&acc, (size_t)n_pix * 4));
  CK(cudaMalloc(&disp, (size_t)n_pix * 4));
  const size_t ws_bytes = scade_render_rays_workspace_bytes(&cfg, &coarse.net.desc, &fine.net.desc, chunk);
  void* ws;
  CK(cudaMalloc(&ws, ws_bytes));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  auto frame = [&]() {
    for (int64_t p0 = 0; p0 < n_pix; p0 += chunk) {
      const int64_t n = std::min<int64_t>(chunk, n_pix - p0);
      SK(scade_camera_ray_batch(H, W, intrinsic, c2w, 0, W, p0, n, 0.1f, 5.0f, rays, st));
      scade_render_out out{};
      out.rgb_map = rgb + p0 * 3; out.depth_map = depth + p0; out.acc_map = acc + p0; out.disp_map = disp + p0;
      SK(scade_render_rays_forward(&cfg, rays, n, &coarse.net, &fine.net, nullptr, nullptr, nullptr, &out, ws, ws_bytes, st));
    }
  };
  frame();                                                       // warm-up
  CK(cudaStreamSynchronize(st));
  const uint64_t l0 = scade_kernel_launch_count();
  const int frames = 3;
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < frames; ++i) frame();
  CK(cudaStreamSynchronize(st));
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / frames;

  std::vector<float> h_rgb((size_t)n_pix * 3), h_depth(n_pix), h_acc(n_pix);
  CK(cudaMemcpy(h_rgb.data(), rgb, h_rgb.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_depth.data(), depth, h_depth.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_acc.data(), acc, h_acc.size() * 4, cudaMemcpyDeviceToHost));
  double s_rgb = 0, s_depth = 0, s_acc = 0;
  bool finite = true;
  for (float v : h_rgb) { s_rgb += v; finite &= std::isfinite(v) && v >= 0.f && v <= 1.0001f; }
  for (float v : h_depth) { s_depth += v; finite &= std::isfinite(v); }
  for (float v : h_acc) { s_acc += v; finite &= std::isfinite(v) && v >= 0.f && v <= 1.0001f; }
  std::printf("%dx%d frame, %d+%d samples/ray: %.2f ms/frame = %.2f M rays/s, %llu kernel launches/frame\n", W, H, Nc, Nf, sec * 1e3,
              n_pix / sec / 1e6, (unsigned long long)((scade_kernel_launch_count() - l0) / frames));
  std::printf("mean rgb %.6f  mean depth %.6f  mean acc %.6f  %s\n", s_rgb / h_rgb.size(), s_depth / n_pix, s_acc / n_pix,
              finite ? "OK" : "NON-FINITE OR OUT-OF-RANGE VALUES");
  return finite ? 0 : 1;
}
