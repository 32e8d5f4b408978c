"""Forward-with-stash + tensor-core backward of one network at the C3 fine-pass size (4096 rays x 192 samples), timed per
phase with CUDA events -- target for ncu (-k regex:'pp_kernel|dgrad|wgrad|head_wgrad')."""
import os
import sys
from ctypes import byref, c_void_p

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scade_b200 import _lib, nerf_helpers as NH, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N, S = 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 192
pf = syn.make_nerf_params(seed=11, bias_scale=0.05, alpha_bias=0.5, weight_gain=1.3)
net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
net.load_state_dict({k: torch.from_numpy(v) for k, v in pf.items()})
net = net.to(dev)
h, L, prec = net.handle(), _lib.load(), _lib.PREC_TC_F16
bb_center, bb_scale = syn.bounding_box()
rb = torch.from_numpy(syn.make_ray_batch(N, seed=50)).to(dev)
z = torch.sort(torch.rand(N, S, device=dev) * 4.9 + 0.1, -1).values
P = N * S
ws = torch.empty(h.workspace_bytes(P, prec, 1), dtype=torch.uint8, device=dev)
raw = torch.empty((N, S, 4), device=dev)
d_out = torch.randn((N, S, 4), device=dev) * 1e-5
grads = [torch.zeros_like(p) for p in h.params]
arr = (c_void_p * len(grads))(*[g.data_ptr() for g in grads])
cnet = h.struct(prec)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
print(f"P = {P} points, stash = {ws.numel() / 2**30:.2f} GiB")
for i in range(n):
    flush.zero_()
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    _lib.check(L.scade_mlp_forward_rays(byref(cnet), prec, _lib.ptr(rb), 11, _lib.ptr(z), N, S, _lib.host_floats(bb_center),
                                        float(bb_scale), _lib.ptr(raw), _lib.ptr(ws), ws.numel(), 1, _lib.stream_ptr()), "fwd")
    e1.record()
    _lib.check(L.scade_mlp_backward(byref(cnet), prec, _lib.ptr(d_out), P, arr, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "bwd")
    e2.record()
    torch.cuda.synchronize()
    f, b = e0.elapsed_time(e1), e1.elapsed_time(e2)
    print(f"iter {i}: fwd+stash {f:.3f} ms ({P * 2 * 587264 / f / 1e9:.0f} TFLOP/s)   bwd {b:.3f} ms ({P * 2 * (587264 + 557696) / b / 1e9:.0f} TFLOP/s)"
          f"   total {P * 3464448 / (f + b) / 1e9:.0f} TFLOP/s")
