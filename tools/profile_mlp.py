"""Launch the fine-pass MLP kernel (4096 rays x 256 samples; argv[4] = another sample count) a few times -- target for ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scade_b200 import functional as F_, nerf_helpers as NH, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "tc_f16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
comp = len(sys.argv) > 3 and sys.argv[3] == "comp"          # the variant with the alpha compositing fused into the epilogue
S = int(sys.argv[4]) if len(sys.argv) > 4 else 256
pf = syn.make_nerf_params(seed=11, bias_scale=0.05, alpha_bias=0.5, weight_gain=1.3)
net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision=prec)
net.load_state_dict({k: torch.from_numpy(v) for k, v in pf.items()})
net = net.to(dev).requires_grad_(False)
bb_center, bb_scale = syn.bounding_box()
rb = torch.from_numpy(syn.make_ray_batch(4096, seed=50)).to(dev)
z = torch.sort(torch.rand(4096, S, device=dev) * 4.9 + 0.1, -1).values
with torch.no_grad():
    for i in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if comp:
            out = F_.mlp_forward_rays_composite(net.handle(), rb, z, bb_center, bb_scale, prec)
        else:
            raw = F_.mlp_forward_rays(net.handle(), rb, z, bb_center, bb_scale, prec)
        e.record()
        torch.cuda.synchronize()
        print(f"launch {i}: {s.elapsed_time(e):.3f} ms  -> {4096 * S * 2 * 587264 / s.elapsed_time(e) / 1e9:.1f} TFLOP/s")
