"""Standalone check of the tcgen05 MLP kernel against the CPU emulation of its arithmetic
(oracle.nerf_forward_f16).  Run on the GPU box under `timeout`; prints diagnostics instead of asserting."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import scade_oracle as O  # noqa: E402
from scade_b200 import synthetic as syn  # noqa: E402
from scade_b200.nerf_helpers import NeRF  # noqa: E402


def run(D, P, seed=3):
    dev = torch.device("cuda:0")
    params = syn.make_nerf_params(seed=seed, D=D, W=256, bias_scale=0.1, alpha_bias=0.3)
    net = NeRF(D=D, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    net = net.to(dev)
    x = np.random.default_rng(7).uniform(-1, 1, (P, 60)).astype(np.float32)
    with torch.no_grad():
        t0 = time.time()
        out = net(torch.from_numpy(x).to(dev))
        torch.cuda.synchronize()
        out = out.cpu().numpy()
    emu = O.nerf_forward_f16(params, x)
    ref = O.nerf_forward(params, x, dtype=np.float64)
    e = np.abs(out - emu)
    print(f"D={D} P={P}: {time.time() - t0:.3f}s  max|out-emu|={e.max():.3e} mean={e.mean():.3e}  "
          f"max|out-ref|={np.abs(out - ref).max():.3e}  per-col max {e.max(0)}", flush=True)
    if e.max() > 5e-3:
        bad = np.argwhere(e > 5e-3)
        print("  first bad entries:", bad[:10].tolist())
        print("  out[0..3]:", out[:4].tolist())
        print("  emu[0..3]:", emu[:4].tolist())
        rows = np.unique(bad[:, 0])
        print("  bad rows: count", len(rows), "min", rows.min(), "max", rows.max(), "first", rows[:20].tolist())
    return e.max()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    worst = 0.0
    for D, P in [(2, 256), (2, 1000), (8, 256), (8, 5000), (8, 148 * 256 * 3 + 17)]:
        worst = max(worst, run(D, P))
    print("WORST", worst)
