"""Achieved HBM bandwidth of the per-ray kernels (compositing, resampling, loss) against their algorithmic bytes.

    python tools/bench_perray.py [n_rays ...]        default: 4096 (headline batch), 32768 (one render chunk), 307200 (640x480 frame)

CUDA events around each call, L2 flushed before every call, mean of 10 after 3 warm-ups.  Algorithmic bytes per ray as in DESIGN §3.3.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scade_b200 import functional as F_  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = 6536.7
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = json.load(open(pp)).get("hbm_gbs", peak)


def timeit(fn, n=10, warm=3):
    ts = []
    for i in range(warm + n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts)


sizes = [int(a) for a in sys.argv[1:]] or [4096, 32768, 307200]
S, Nimp, K = 192, 128, 20
print(f"HBM peak {peak:.0f} GB/s; S = {S} samples, Nimp = {Nimp}, K = {K}")
print("| kernel | rays | us | algorithmic MB | GB/s | of peak |")
print("|---|---|---|---|---|---|")
with torch.no_grad():
    for N in sizes:
        g = torch.Generator(device=dev).manual_seed(1)
        raw = torch.randn(N, S, 4, device=dev, generator=g)
        z = torch.sort(torch.rand(N, S, device=dev, generator=g) * 4.9 + 0.1, -1).values
        rays_d = torch.randn(N, 3, device=dev, generator=g)
        w = torch.rand(N, S, device=dev, generator=g)
        pred = torch.rand(N, Nimp, device=dev, generator=g) * 4.9 + 0.1
        hyp = torch.rand(K, N, 1, device=dev, generator=g) * 4.9 + 0.1
        zc = z[:, :64].contiguous()
        wc = w[:, :64].contiguous()
        rows = [
            ("raw2outputs_fwd", lambda: F_.raw2outputs(raw, z, rays_d), N * (S * 24 + 36)),            # raw 16 + z 4 in, w 4 out / sample
            ("sample_pdf (+merge, z_std) 64 -> 128", lambda: F_.resample_from_z(zc, wc, Nimp, merge=True, std=True),
             N * 4 * (2 * 64 + 2 * Nimp + (64 + Nimp) + 1)),
            ("sample_pdf 192 -> 128", lambda: F_.resample_from_z(z, w, Nimp), N * 4 * (2 * S + 2 * Nimp)),
            ("space_carving fwd", lambda: F_.space_carving_loss(pred, hyp), N * 4 * (Nimp + K)),
        ]
        for name, fn, nbytes in rows:
            ms = timeit(fn)
            gbs = nbytes / (ms * 1e-3) / 1e9
            print(f"| {name} | {N} | {ms * 1e3:.1f} | {nbytes / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.1f}% |")
