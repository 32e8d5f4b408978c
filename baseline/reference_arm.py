#!/usr/bin/env python
"""Reference arm of bench.py: the UNMODIFIED reference (mikacuy/scade @ 23139b1, MIT licence) timed on this box's host cores.

    python baseline/reference_arm.py --install            # copy the reference's hot-path packages into baseline/_ref/
    python baseline/reference_arm.py --rays 4096 --steps 2 --warmup 1 [--threads T]

`--install` copies run_scade_scannet.py and the packages it imports (model/, data/, train_utils/, metric/, LICENSE;
~190 KB of Python) from /root/reference into baseline/_ref/, byte for byte.  baseline/_ref/ is git-ignored (the reference's
sources never enter this repository's history) but travels to the GPU box with the gpurun snapshot.  __graft_entry__.build()
runs the install whenever /root/reference is present.

The timed call is the reference's own public entry point for a ray batch, run_scade_scannet.render (RS:80-155) with
`rays=` (the form train_nerf uses through render_hyp, RS:963) under torch.no_grad(): viewdir normalisation, batchify_rays,
render_rays (coarse net -> raw2outputs -> sample_pdf -> sort -> fine net -> raw2outputs -> sample_pdf_return_u), on the same
synthetic rays / weights / sample counts as the CUDA arm (BASELINE.json metric: 4096 rays x (128c + 128f), two 8x256 nets,
perturb = 0).  Third-party modules the renderer never touches (configargparse, skimage, lpips, imageio, pandas) are stubbed
by the import shim of SURVEY 8(c) when they are not installed.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DST = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"
PIECES = ["run_scade_scannet.py", "model", "data", "train_utils", "metric", "LICENSE"]


def install(src=REF_SRC, dst=REF_DST):
    """Byte-for-byte copy of the reference's renderer + the packages it imports.  Returns dst, or None if src is absent."""
    if not os.path.exists(os.path.join(src, "run_scade_scannet.py")):
        return None
    os.makedirs(dst, exist_ok=True)
    for name in PIECES:
        s, d = os.path.join(src, name), os.path.join(dst, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copy2(s, d)
    return dst


def locate():
    for path in (REF_DST, REF_SRC):
        if os.path.exists(os.path.join(path, "run_scade_scannet.py")):
            return path
    return None


def run(n_rays, steps, warmup, threads, n_coarse=128, n_fine=128):
    ref = locate()
    if ref is None:
        raise SystemExit("reference_arm: neither baseline/_ref nor /root/reference holds the reference")
    os.environ["SCADE_REFERENCE"] = ref
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    torch.set_num_threads(threads)
    from scade_b200 import synthetic as syn
    from tests.golden import generate_goldens as G        # import shim + reference-module builders (no goldens are written)
    R, H = G.import_reference()
    pc, pf = G.net_pair(8, 256)
    netc, netf = G.build_ref_nerf(H, pc, 8, 256), G.build_ref_nerf(H, pf, 8, 256)
    bb_center, bb_scale = syn.bounding_box()
    qf = G.make_query_fn(R, H, bb_center, bb_scale)
    rb = syn.make_ray_batch(4096, seed=50)[:n_rays]
    rays = torch.from_numpy(np.ascontiguousarray(np.stack([rb[:, 0:3], rb[:, 3:6]], 0)))      # batch_rays [2, N, 3]  (RS:824)
    kwargs = dict(network_fn=netc, network_query_fn=qf, N_samples=n_coarse, embedded_cam=torch.tensor(()), retraw=False,
                  perturb=0.0, N_importance=n_fine, network_fine=netf, raw_noise_std=0.0)
    times, rgb = [], None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            rgb, disp, acc, extras = R.render(480, 640, None, chunk=32768, rays=rays, ndc=False, near=float(rb[0, 6]),
                                              far=float(rb[0, 7]), use_viewdirs=True, **kwargs)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    sec = float(np.mean(times))
    return {"value": n_rays / sec, "unit": "rays/s", "cores": threads, "kind": "reference", "sec_per_step": sec,
            "sample": f"all {n_rays} rays of the workload per step through the unmodified reference's run_scade_scannet.render "
                      f"(rays=, no_grad), {steps} step(s) after {warmup} warm-up, torch {torch.__version__} CPU fp32, {threads} threads",
            "rgb_mean": float(rgb.mean()), "reference_path": ref}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--install", action="store_true")
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--coarse", type=int, default=128)
    ap.add_argument("--fine", type=int, default=128)
    a = ap.parse_args()
    if a.install:
        print(install() or "no reference at " + REF_SRC)
        return
    threads = a.threads or len(os.sched_getaffinity(0))
    print(json.dumps(run(a.rays, a.steps, a.warmup, threads, a.coarse, a.fine)), flush=True)


if __name__ == "__main__":
    main()
