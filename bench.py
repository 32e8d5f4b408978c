#!/usr/bin/env python
"""Headline benchmark: rays/sec of the SCADE render path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision tc_f16|fp32]

Workload (BASELINE.json "metric"): one step = render_rays forward (coarse 8x256 MLP on 128 samples ->
compositing -> sample_pdf -> sort-merge -> fine 8x256 MLP on 256 samples -> compositing -> second
sample_pdf) over 4096 synthetic rays per GPU, all reference outputs materialised (RS:733-744, no `raw`).
Weak scaling: every rank renders its own 4096 rays, no data-path collective (rays are independent).

Prints ONE JSON line (rank 0).  `value` = rays/s with the ray batch already in HBM, timed with CUDA events
on the launching stream around each step (L2 flushed between steps, untimed); `e2e` = the same metric through
the public Python API (scade_b200.render.render_rays -> C ABI) with the ray batch in pinned HOST memory and the
image outputs read back every step.  `roofline` is for the dominant kernel (the fine-pass tcgen05 MLP launch),
`cpu_baseline` is the torch-op port of the reference's CPU path on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS, N_COARSE, N_FINE, NET_D, NET_W = 4096, 128, 128, 8, 256
FLOP_PER_EVAL = 2 * 587264            # 587,264 MAC per MLP evaluation
WORKLOAD = f"render_rays fwd, {N_RAYS} rays x ({N_COARSE}c+{N_FINE}f), {NET_D}x{NET_W} MLP x2 nets, det sampling"
METRIC = "rays/sec (4096 rays x 256 samples, 8x256 MLP)"


def set_shape(n_rays, n_coarse, n_fine, tag):
    """Re-point the render workload at another BASELINE config (C2: 1024 rays x (64c + 128f))."""
    global N_RAYS, N_COARSE, N_FINE, WORKLOAD, METRIC
    N_RAYS, N_COARSE, N_FINE = n_rays, n_coarse, n_fine
    WORKLOAD = f"{tag}: render_rays fwd, {N_RAYS} rays x ({N_COARSE}c+{N_FINE}f), {NET_D}x{NET_W} MLP x2 nets, det sampling"
    METRIC = f"rays/sec ({N_RAYS} rays x {N_COARSE + N_FINE} samples, 8x256 MLP)"


_REAL_STDOUT = None


def emit(line):
    """The one JSON line of the run, on the process's real stdout (see main())."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("bf16_tflops", 1590.0), p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"       # B200_PROFILING.md fallback figures


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (baseline/_ref, installed by __graft_entry__.build() from
# /root/reference) on this box's host cores, in a subprocess whose thread environment is set explicitly (torchrun exports
# OMP_NUM_THREADS=1 to its ranks).  Only if the reference did not travel: the torch-op port under oracle/.
# ------------------------------------------------------------------------------------------------
def workload_config():
    """`config` of the JSON line -- identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "rays_per_step": N_RAYS, "samples": f"{N_COARSE}c+{N_FINE}f", "net": f"{NET_D}x{NET_W} x2",
            "rays": f"synthetic.make_ray_batch(4096, seed=50)[:{N_RAYS}]: 640x480 pinhole camera, near 0.1, far 5.0",
            "weights": "Xavier-uniform random init (synthetic.make_nerf_params, seeds 10/11)", "sampling": "perturb=0 (deterministic)"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_arm(n_rays, steps, warmup):
    cores = host_cores()
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        env[k] = str(cores)
    env["CUDA_VISIBLE_DEVICES"] = ""                     # the reference arm is the CPU path
    script = os.path.join(ROOT, "baseline", "reference_arm.py")
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import reference_arm
    if reference_arm.locate() is not None:
        r = subprocess.run([sys.executable, script, "--rays", str(n_rays), "--steps", str(steps), "--warmup", str(warmup),
                            "--threads", str(cores), "--coarse", str(N_COARSE), "--fine", str(N_FINE)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        sys.stderr.write("bench.py: reference arm failed, falling back to the oracle port\n" + r.stderr[-2000:] + "\n")
    return cpu_arm_port(n_rays, steps, warmup, cores)


def cpu_arm_port(n_rays_sample, steps, warmup, cores):
    import torch
    from oracle import torch_port as TP
    from scade_b200 import synthetic as syn
    from tests.golden.generate_goldens import net_pair
    torch.set_num_threads(cores)
    pc, pf = net_pair(NET_D, NET_W)
    pc = {k: torch.from_numpy(v) for k, v in pc.items()}
    pf = {k: torch.from_numpy(v) for k, v in pf.items()}
    bb_center, bb_scale = syn.bounding_box()
    rb = torch.from_numpy(syn.make_ray_batch(N_RAYS, seed=50)[:n_rays_sample])
    bbc = torch.from_numpy(bb_center)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            TP.render_rays(rb, pc, pf, bbc, float(bb_scale), N_COARSE, N_FINE)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    sec = float(np.mean(times))
    return {"value": n_rays_sample / sec, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{n_rays_sample} of {N_RAYS} rays of the same workload per step, {steps} steps after {warmup} warm-up, "
                      f"torch {torch.__version__} CPU fp32, {cores} threads (oracle/torch_port.py; the reference was not found)",
            "sec_per_step": sec}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation on the host cores, the FULL 4096-ray workload per step.
    Steps are bounded (a step is ~3-12 s of CPU work) so that the run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    base = cpu_arm(N_RAYS, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "rays/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": base["sec_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(), "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from scade_b200 import _lib, functional as F_, synthetic as syn
    from scade_b200 import nerf_helpers as NH
    from scade_b200 import render as R_
    from tests.golden.generate_goldens import net_pair

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    prec = args.precision

    pc, pf = net_pair(NET_D, NET_W)

    def mk(params):
        net = NH.NeRF(D=NET_D, W=NET_W, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True,
                      precision=prec)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        net = net.to(dev)
        for p in net.parameters():
            p.requires_grad_(False)
        return net
    netc, netf = mk(pc), mk(pf)
    bb_center, bb_scale = syn.bounding_box()
    qf = R_.NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision=prec)
    kwargs = dict(network_fn=netc, network_query_fn=qf, N_samples=N_COARSE, embedded_cam=torch.tensor((), device=dev),
                  retraw=False, perturb=0.0, N_importance=N_FINE, network_fine=netf, raw_noise_std=0.0)
    rb_host = torch.from_numpy(np.ascontiguousarray(syn.make_ray_batch(4096, seed=50 + rank)[:N_RAYS])).pin_memory()
    rb_dev = rb_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return F_.render_rays_forward(rb_dev, netc.handle(), netf.handle(), N_COARSE, N_FINE, bb_center, bb_scale,
                                      precision=prec)

    out_host = {k: torch.empty(s, dtype=torch.float32).pin_memory()
                for k, s in {"rgb_map": (N_RAYS, 3), "disp_map": (N_RAYS,), "acc_map": (N_RAYS,), "depth_map": (N_RAYS,)}.items()}

    def step_e2e_eager():
        rb = rb_host.to(dev, non_blocking=True)
        with torch.no_grad():
            ret = R_.render_rays(rb, True, **kwargs)
        for k, buf in out_host.items():
            buf.copy_(ret[k], non_blocking=True)
        return ret

    # public API: render_rays for a fixed chunk size as one CUDA graph that starts with the host->device copy of the pinned ray
    # batch and ends with the device->host copies of the image maps
    graphed = R_.GraphedRenderRays(N_RAYS, host_outputs=tuple(out_host), **kwargs)
    graphed.rays_host.copy_(rb_host)

    def step_e2e():
        return graphed()

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
        step_e2e_eager()
    sync_all()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- timed: device-resident ----
    launches0 = lib.scade_kernel_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    wall0 = time.perf_counter()
    for s, e in ev:
        flush.zero_()                      # L2 flush between steps (untimed); twice = ~130 us of queued memsets, so the step's
        flush.zero_()                      # first launch is enqueued before the stream runs dry (device time, not host latency)
        s.record()
        step_resident()
        e.record()
    sync_all()
    wall = time.perf_counter() - wall0
    launches = lib.scade_kernel_launch_count() - launches0
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---- timed: end to end through the public API with host buffers ----
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
        torch.cuda.current_stream().synchronize()      # the result is on the host when the step ends
    sync_all()
    e2e_sec = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e_eager()
        torch.cuda.current_stream().synchronize()
    sync_all()
    e2e_eager_sec = time.perf_counter() - t0

    # ---- the same end-to-end step with depth-2 pipelining (scade_b200.render.PipelinedRenderRays): step s+1's rays cross PCIe
    #      and step s-1's maps travel back while step s computes; every step still copies its own inputs and results ----
    pipe = R_.PipelinedRenderRays(N_RAYS, depth=2, host_outputs=tuple(out_host), **kwargs)
    for _ in range(4):
        pipe.submit(rb_host)
    pipe.drain()
    e2e_pipe_sec = float("inf")
    for _ in range(3):                     # extra key, host-clocked: best of three passes (one host hiccup spoils a 30 ms pass)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pipe.submit(rb_host)
        pipe.drain()
        sync_all()
        e2e_pipe_sec = min(e2e_pipe_sec, time.perf_counter() - t0)
    del pipe

    # ---- dominant kernel alone: fine-pass MLP launch (4096 x 256 points) ----
    z_f = step_resident()["z_vals"]
    fused_comp = F_.composite_fusable(netf.handle(), prec, N_COARSE + N_FINE)
    mlp_ev = []
    comp_buf = F_.composite_buffers(netf.handle(), N_RAYS, N_COARSE + N_FINE, prec, device=dev) if fused_comp else None
    for i in range(3 + args.steps):
        # two flushes (~130 us of queued memsets): the start event and the launch are both enqueued while the stream is still
        # busy, so the timed interval holds the kernel and not the host's launch latency
        flush.zero_()
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if fused_comp:       # the launch the step actually makes: network + alpha compositing in one kernel
            F_.mlp_forward_rays_composite(netf.handle(), rb_dev, z_f, bb_center, bb_scale, prec, out=comp_buf)
        else:
            F_.mlp_forward_rays(netf.handle(), rb_dev, z_f, bb_center, bb_scale, prec)
        e.record()
        if i >= 3:
            mlp_ev.append((s, e))
    torch.cuda.synchronize()
    mlp_ms = float(np.mean([s.elapsed_time(e) for s, e in mlp_ev]))
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_sec * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    # ---- the path that has the collective: BASELINE config 3 train step, the 4096 rays split across the ranks ----
    train_rec = None
    if not args.no_train:
        graphed.graph, graphed.out, flush = None, None, None      # free the render graph's pool before the train step allocates its stash
        torch.cuda.empty_cache()
        train_rec = measure_train(args, dev, rank, world, args.steps)

    if rank == 0:
        burst, sustained, hbm, src = load_peaks()
        total_rays = N_RAYS * world * args.steps
        flop_launch = N_RAYS * (N_COARSE + N_FINE) * FLOP_PER_EVAL
        achieved = flop_launch / (mlp_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("mlp_fine_x3_dram_bytes_per_launch" if prec == "tc_f16x3" else "mlp_fine_dram_bytes_per_launch")
        is_tc = prec in ("tc_f16", "tc_f16x3")
        dtype = {"tc_f16": "f16 operands / f32 accumulate (tcgen05)", "fp32": "f32",
                 "tc_f16x3": "f16 hi/lo operand pairs, 3 tcgen05 passes / f32 accumulate (fp32-level tolerance)"}[prec]
        kernel = {"tc_f16": "nerf_mlp_tc_pp_kernel<false,true> = field network + alpha compositing" if fused_comp else "nerf_mlp_tc_pp_kernel",
                  "tc_f16x3": "nerf_mlp_tc_x3_kernel", "fp32": "sgemm_kernel chain"}[prec]
        line = {
            "metric": METRIC, "value": total_rays / (dev_ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": workload_config(),
            "arm": {"precision": prec, "rays_per_gpu": N_RAYS, "parallelism": f"rays x{world} (every rank renders its own 4096 rays, "
                    "no data-path collective)",
                    "l2": "flushed between steps (two 256 MiB memsets, untimed: the step is enqueued while they run); weights (2.3 MB fp16) are meant to be L2-resident"},
            "e2e": {"value": total_rays / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": int(rb_host.numel() * 4),
                    "d2h_bytes_per_step": int(sum(b.numel() * 4 for b in out_host.values())),
                    "api": "scade_b200.render.GraphedRenderRays (render_rays for a fixed chunk size replayed as one CUDA graph that "
                           "includes the H2D copy of the pinned host ray batch and the D2H copies of the rgb/disp/acc/depth maps; "
                           "stream sync every step)",
                    "pipelined_value": N_RAYS * world * args.steps / e2e_pipe_sec,
                    "pipelined_api": "scade_b200.render.PipelinedRenderRays(depth=2): two graph slots on two streams, the copies of one step "
                                     "overlap the kernels of the other; same bytes per step (rank-0 clock, best of three passes of `steps` submissions)",
                    "eager_value": N_RAYS * world * args.steps / e2e_eager_sec,
                    "eager_api": "scade_b200.render.render_rays called eagerly every step (same copies; rank-0 clock)"},
            "gpu_launches": int(launches),
            "wall_ms_timed_region": wall * 1e3,
            "roofline": {"bound": "tensor", "kernel": f"{kernel} (fine pass, {N_RAYS}x{N_COARSE + N_FINE} points)",
                         "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                         "peak_source": f"{src} bf16 dense burst (kernel timed alone)", "traffic": traffic,
                         "flop_per_launch": flop_launch, "ms_per_launch": mlp_ms,
                         "step_frac_of_sustained_peak": (N_RAYS * (2 * N_COARSE + N_FINE) * FLOP_PER_EVAL)
                         / (dev_ms / args.steps * 1e-3) / 1e12 / sustained},
            "clocks": clocks,
        }
        if train_rec is not None:
            line["train"] = train_rec
        if not args.no_cpu_baseline:
            base = cpu_arm(N_RAYS, 2, 1)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    if world > 1:
        # the other ranks must not busy-wait in an NCCL barrier (one spinning host thread each) while rank 0 times the CPU
        # baseline on the same cores: they block on the rendezvous store's socket instead
        import datetime
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("scade_bench_rank0_done", "1")
        else:
            store.wait(["scade_bench_rank0_done"], datetime.timedelta(minutes=15))
    finish(world)


def measure_train(args, dev, rank, world, steps):
    """BASELINE config 3: train step on 4096 rays (global), 64 coarse + 128 importance, K=20 hypotheses:
    forward + losses + backward + the NCCL all-reduce of the flat gradient buffer + Adam (RS:954-997).
    Strong scaling: the 4096 rays of the step are split across ranks.  The process group (world > 1) is the caller's.
    Returns the record on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from scade_b200 import nerf_helpers as NH, render as R_, synthetic as syn
    from scade_b200 import _lib
    from scade_b200.dist import shard_range, sharded_train_step
    from tests.golden.generate_goldens import net_pair
    lib = _lib.load()
    N, Nc, Nf, K = int(args.train_rays), 64, 128, 20
    pc, pf = net_pair(NET_D, NET_W)
    nets = []
    for p in (pc, pf):
        net = NH.NeRF(D=NET_D, W=NET_W, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision=args.precision)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
        nets.append(net.to(dev))
    bb_center, bb_scale = syn.bounding_box()
    qf = R_.NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision=args.precision)
    kw = dict(network_fn=nets[0], network_query_fn=qf, N_samples=Nc, embedded_cam=torch.tensor((), device=dev), perturb=1.0,
              N_importance=Nf, network_fine=nets[1], raw_noise_std=0.0)
    scale = torch.ones(1, device=dev, requires_grad=True)
    shift = torch.zeros(1, device=dev, requires_grad=True)
    flat = None
    if args.optimizer == "fused":
        # parameters of both networks + scale / shift as views of one flat buffer: one memset, one all-reduce, one Adam launch
        from scade_b200.optim import FusedAdam, flatten_parameters
        # fine net first: its gradient range is all-reduced early (under the coarse backward), the rest in one piece
        net_params = [p for n in (nets[1], nets[0]) for p in n.parameters()]
        flat = flatten_parameters(net_params, [scale, shift])
        opt = FusedAdam(net_params, lr=5e-4, betas=(0.9, 0.999), flat=flat, capturable=args.train_graph)     # RS:469
        opt_ss = FusedAdam([scale, shift], lr=1e-6, flat=flat, capturable=args.train_graph)                  # RS:888
    else:
        opt = torch.optim.Adam([p for n in nets for p in n.parameters()], lr=5e-4, betas=(0.9, 0.999))
        opt_ss = torch.optim.Adam([scale, shift], lr=1e-6)
    lo, hi = shard_range(N, rank, world)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    rb = to(syn.make_ray_batch(N, seed=80)[lo:hi])
    target_s, target_h = syn.make_train_targets(N, K=K, seed=82)
    target_s, target_h = to(target_s[lo:hi]), to(target_h[:, lo:hi])

    graphed = None
    if args.optimizer == "fused" and args.train_graph:
        # the whole step (zero_grad, forward, losses, backward, all-reduce, both Adam launches) as one CUDA graph
        from scade_b200.dist import GraphedTrainStep
        graphed = GraphedTrainStep(kw, scale, shift, flat, [opt, opt_ss], n_global=N, warmup=3, overlap=bool(args.train_overlap))

    def step():
        if graphed is not None:
            return graphed(rb, target_s, target_h)
        opt.zero_grad(set_to_none=False)
        opt_ss.zero_grad(set_to_none=False)
        losses = sharded_train_step(rb, target_s, target_h, scale, shift, kw, n_global=N, flat=flat, overlap=bool(args.train_overlap))
        opt.step()
        opt_ss.step()
        return losses
    for _ in range(max(args.warmup, 3) + (2 if graphed is not None else 0)):      # (3 eager steps, then the capturing one)
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = lib.scade_kernel_launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        losses = step()
    e.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec = None
    if rank == 0:
        flop_step = N * (Nc + Nc + Nf) * 3464448          # fwd + dgrad + wgrad per evaluation
        burst, sustained, _, src = load_peaks()
        sec = float(ms) * 1e-3 / steps
        rec = {
            "metric": f"train rays/sec ({N} rays, 64c+128f, K=20, fwd+loss+bwd+allreduce+Adam)", "value": N / sec, "unit": "rays/s",
            "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (tcgen05 fwd+dgrad+wgrad)" if args.precision == "tc_f16" else "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 3 train step", "global_rays": N, "rays_per_gpu": hi - lo,
                       "precision": "tcgen05 fwd + dgrad + wgrad, fp32 master weights / gradients / Adam" if args.precision == "tc_f16"
                       else "fp32 FFMA GEMMs (fwd+bwd)",
                       "collective": "NCCL all-reduce of the flat fp32 gradient buffer, fine-net bucket overlapped with the coarse-net backward"
                       if world > 1 else "none (1 GPU)",
                       "allreduce_bytes_per_step": int(flat.flat_grad.numel() * 4) if flat is not None else None,
                       "optimizer": args.optimizer,
                       "launch": "one CUDA graph per step (scade_b200.dist.GraphedTrainStep)" if graphed is not None else "eager"},
            "gpu_launches": int(lib.scade_kernel_launch_count() - l0) if graphed is None else int(graphed.launches_per_step * steps),
            "loss": float(losses["loss"]),
            "roofline": {"bound": "tensor", "achieved": flop_step / sec / 1e12 / world, "peak": sustained, "unit": "TFLOP/s per GPU",
                         "frac": flop_step / sec / 1e12 / world / sustained, "peak_source": f"{src} bf16 sustained", "traffic": None,
                         "note": "the dominant training kernels are HBM-bound (10 KB/point activation stash): wgrad runs at 99% of the "
                                 "measured HBM bandwidth (profiles/r01_v15_train_launches.csv, DESIGN.md 3.1b)"}}
    if graphed is not None:
        graphed.release()                                   # a live graph holds captured NCCL work: drop it before the group goes
    return rec


def run_train(args):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rec = measure_train(args, dev, rank, world, args.steps)
    if rank == 0:
        emit(rec)
    finish(world)


def finish(world):
    """Tear the process group down without ever holding the result hostage: the JSON line is out; if destroying the
    communicator hangs, leave anyway."""
    if world > 1:
        import torch.distributed as dist
        killer = threading.Timer(20.0, os._exit, (0,))
        killer.daemon = True
        killer.start()
        dist.barrier()
        dist.destroy_process_group()
        killer.cancel()


def run_image(args):
    """BASELINE configs 4 / 5: full 640x480 frames (64 coarse + 128 importance -> 192 fine samples per ray), pixels sharded
    across the ranks (strong scaling: the frame is fixed), rays generated on the device from the pixel index, one all_gather of
    the finished maps per frame.  `--workload video` additionally builds the three-panel uint8 video frame on the device
    (scade_video_frame) and reads it back to pinned host memory every frame (what render_video RS:236-260 writes to disk),
    and reports PSNR of the tensor-core frames against the fp32-arithmetic path (the reference's arithmetic) on a few frames."""
    import torch
    import torch.distributed as dist
    from scade_b200 import _lib, nerf_helpers as NH, postprocess as PP, render as R_, synthetic as syn
    from scade_b200.dist import render_image_sharded
    from tests.golden.generate_goldens import net_pair
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    H, W, Nc, Nf = 480, 640, 64, 128
    video = args.workload == "video"
    bb_center, bb_scale = syn.bounding_box()
    pc, pf = net_pair(NET_D, NET_W)

    def make_kwargs(prec):
        nets = []
        for p in (pc, pf):
            net = NH.NeRF(D=NET_D, W=NET_W, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision=prec)
            net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
            nets.append(net.to(dev).requires_grad_(False))
        qf = R_.NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision=prec)
        return dict(network_fn=nets[0], network_query_fn=qf, N_samples=Nc, embedded_cam=torch.tensor((), device=dev), retraw=False,
                    perturb=0.0, N_importance=Nf, network_fine=nets[1], raw_noise_std=0.0)
    kw = make_kwargs(args.precision)
    poses = [torch.from_numpy(p) for p in syn.spiral_poses(120)]
    keys = ("rgb_map", "depth_map", "acc_map", "z_vals", "weights") if video else ("rgb_map", "depth_map", "acc_map")
    frame_host = torch.empty((H, 3 * W, 3), dtype=torch.uint8).pin_memory() if video else None
    rgb_host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    chunk = 32768                                        # rays per render_rays call (the reference's eval chunk is 16384, RS:347)

    graphs = {}                                          # chunk size -> GraphedRenderRays, kept across frames (weights are static)

    def frame(i, kwargs=kw):
        out = render_image_sharded(H, W, syn.CAM_INTRINSIC, poses[i % len(poses)], 0.1, 5.0, kwargs, chunk=chunk, keys=keys,
                                   graph_cache=graphs if kwargs is kw else None)
        out = {k: v.reshape((H, W) + ((v.shape[-1],) if k in ("rgb_map", "z_vals", "weights") else ())) for k, v in out.items()}
        if rank == 0:
            if video:
                fr = PP.video_frame(out["rgb_map"], out["depth_map"], out["z_vals"], out["weights"], depth_scale=5.0)
                frame_host.copy_(fr["frame"], non_blocking=True)
            else:
                rgb_host.copy_(out["rgb_map"], non_blocking=True)
        return out
    for i in range(max(args.warmup, 3)):
        frame(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = lib.scade_kernel_launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        frame(i)
    e.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    psnr = None
    if video:
        # BASELINE config 5 "PSNR vs reference": the rendered frames against the fp32 ORACLE (oracle/scade_oracle.py, the CPU
        # restatement pinned to the reference's goldens) on a random 1024-pixel sub-sample of three frames -- the checker leg
        from oracle import scade_oracle as O
        from scade_b200 import functional as F_
        vals = []
        for i in (0, 40, 80):
            a = frame(i)["rgb_map"].reshape(H * W, 3)
            if rank == 0:
                pix = np.random.default_rng(1000 + i).choice(H * W, 1024, replace=False)
                rays = F_.camera_ray_batch(H, W, syn.CAM_INTRINSIC, poses[i % len(poses)], 0.1, 5.0, device=dev)
                rb_np = rays[torch.from_numpy(pix).to(dev)].cpu().numpy()
                ref = O.render_rays(rb_np, pc, pf, bb_center, bb_scale, Nc, Nf)["rgb_map"]
                got = a[torch.from_numpy(pix).to(dev)].cpu().numpy()
                vals.append(float(-10.0 * np.log10(np.mean((got.astype(np.float64) - ref) ** 2) + 1e-30)))
        psnr = vals
    if rank == 0:
        sec = float(ms) * 1e-3 / args.steps
        burst, sustained, _, src = load_peaks()
        flop_frame = H * W * (Nc + Nc + Nf) * FLOP_PER_EVAL
        line = {"metric": "rays/sec, full 640x480 frames (64c+128f -> 192 samples/ray), pixels sharded across GPUs",
                "value": H * W / sec, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": sec * 1e3, "frames_per_s": 1.0 / sec, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16 operands / f32 accumulate (tcgen05)" if args.precision == "tc_f16" else "f32", "data": "synthetic",
                "config": {"workload": "BASELINE config 5 (spiral video frames incl. device post-processing + frame read-back)" if video
                           else "BASELINE config 4 (full image render, maps gathered on every rank, rgb read back)",
                           "H": H, "W": W, "samples": f"{Nc}c+{Nf}f", "chunk": chunk, "parallelism": f"pixels x{world}",
                           "poses": "120-pose synthetic spiral (synthetic.spiral_poses)"},
                "gpu_launches": int(lib.scade_kernel_launch_count() - l0),
                "roofline": {"bound": "tensor", "achieved": flop_frame / sec / 1e12 / world, "peak": sustained, "unit": "TFLOP/s per GPU",
                             "frac": flop_frame / sec / 1e12 / world / sustained, "peak_source": f"{src} bf16 sustained", "traffic": None}}
        if psnr is not None:
            line["psnr_vs_oracle_db"] = psnr
            line["psnr_note"] = "rgb of frames 0 / 40 / 80 vs the fp32 CPU oracle on 1024 random pixels per frame"
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tc_f16", choices=["tc_f16", "tc_f16x3", "fp32"],
                    help="tc_f16 = tcgen05, fp16 operands (fast, default); tc_f16x3 = tcgen05 at an fp32-level tolerance (fp16 hi/lo operand "
                         "pairs, three MMA passes); fp32 = FFMA GEMMs (the reference's arithmetic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="render workload: skip the attached config-3 train-step record")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="train workload: fused = flat parameters + scade_adam_step (default); torch = torch.optim.Adam on 48 tensors")
    ap.add_argument("--train-graph", type=int, default=1, help="train workload: replay the step as one CUDA graph (0 = eager)")
    ap.add_argument("--train-rays", type=int, default=4096, help="train workload: global rays per step (BASELINE config 3: 4096)")
    ap.add_argument("--train-overlap", type=int, default=1, help="train workload: all-reduce the fine net's gradient bucket under the coarse backward")
    ap.add_argument("--workload", default="render", choices=["render", "render_c2", "train", "image", "video"],
                    help="render = BASELINE metric (default); render_c2 = config 2 (1024 rays x (64c+128f), 1 GPU); train = config 3 (4096 rays, 64c+128f, K=20, fwd+loss+bwd+allreduce+Adam); "
                         "image / video = configs 4 / 5 (full 640x480 frames, pixels sharded across the GPUs)")
    args = ap.parse_args()
    # stdout carries ONE JSON line (rank 0): everything else that writes to file descriptor 1 during the run -- NCCL's banner
    # ("NCCL version ...", whenever NCCL_DEBUG is set on the box), library chatter -- is sent to stderr; the line itself goes to
    # the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.workload == "render_c2":
        set_shape(1024, 64, 128, "BASELINE config 2")
        args.no_train = True
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    elif args.workload in ("image", "video"):
        run_image(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
