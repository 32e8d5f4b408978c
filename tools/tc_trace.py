"""Timeline of one cluster of the ping-pong tcgen05 MLP kernel, from in-kernel clock stamps.

    SCADE_TC_TRACE=1 python tools/tc_trace.py [n_rays]     (builds/loads the tracing variant of the library)

Lane 0 of every warp of CTA 0 (the leader of cluster 0) records (tag, clock64); tags are listed next to the TRACE()
calls in csrc/mlp_tc.cu.  Prints, per layer, the mean over steps of: epilogue duration (accumulators visible -> operand
chunks written -> signalled), the time the epilogue warps wait for accumulators, and the time the MMA-issuing warp waits
for each tile's operand.  All on one SM, so one clock.
"""
import ctypes
import os
import sys

os.environ["SCADE_TC_TRACE"] = "1"
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scade_b200 import _lib, functional as F_, nerf_helpers as NH, synthetic as syn  # noqa: E402

CAP = 8192
dev = torch.device("cuda:0")
n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
stash = len(sys.argv) > 2 and sys.argv[2] == "stash"       # trace the training forward (kStash) instead
comp = len(sys.argv) > 2 and sys.argv[2] == "comp"         # trace the kernel with the compositor warp (kComp)
S = int(sys.argv[3]) if len(sys.argv) > 3 else (192 if stash else 256)
pf = syn.make_nerf_params(seed=11, bias_scale=0.05, alpha_bias=0.5, weight_gain=1.3)
net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
net.load_state_dict({k: torch.from_numpy(v) for k, v in pf.items()})
net = net.to(dev).requires_grad_(False)
bb_center, bb_scale = syn.bounding_box()
rb = torch.from_numpy(syn.make_ray_batch(n_rays, seed=50)).to(dev)
z = torch.sort(torch.rand(n_rays, S, device=dev) * 4.9 + 0.1, -1).values
lib = _lib.load()
buf = torch.zeros(32 * CAP, dtype=torch.int64, device=dev)
lib.scade_debug_tc_trace.argtypes = [ctypes.c_void_p]
print("SCADE_TC_DBG =", os.environ.get("SCADE_TC_DBG", "0"))
with torch.no_grad():
    for i in range(3):
        if i == 2:
            assert lib.scade_debug_tc_trace(ctypes.c_void_p(buf.data_ptr())) == 0
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if stash:
            h = net.handle()
            if i == 0:
                ws = torch.empty(h.workspace_bytes(n_rays * S, _lib.PREC_TC_F16, 1), dtype=torch.uint8, device=dev)
                raw = torch.empty((n_rays, S, 4), device=dev)
                cnet = h.struct(_lib.PREC_TC_F16)
            _lib.check(lib.scade_mlp_forward_rays(ctypes.byref(cnet), _lib.PREC_TC_F16, _lib.ptr(rb), 11, _lib.ptr(z), n_rays, S,
                                                  _lib.host_floats(bb_center), float(bb_scale), _lib.ptr(raw), _lib.ptr(ws), ws.numel(), 1,
                                                  _lib.stream_ptr()), "fwd")
        elif comp:
            out = F_.mlp_forward_rays_composite(net.handle(), rb, z, bb_center, bb_scale, "tc_f16")
        else:
            raw = F_.mlp_forward_rays(net.handle(), rb, z, bb_center, bb_scale, "tc_f16")
        e.record()
        torch.cuda.synchronize()
        print(f"launch {i}: {s.elapsed_time(e):.3f} ms")
t = buf.cpu().numpy().astype(np.uint64).reshape(32, CAP)


def events(w):
    n = int(t[w, 0])
    ev = t[w, 1:n]
    return (ev >> np.uint64(48)).astype(np.int64), (ev & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64)


NL = 10
# ---- issuer ----
tag, clk = events(1)
print(f"issuer: {len(tag)} events, span {clk[-1] - clk[0]} clk")
wait = np.zeros((NL, 2)); cnt = np.zeros((NL, 2)); issue = np.zeros((NL, 2))
begin = {}
ready = {}
steps = 0
for g, c in zip(tag, clk):
    kind, lt = g >> 8, g & 0xFF
    l, tl = lt >> 1, lt & 1
    if kind == 1:
        begin[(l, tl)] = c
        if l == 0 and tl == 0:
            steps += 1
    elif kind == 2:
        if steps > 1:
            wait[l, tl] += c - begin[(l, tl)]; cnt[l, tl] += 1
        ready[(l, tl)] = c
    elif kind == 3 and steps > 1:
        issue[l, tl] += c - ready[(l, tl)]
print(f"steps traced: {steps};  clk per step: {(clk[-1] - clk[0]) / max(steps, 1):.0f}")
print("issuer, mean clk per step:  layer | wait a(0) | issue t0 | wait a(1) | issue t1")
for l in range(NL):
    c0, c1 = max(cnt[l, 0], 1), max(cnt[l, 1], 1)
    print(f"   L{l}: {wait[l, 0] / c0:8.0f} {issue[l, 0] / c0:8.0f} {wait[l, 1] / c1:8.0f} {issue[l, 1] / c1:8.0f}")
print(f"   total wait per step: {(wait[:, 0] / np.maximum(cnt[:, 0], 1)).sum() + (wait[:, 1] / np.maximum(cnt[:, 1], 1)).sum():.0f}")

# ---- epilogue warps ----
for w, name in [(4, "tile0 half0"), (8, "tile0 half1"), (12, "tile1 half0"), (16, "tile1 half1")]:
    tag, clk = events(w)
    ep = np.zeros(NL); sig = np.zeros(NL); acc_wait = np.zeros(NL); n = np.zeros(NL)
    pro = []; last = None; acc_t = {}; wr_t = {}; st = 0; p0 = None
    for g, c in zip(tag, clk):
        kind, l = g >> 8, g & 0xFF
        if g == 0x400:
            p0 = c
        elif g == 0x401:
            pro.append(c - p0)
            last = c
        elif kind == 5:
            if l == 0: st += 1
            if st > 1: acc_wait[l] += c - last; n[l] += 1
            acc_t[l] = c
        elif kind == 6:
            wr_t[l] = c
        elif kind == 7:
            if st > 1:
                ep[l] += (wr_t.get(l, c) if l in wr_t else c) - acc_t[l]
                sig[l] += c - (wr_t[l] if l in wr_t else c)
            wr_t.pop(l, None)
            last = c
    n = np.maximum(n, 1)
    print(f"warp {w} ({name}): prologue {np.mean(pro):.0f} clk;  per layer: wait-acc | acc->written | written->signalled")
    for l in range(NL):
        print(f"   L{l}: {acc_wait[l] / n[l]:8.0f} {ep[l] / n[l]:8.0f} {sig[l] / n[l]:8.0f}")
    print(f"   sums: wait {np.sum(acc_wait / n):.0f}  epilogue {np.sum(ep / n):.0f}  signal {np.sum(sig / n):.0f}  prologue {np.mean(pro):.0f}")

if comp:
    # ---- compositor warp (warp 2): per tile, wait for the parked values -> composited
    tag, clk = events(2)
    t_wait = {}; t_ready = {}; work = [[], []]; waits = [[], []]
    for g, c in zip(tag, clk):
        kind, tl = g >> 8, g & 0xFF
        if kind == 8: t_wait[tl] = c
        elif kind == 9: t_ready[tl] = c; waits[tl].append(c - t_wait[tl])
        elif kind == 10: work[tl].append(c - t_ready[tl]); last_done = c
    for tl in (0, 1):
        w_, k_ = np.array(waits[tl]), np.array(work[tl])
        print(f"compositor tile {tl}: {len(k_)} steps; work clk mean {k_.mean():.0f} (first {k_[0]}, last {k_[-1]}); wait mean {w_.mean():.0f} min {w_.min()}")
    # last epilogue stamp of CTA 0 against the compositor's end
    ends = []
    for w in range(4, 20):
        tg, ck = events(w)
        if len(ck): ends.append(ck[-1])
    print(f"compositor ends {last_done - max(ends)} clk after the last epilogue stamp of the CTA")
