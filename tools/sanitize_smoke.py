"""Small eval + train invocations of every kernel family for `compute-sanitizer --tool memcheck` (run on a B200):
fast tensor-core forward with the fused compositing (64 / 128 / 256 samples per ray, and 192 / 96: chain mode) and without
(64 samples: not a multiple of 32), the tight mode, the
stand-alone per-ray kernels, one tensor-core train step through the flat-storage loss heads, and the f4 / loss entry points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scade_b200 import synthetic as syn, render as R_, nerf_helpers as NH, functional as F_, checkpoint as C
from scade_b200.dist import sharded_train_step
from scade_b200.optim import FusedAdam, flatten_parameters
from tests.golden.generate_goldens import net_pair
dev = torch.device("cuda:0")
pc, pf = net_pair(8, 256)
bb_center, bb_scale = syn.bounding_box()


def mk(p, prec):
    net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision=prec)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    return net.to(dev)


def kwargs(prec, Nc, Nf, perturb=0.0, grad=False):
    qf = R_.NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision=prec)
    return dict(network_fn=mk(pc, prec).requires_grad_(grad), network_query_fn=qf, N_samples=Nc, embedded_cam=torch.tensor((), device=dev),
                retraw=False, perturb=perturb, N_importance=Nf, network_fine=mk(pf, prec).requires_grad_(grad), raw_noise_std=0.0)


rb = torch.from_numpy(syn.make_ray_batch(77, seed=3)).to(dev)
with torch.no_grad():
    for prec, Nc, Nf in (("tc_f16", 24, 40), ("tc_f16", 64, 64), ("tc_f16", 128, 128), ("tc_f16", 64, 128), ("tc_f16", 32, 64),
                         ("tc_f16x3", 64, 64), ("fp32", 16, 16)):
        out = R_.render_rays(rb, True, **kwargs(prec, Nc, Nf))
        torch.cuda.synchronize()
        print(f"eval {prec} {Nc}c+{Nf}f ok", float(out["rgb_map"].sum()))
    out = R_.render_rays(rb, True, **dict(kwargs("tc_f16", 128, 128), retraw=True))
    torch.cuda.synchronize()
    print("eval retraw ok", tuple(out["raw"].shape))
# train step: flat storage, loss heads, fused Adam
kw = kwargs("tc_f16", 24, 40, perturb=1.0, grad=True)
scale = torch.ones(1, device=dev, requires_grad=True)
shift = torch.zeros(1, device=dev, requires_grad=True)
params = [p for n in (kw["network_fine"], kw["network_fn"]) for p in n.parameters()]
flat = flatten_parameters(params, [scale, shift])
opt = FusedAdam(params, lr=1e-4, flat=flat)
ts, th = syn.make_train_targets(77, K=7, seed=5)
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for is_joint in (False, True):
    opt.zero_grad()
    losses = sharded_train_step(rb, to(ts), to(th), scale, shift, kw, n_global=77, flat=flat, is_joint=is_joint)
    opt.step()
    torch.cuda.synchronize()
    print(f"train is_joint={is_joint} ok", float(losses["loss"]), float(scale.grad))
# f4: hypothesis store
store = C.HypothesisStore.from_float(np.random.default_rng(0).uniform(0, 6, (1, 7, 20, 30, 1)).astype(np.float32), 0.1, 5.0, device=dev)
torch.cuda.synchronize()
print("store ok", store.nbytes)
