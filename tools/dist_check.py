"""Multi-GPU correctness check, launched with torchrun (NCCL):
  (1) ray-sharded train step + single flat all-reduce == single-GPU train step (losses and every gradient);
  (2) pixel-sharded full-image render == single-GPU render.
Prints one summary line per check from rank 0; exits non-zero on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scade_b200 import nerf_helpers as NH, render as R_, synthetic as syn  # noqa: E402
from scade_b200.dist import render_image_sharded, shard_range, sharded_train_step  # noqa: E402
from tests.golden.generate_goldens import net_pair  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def make(D, W, precision, requires_grad):
    pc, pf = net_pair(D, W)
    nets = []
    for p in (pc, pf):
        net = NH.NeRF(D=D, W=W, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision=precision)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
        nets.append(net.to(dev).requires_grad_(requires_grad))
    bb_center, bb_scale = syn.bounding_box()
    qf = R_.NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision=precision)
    return dict(network_fn=nets[0], network_query_fn=qf, N_samples=64, embedded_cam=torch.tensor((), device=dev),
                perturb=1.0 if requires_grad else 0.0, N_importance=128, network_fine=nets[1], raw_noise_std=0.0)


ok = True
# ---- (1) train step ----
N, K = 1024, 20
rb = syn.make_ray_batch(N, seed=70)
t_rand, u_c, u_f = syn.make_uniforms(N, 64, 128, seed=71)
target_s, target_h = syn.make_train_targets(N, K=K, seed=72)


def run(lo, hi, n_global, group_reduce):
    kw = make(8, 256, "fp32", True)
    scale = torch.tensor([1.1], device=dev, requires_grad=True)
    shift = torch.tensor([-0.05], device=dev, requires_grad=True)
    if not group_reduce:
        # single-GPU reference: same code, world of one (no collective)
        import scade_b200.dist as D_
        saved = D_._world
        D_._world = lambda group=None: (0, 1)
        saved_ar = D_.FlatAllReduce.all_reduce
        D_.FlatAllReduce.all_reduce = lambda self, group=None: self.flat
    losses = sharded_train_step(T(rb[lo:hi]), T(target_s[lo:hi]), T(target_h[:, lo:hi]), scale, shift, kw, n_global=n_global,
                                t_rand=T(t_rand[lo:hi]), u_coarse=T(u_c[lo:hi]), u_fine=T(u_f[lo:hi]))
    if not group_reduce:
        D_._world, D_.FlatAllReduce.all_reduce = saved, saved_ar
    grads = [p.grad.clone() for net in (kw["network_fn"], kw["network_fine"]) for p in net.parameters()] + [scale.grad, shift.grad]
    return losses, grads


lo, hi = shard_range(N, rank, world)
l_sh, g_sh = run(lo, hi, N, True)
l_1, g_1 = run(0, N, N, False)
worst = 0.0
for a, b in zip(g_sh, g_1):
    worst = max(worst, float((a - b).abs().max() / (b.abs().max() + 1e-12)))
dl = abs(float(l_sh["loss"]) - float(l_1["loss"])) / float(l_1["loss"])
if rank == 0:
    print(f"train step x{world}: loss {float(l_sh['loss']):.6f} vs single {float(l_1['loss']):.6f} (rel {dl:.2e}); "
          f"worst gradient rel err {worst:.2e}", flush=True)
ok &= dl < 1e-5 and worst < 2e-3     # fp32 summation order (atomics / split-K over different row counts)

# ---- (1b) flat storage: two-bucket exchange (fine net early, under the coarse backward), default and is_joint loss ----
from scade_b200.optim import flatten_parameters  # noqa: E402


def run_flat(lo, hi, n_global, group_reduce, is_joint):
    kw = make(8, 256, "fp32", True)
    scale = torch.tensor([1.1], device=dev, requires_grad=True)
    shift = torch.tensor([-0.05], device=dev, requires_grad=True)
    flat = flatten_parameters(kw["network_fine"], kw["network_fn"], [scale, shift])
    import scade_b200.dist as D_
    saved = D_._world
    if not group_reduce:
        D_._world = lambda group=None: (0, 1)
    losses = sharded_train_step(T(rb[lo:hi]), T(target_s[lo:hi]), T(target_h[:, lo:hi]), scale, shift, kw, n_global=n_global,
                                t_rand=T(t_rand[lo:hi]), u_coarse=T(u_c[lo:hi]), u_fine=T(u_f[lo:hi]), flat=flat, is_joint=is_joint)
    D_._world = saved
    torch.cuda.synchronize()
    return losses, flat.flat_grad[:flat.numel].clone()


for is_joint in (False, True):
    l_sh, g_sh = run_flat(lo, hi, N, True, is_joint)
    l_1, g_1 = run_flat(0, N, N, False, is_joint)
    worst = float((g_sh - g_1).abs().max() / g_1.abs().max())
    dl = max(abs(float(l_sh[k]) - float(l_1[k])) / abs(float(l_1[k])) for k in ("loss", "space_carving", "img_loss", "img_loss0"))
    if rank == 0:
        print(f"flat two-bucket train step x{world} (is_joint={is_joint}): loss {float(l_sh['loss']):.6f} vs single {float(l_1['loss']):.6f}, "
              f"space_carving {float(l_sh['space_carving']):.6f} vs {float(l_1['space_carving']):.6f} (worst rel {dl:.2e}); "
              f"flat gradient rel err {worst:.2e}", flush=True)
    ok &= dl < 1e-5 and worst < 2e-3

# ---- (2) image render ----
kw = make(8, 256, "tc_f16", False)
c2w = torch.from_numpy(syn.spiral_poses(8)[3])
out = render_image_sharded(60, 80, syn.CAM_INTRINSIC, c2w, 0.1, 5.0, kw, chunk=1000, graph_cache={})
with torch.no_grad():
    rgb, disp, acc, extras = R_.render(60, 80, syn.CAM_INTRINSIC, chunk=4800, c2w=c2w, near=0.1, far=5.0, use_viewdirs=True,
                                       **{k: v for k, v in kw.items() if k != "use_viewdirs"})
same = torch.equal(out["rgb_map"], rgb) and torch.equal(out["depth_map"], extras["depth_map"])
if rank == 0:
    print(f"image render x{world}: sharded == single-GPU bit-for-bit: {same}", flush=True)
ok &= same
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
