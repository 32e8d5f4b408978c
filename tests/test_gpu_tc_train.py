"""GPU parity of the tensor-core TRAINING path (SCADE_PREC_TC_F16 with save_for_backward): forward stash, tcgen05 dgrad
chain, MN-major tcgen05 wgrad, head gradients -- called through the C ABI.

Two references:
  * teacher-forced (oracle.nerf_backward_teacher_forced): float64 autograd on the activations and ReLU masks the
    forward kernel stashed.  Isolates the backward kernels; tolerance = fp16 rounding of the gradient operands.
  * the plain oracle (oracle.nerf_backward, float64): includes the few ReLU sign flips the fp16 forward causes next to
    z = 0; stated as a cosine similarity per parameter tensor.
"""
from ctypes import byref, c_void_p

import numpy as np
import pytest
import torch

from oracle import scade_oracle as O
from scade_b200 import synthetic as syn
from tests.util import decode_sign_mask, stash_unswizzle

pytestmark = pytest.mark.gpu

PARAM_ORDER = [f"pts_linears.{i}.{k}" for i in range(8) for k in ("weight", "bias")] + [
    "views_linears.0.weight", "views_linears.0.bias", "feature_linear.weight", "feature_linear.bias",
    "alpha_linear.weight", "alpha_linear.bias", "rgb_linear.weight", "rgb_linear.bias"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from scade_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def cosine(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))


def build_net(params, dev, requires_grad=True):
    from scade_b200.nerf_helpers import NeRF
    net = NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    net = net.to(dev)
    for p in net.parameters():
        p.requires_grad_(requires_grad)
    return net


@pytest.mark.parametrize("P,gscale", [(700, 1e-7), (2500, 3e5)])
def test_tc_backward_teacher_forced(dev, P, gscale):
    """Forward with stash -> backward, straight through the C ABI; every stashed tensor and every gradient is checked.
    P = 700 leaves ragged tiles (dead rows) and an odd number of tile pairs; gscale exercises gradient
    magnitudes far below and above fp16's range (handled by the per-call power-of-two scale)."""
    from scade_b200 import _lib, functional as F_
    params = syn.make_nerf_params(seed=12, D=8, W=256, bias_scale=0.1, alpha_bias=0.3)
    net = build_net(params, dev)
    rng = np.random.default_rng(13)
    x = rng.uniform(-1, 1, (P, 60)).astype(np.float32)
    d_out = (rng.standard_normal((P, 4)) * gscale).astype(np.float32)
    h, L, prec = net.handle(), _lib.load(), _lib.PREC_TC_F16
    ws = torch.zeros(h.workspace_bytes(P, prec, 1), dtype=torch.uint8, device=dev)
    out = torch.empty((P, 4), dtype=torch.float32, device=dev)
    xs = torch.from_numpy(x).to(dev)
    cnet = h.struct(prec)
    _lib.check(L.scade_mlp_forward_embedded(byref(cnet), prec, _lib.ptr(xs), P, _lib.ptr(out), _lib.ptr(ws), ws.numel(), 1,
                                            _lib.stream_ptr()), "forward")
    grads = [torch.zeros_like(p) for p in h.params]
    arr = (c_void_p * len(grads))(*[g.data_ptr() for g in grads])
    ds = torch.from_numpy(d_out).to(dev)
    _lib.check(L.scade_mlp_backward(byref(cnet), prec, _lib.ptr(ds), P, arr, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
               "backward")
    torch.cuda.synchronize()
    lay = F_.stash_layout(h, P)
    T, D, W = lay["T"], 8, 256
    assert T % 4 == 0 and T * 128 >= P and lay["total"] == ws.numel()
    buf = ws.cpu().numpy()

    # ---- forward stash against the fp32 oracle (fp16 operand rounding: 2e-3 of each tensor's scale) ----
    ref_out, acts = O.nerf_forward(params, x, return_acts=True)
    assert rel(out.cpu().numpy(), ref_out) < 4e-3
    emb = stash_unswizzle(buf, lay["emb"], T, 1)[:P]
    assert rel(emb[:, :60], x) < 5e-4 and (emb[:, 60:62] == 1.0).all() and (emb[:, 62:] == 0).all()
    hs, inact = [], []
    for l in range(D):
        hs.append(stash_unswizzle(buf, lay["h"][l], T, 4)[:P])
        assert rel(hs[l], np.maximum(acts["pre"][l], 0)) < 2e-3, l
        m = decode_sign_mask(buf[lay["maskh"][l]:lay["maskh"][l] + T * 128 * 32].view(np.uint32).reshape(T, 2, 128, 4).transpose(0, 2, 1, 3).reshape(T * 128, 8))[:P]
        inact.append(m)
        flips = m != (acts["pre"][l] < 0)
        assert flips.mean() < 1e-3 and (not flips.any() or np.abs(acts["pre"][l][flips]).max() < 5e-3), l   # only next to zero
        assert ((hs[l] == 0) | ~m).all(), l                                     # a masked element was stashed as 0
    feat = stash_unswizzle(buf, lay["feat"], T, 4)[:P]
    hv = stash_unswizzle(buf, lay["hv"], T, 2)[:P]
    al = buf[lay["alpha"]:lay["alpha"] + T * 128 * 4].view(np.float32)[:P]
    mv = decode_sign_mask(buf[lay["maskv"]:lay["maskv"] + T * 128 * 16].view(np.uint32).reshape(T * 128, 4))[:P]
    assert rel(feat, acts["hv_in"][:, :W]) < 2e-3 and rel(hv, acts["hv"]) < 3e-3 and rel(al, acts["alpha"][:, 0]) < 3e-3
    assert (mv != (acts["zv"] < 0)).mean() < 1e-3

    # ---- backward against the teacher-forced float64 reference (fp16 gradient operands, power-of-two scale) ----
    g_tf, dz = O.nerf_backward_teacher_forced(params, emb, hs, feat, hv, al, inact, mv, d_out)
    maxbits = int(buf[lay["gs"]:lay["gs"] + 4].view(np.uint32)[0])
    d_alpha = d_out[:, 3] * np.where(al * 10 > 20, 1.0, 1.0 / (1.0 + np.exp(-al.astype(np.float64) * 10)))
    gmax = max(np.abs(d_out[:, :3]).max(), np.abs(d_alpha).max())
    assert abs(np.array([maxbits], np.uint32).view(np.float32)[0] / gmax - 1) < 1e-5
    scale = 2.0 ** (5 - ((maxbits >> 23) - 127))
    assert 32.0 <= gmax * scale * (1 + 1e-5) and gmax * scale < 64.0 * (1 + 1e-5)
    assert rel(stash_unswizzle(buf, lay["dzv"], T, 2)[:P] / scale, dz["v"]) < 1e-3
    assert rel(stash_unswizzle(buf, lay["dzf"], T, 4)[:P] / scale, dz["feature"]) < 1.5e-3
    for l in range(D):
        assert rel(stash_unswizzle(buf, lay["dz"][l], T, 4)[:P] / scale, dz[l]) < 3e-3, l
    dead = stash_unswizzle(buf, lay["dz"][0], T, 4)[P:]
    assert (dead == 0).all()                                                   # rows beyond P carry no gradient
    g_ref = O.nerf_backward(params, x, d_out, dtype=np.float64)
    for name, g in zip(PARAM_ORDER, grads):
        gn = g.cpu().numpy().astype(np.float64)
        tol = 1e-5 if name.startswith(("alpha_linear", "rgb_linear")) else 3e-3      # heads: fp32 arithmetic on fp16 activations
        assert rel(gn, g_tf[name].reshape(gn.shape)) < tol, (name, rel(gn, g_tf[name].reshape(gn.shape)))
        assert cosine(gn, g_ref[name]) > 0.995, (name, cosine(gn, g_ref[name]))


def test_tc_backward_accumulates_and_autograd(dev):
    """The C call ACCUMULATES into the gradient tensors; the autograd wrapper (NeRF.forward on embedded inputs,
    H:223-247) produces the same gradients as two direct calls summed."""
    params = syn.make_nerf_params(seed=5, D=8, W=256, bias_scale=0.1, alpha_bias=0.3)
    net = build_net(params, dev)
    rng = np.random.default_rng(6)
    x = torch.from_numpy(rng.uniform(-1, 1, (3, 300, 60)).astype(np.float32)).to(dev)
    d = torch.from_numpy(rng.standard_normal((3, 300, 4)).astype(np.float32)).to(dev)
    out = net(x)
    out.backward(d)
    g1 = [p.grad.clone() for p in net.parameters()]
    out = net(x)
    out.backward(d)                                  # second backward adds to .grad
    for a, p in zip(g1, net.parameters()):
        assert torch.allclose(p.grad, 2 * a, rtol=2e-3, atol=1e-3 * float(a.abs().max()))      # fp32 atomics: order-dependent round-off only
    ref = O.nerf_backward(params, x.reshape(-1, 60).cpu().numpy(), d.reshape(-1, 4).cpu().numpy(), dtype=np.float64)
    for (name, p), a in zip(net.named_parameters(), g1):
        assert cosine(a.cpu().numpy(), ref[name]) > 0.995, name


def test_tc_train_step_matches_fp32_path(dev):
    """RS:954-985 with both networks on the tensor-core training path vs the same step on the fp32 FFMA path (itself
    checked against the oracle in test_gpu_parity.py): losses within fp16-forward tolerance, gradient directions agree.
    The fine network sees resampled z (inverse-CDF of the coarse weights), which amplifies forward differences, hence the
    looser bound there."""
    from scade_b200 import nerf_helpers as NH
    from scade_b200 import render as R_
    from scade_b200.render import NetworkQuery
    from tests.golden.generate_goldens import net_pair
    n, Nc, Nf = 256, 64, 128
    pc, pf = net_pair(8, 256)
    bb_center, bb_scale = syn.bounding_box()
    rb = syn.make_ray_batch(n, seed=30)
    t_rand, u_c, u_f = syn.make_uniforms(n, Nc, Nf, seed=31)
    target_s, target_h = syn.make_train_targets(n, K=20, seed=32)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    results = {}
    for prec in ("fp32", "tc_f16"):
        nets = []
        for p in (pc, pf):
            net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision=prec)
            net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
            nets.append(net.to(dev))
        qf = NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision=prec)
        kw = dict(network_fn=nets[0], network_query_fn=qf, N_samples=Nc, embedded_cam=torch.tensor((), device=dev),
                  perturb=1.0, N_importance=Nf, network_fine=nets[1], raw_noise_std=0.0)
        ret = R_.render_rays(to(rb), True, cached_u=to(u_f), t_rand=to(t_rand), u_coarse=to(u_c), **kw)
        img_loss = NH.img2mse(ret["rgb_map"], to(target_s))
        sc = NH.compute_space_carving_loss(ret["pred_hyp"], to(target_h), is_joint=False, norm_p=2, threshold=0.0)
        img_loss0 = NH.img2mse(ret["rgb0"], to(target_s))
        loss = img_loss + 0.007 * sc + img_loss0
        loss.backward()
        results[prec] = (float(img_loss.detach()), float(sc.detach()), float(img_loss0.detach()),
                         [[p.grad.cpu().numpy() for p in net.parameters()] for net in nets])
    a, b = results["fp32"], results["tc_f16"]
    print("losses fp32", a[:3], "tc", b[:3])
    print("cos coarse", [round(cosine(x, y), 4) for x, y in zip(a[3][0], b[3][0])])
    print("cos fine  ", [round(cosine(x, y), 4) for x, y in zip(a[3][1], b[3][1])])
    assert abs(a[2] - b[2]) < 2e-3 * abs(a[2]) and abs(a[0] - b[0]) < 2e-2 * abs(a[0]) and abs(a[1] - b[1]) < 2e-2 * abs(a[1])
    for ga, gb in zip(a[3][0], b[3][0]):                     # coarse network: deterministic sample placement
        assert cosine(ga, gb) > 0.99
    for ga, gb in zip(a[3][1], b[3][1]):                     # fine network
        assert cosine(ga, gb) > 0.9


def test_graphed_train_step_matches_eager(dev):
    """GraphedTrainStep (the whole step as one CUDA graph: zero_grad, forward, losses, backward, exchange, capturable fused Adam)
    follows the eager step: same losses and parameters after 7 steps (deterministic sampling; fp32 atomics in the weight
    gradients reorder sums, hence a tolerance), and the device-side step count / learning-rate upload work."""
    from scade_b200 import nerf_helpers as NH, render as R_
    from scade_b200.dist import GraphedTrainStep, sharded_train_step
    from scade_b200.optim import FusedAdam, flatten_parameters, update_learning_rate
    from tests.golden.generate_goldens import net_pair
    N, Nc, Nf, K = 256, 16, 32, 5
    T = lambda a, d: torch.from_numpy(np.ascontiguousarray(a)).to(d)
    rb = T(syn.make_ray_batch(N, seed=90), dev)
    target_s, target_h = syn.make_train_targets(N, K=K, seed=91)
    target_s, target_h = T(target_s, dev), T(target_h, dev)
    bb_center, bb_scale = syn.bounding_box()
    pc, pf = net_pair(8, 256)

    def setup(capturable):
        nets = []
        for p in (pc, pf):
            net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
            net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
            nets.append(net.to(dev))
        qf = R_.NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision="tc_f16")
        kw = dict(network_fn=nets[0], network_query_fn=qf, N_samples=Nc, embedded_cam=torch.tensor((), device=dev), perturb=0.0,
                  N_importance=Nf, network_fine=nets[1], raw_noise_std=0.0)
        scale = torch.ones(1, device=dev, requires_grad=True)
        shift = torch.zeros(1, device=dev, requires_grad=True)
        params = [p for n in nets for p in n.parameters()]
        flat = flatten_parameters(params, [scale, shift])
        opt = FusedAdam(params, lr=2e-5, flat=flat, capturable=capturable)        # small steps: a smooth, comparable trajectory
        opt_ss = FusedAdam([scale, shift], lr=1e-4, flat=flat, capturable=capturable)
        return kw, scale, shift, flat, opt, opt_ss

    kw, scale, shift, flat, opt, opt_ss = setup(False)
    ref_losses = []
    for i in range(7):
        if i == 5:
            update_learning_rate(opt, 5e-6)
        opt.zero_grad(); opt_ss.zero_grad()
        ls = sharded_train_step(rb, target_s, target_h, scale, shift, kw, n_global=N, flat=flat)
        opt.step(); opt_ss.step()
        ref_losses.append(float(ls["loss"]))
    ref_flat = flat.flat.detach().clone()

    kw, scale, shift, flat, opt, opt_ss = setup(True)
    step = GraphedTrainStep(kw, scale, shift, flat, [opt, opt_ss], n_global=N, warmup=2)
    losses = []
    for i in range(7):
        if i == 5:
            update_learning_rate(opt, 5e-6)
        losses.append(float(step(rb, target_s, target_h)["loss"]))
    assert step.graph is not None and opt._step == 7 and int(opt._step_t) == 7
    # Adam normalises every gradient to ~lr per step, so the atomics-order noise of the weight gradients moves parameters whose
    # gradient is ~0 by up to lr per step: the trajectories agree closely in the mean, not bit for bit
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-3)
    diff = (flat.flat.detach() - ref_flat).abs()
    assert float(diff.mean()) < 2e-6 and float(diff.max()) < 7 * 1e-4 * 2, (float(diff.mean()), float(diff.max()))   # <= 2 lr per step


def test_graphed_train_step_per_image_scale_shift_and_checkpoint(dev):
    """(1) the graphed step follows the step's image: DEPTH_SCALES / DEPTH_SHIFTS are [n_img, 1] tables (RS:878-879),
    the graph gathers row img_i through a device-side index and the gradient lands in that row only; (2) the weight streams are
    re-packed inside the graph even when an eval render ran right before the capture; (3) optimizer.state_dict() after replays
    carries the current step count, and load_state_dict() on a used FusedAdam adopts moments and step."""
    from scade_b200 import nerf_helpers as NH, render as R_
    from scade_b200.dist import GraphedTrainStep
    from scade_b200.optim import FusedAdam, flatten_parameters
    from tests.golden.generate_goldens import net_pair
    N, Nc, Nf, K, n_img = 256, 16, 32, 5, 3
    T = lambda a, d: torch.from_numpy(np.ascontiguousarray(a)).to(d)
    rb = T(syn.make_ray_batch(N, seed=90), dev)
    target_s, target_h = syn.make_train_targets(N, K=K, seed=91)
    target_s, target_h = T(target_s, dev), T(target_h, dev)
    bb_center, bb_scale = syn.bounding_box()
    pc, pf = net_pair(8, 256)
    nets = []
    for p in (pc, pf):
        net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
        net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
        nets.append(net.to(dev))
    qf = R_.NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision="tc_f16")
    kw = dict(network_fn=nets[0], network_query_fn=qf, N_samples=Nc, embedded_cam=torch.tensor((), device=dev), perturb=0.0,
              N_importance=Nf, network_fine=nets[1], raw_noise_std=0.0)
    scales = torch.nn.Parameter(torch.tensor([[1.0], [1.3], [0.8]], device=dev))                # DEPTH_SCALES (RS:878)
    shifts = torch.nn.Parameter(torch.tensor([[0.0], [-0.1], [0.2]], device=dev))               # DEPTH_SHIFTS (RS:879)
    params = [p for n in (nets[1], nets[0]) for p in n.parameters()]
    flat = flatten_parameters(params, [scales, shifts])
    opt = FusedAdam(params, lr=2e-5, flat=flat, capturable=True)
    opt_ss = FusedAdam([scales, shifts], lr=1e-3, flat=flat, capturable=True)
    step = GraphedTrainStep(kw, scales, shifts, flat, [opt, opt_ss], n_global=N, warmup=2)
    with pytest.raises(ValueError):
        step(rb, target_s, target_h)                                                              # per-image tables need img_i
    step.calls = 0
    s0, h0 = scales.detach().clone(), shifts.detach().clone()
    losses = []
    for i, img in enumerate([1, 1, 2, 0, 2]):                                                     # 2 eager steps, capture, 2 replays
        if i == 2:
            with torch.no_grad():                                                                 # an eval render right before the capture
                R_.render_rays(rb[:64], True, **kw)
        before_s = scales.detach().clone()
        losses.append(float(step(rb, target_s, target_h, img)["loss"]))
        moved = (scales.detach() - before_s).abs().reshape(-1) > 0
        # Adam moves every row that has ever had a gradient (momentum); rows never selected stay put
        seen = sorted(set([1, 1, 2, 0, 2][:i + 1]))
        assert all(bool(moved[j]) == (j in seen) for j in range(n_img)), (i, img, moved.tolist())
        g = scales.grad.detach().reshape(-1)
        assert float(g[img].abs()) > 0 and all(float(g[j]) == 0.0 for j in range(n_img) if j != img), (i, img, g.tolist())
    assert step.graph is not None and np.isfinite(losses).all()
    # (2) the captured graph contains the two weight re-packs (fast stream: 2 launches per net)
    assert step.launches_per_step is not None and step.launches_per_step >= 4
    w_before = nets[1].pts_linears[3].weight.detach().clone()
    l_a = float(step(rb, target_s, target_h, 1)["loss"])
    l_b = float(step(rb, target_s, target_h, 1)["loss"])
    assert not torch.equal(w_before, nets[1].pts_linears[3].weight.detach()) and l_a != l_b      # replays see the updated weights
    # (3) checkpoint round trip
    import copy
    sd = copy.deepcopy(opt.state_dict())                    # what torch.save / torch.load do (state_dict() aliases the live moments)
    steps = {int(v["step"]) for v in sd["state"].values()}
    assert steps == {7} and opt._step == 7 and int(opt._step_t) == 7
    m_saved = opt._m.clone()
    opt._m.zero_(); opt._v.mul_(0.5); opt._step = 3
    opt.load_state_dict(sd)
    assert opt._step == 7 and int(opt._step_t) == 7 and torch.equal(opt._m, m_saved)
    assert all(st["exp_avg"].data_ptr() >= opt._m.data_ptr() for st in opt.state.values())        # state views alias the flat moments
    step.release()


@pytest.mark.parametrize("hyp_full", [False, True])
def test_space_carving_joint_sharded_matches_full_batch(dev, hyp_full):
    """is_joint=True (H:115-119) on a ray-sharded step (SURVEY 8(e) "Exception"): per-shard [K,P] sums, their sum over the
    shards (the all-reduce, done by hand here), then per-shard finish with the GLOBAL ray count == the full-batch loss and
    the corresponding slices of its gradient; and both == the oracle."""
    from scade_b200 import _lib, functional as F_
    L = _lib.load()
    rng = np.random.default_rng(3)
    N, P, K = 1000, 96, 7
    pred = rng.uniform(0.5, 4.5, (N, P)).astype(np.float32)
    hyp = rng.uniform(0.1, 5.0, (K, N, P if hyp_full else 1)).astype(np.float32)
    mask = (rng.random(N) > 0.2).astype(np.float32)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    p_t, h_t, m_t = T(pred).requires_grad_(True), T(hyp).requires_grad_(True), T(mask)
    full = F_.space_carving_loss(p_t, h_t, is_joint=True, mask=m_t, threshold=0.05)
    full.backward()
    ref = O.space_carving_loss(pred, hyp, True, mask, 0.05)
    np.testing.assert_allclose(float(full), float(ref), rtol=2e-5)
    cuts = [(0, 333), (333, 1000)]                                                               # two "ranks"
    qs, parts = [], []
    for lo, hi in cuts:
        pp, hh, mm = T(pred[lo:hi]), T(hyp[:, lo:hi]), T(mask[lo:hi])
        q = torch.empty((K, P), dtype=torch.float32, device=dev)
        _lib.check(L.scade_space_carving_joint_accumulate(_lib.ptr(pp), _lib.ptr(hh), int(hyp_full), _lib.ptr(mm), K, hi - lo, P, 0.05,
                                                          _lib.ptr(q), _lib.stream_ptr()), "accumulate")
        qs.append(q)
        parts.append((pp, hh, mm))
    qsum = qs[0] + qs[1]                                                                          # dist.all_reduce(SUM)
    for (lo, hi), (pp, hh, mm) in zip(cuts, parts):
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        d_p, d_h = torch.empty_like(pp), torch.empty_like(hh)
        ks = torch.empty(P, dtype=torch.int32, device=dev)
        _lib.check(L.scade_space_carving_joint_finish(_lib.ptr(pp), _lib.ptr(hh), int(hyp_full), _lib.ptr(mm), _lib.ptr(qsum), K, hi - lo, N,
                                                      P, 0.05, 1.0, _lib.ptr(loss), _lib.ptr(d_p), _lib.ptr(d_h), _lib.ptr(ks),
                                                      _lib.stream_ptr()), "finish")
        np.testing.assert_allclose(float(loss), float(full), rtol=1e-5)
        np.testing.assert_allclose(d_p.cpu().numpy(), p_t.grad[lo:hi].cpu().numpy(), rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(d_h.cpu().numpy(), h_t.grad[:, lo:hi].cpu().numpy(), rtol=1e-5, atol=1e-10)
    # world-1 autograd wrapper == the one-call joint loss
    p2, h2 = T(pred).requires_grad_(True), T(hyp).requires_grad_(True)
    l2 = F_.space_carving_loss_joint_sharded(p2, h2, N, mask=m_t, threshold=0.05)
    l2.backward()
    np.testing.assert_allclose(float(l2.detach()), float(full.detach()), rtol=1e-6)    # the [K,P] sums are float atomics: order varies
    torch.testing.assert_close(p2.grad, p_t.grad, rtol=1e-6, atol=0)


def test_space_carving_affine_matches_composed(dev):
    """RS:954 fused into the loss kernel (scade_space_carving_loss_affine): same loss bits and d_pred bits as
    compute_space_carving_loss(pred, target_h * scale + shift), d_scale / d_shift equal autograd's reductions to fp32 round-off;
    `denominator` rescales like the ray-sharded step needs; both against the oracle."""
    from scade_b200 import functional as F_
    rng = np.random.default_rng(8)
    N, P, K = 777, 128, 20
    pred = rng.uniform(0.3, 4.8, (N, P)).astype(np.float32)
    hyp = rng.uniform(0.1, 5.0, (K, N, 1)).astype(np.float32)
    mask = (rng.random(N) > 0.1).astype(np.float32)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for thr in (0.0, 0.03):
        p1, s1, h1 = T(pred).requires_grad_(True), torch.tensor([1.07], device=dev, requires_grad=True), torch.tensor([-0.04], device=dev, requires_grad=True)
        l1 = F_.space_carving_loss(p1, T(hyp) * s1 + h1, mask=T(mask), threshold=thr)
        l1.backward()
        p2, s2, h2 = T(pred).requires_grad_(True), torch.tensor([1.07], device=dev, requires_grad=True), torch.tensor([-0.04], device=dev, requires_grad=True)
        l2 = F_.space_carving_loss_affine(p2, T(hyp), s2, h2, mask=T(mask), threshold=thr)
        l2.backward()
        np.testing.assert_allclose(float(l1.detach()), float(l2.detach()), rtol=1e-6)      # block partials meet in float atomics
        assert torch.equal(p1.grad, p2.grad)
        np.testing.assert_allclose(float(s2.grad), float(s1.grad), rtol=2e-5)
        np.testing.assert_allclose(float(h2.grad), float(h1.grad), rtol=2e-5, atol=1e-9)
        ref = O.space_carving_loss(pred, hyp * np.float32(1.07) + np.float32(-0.04), False, mask, thr)
        np.testing.assert_allclose(float(l2.detach()), float(ref), rtol=2e-6)
        d_pred_ref, d_h_ref = O.space_carving_loss_bwd(pred, hyp * np.float32(1.07) + np.float32(-0.04), False, mask, thr)
        np.testing.assert_allclose(float(s2.grad), float((d_h_ref * hyp).sum()), rtol=1e-4)
        # a ray shard normalised by the global count
        p3 = T(pred[:300]).requires_grad_(True)
        l3 = F_.space_carving_loss_affine(p3, T(hyp[:, :300]), s2.detach(), h2.detach(), mask=T(mask[:300]), threshold=thr, denominator=N)
        l3f = F_.space_carving_loss(T(pred[:300]), T(hyp[:, :300]) * 1.07 - 0.04, mask=T(mask[:300]), threshold=thr)
        np.testing.assert_allclose(float(l3.detach()), float(l3f) * 300 / N, rtol=1e-6)


@pytest.mark.parametrize("N", [300, 9001])
def test_space_carving_both_block_shapes_vs_oracle(dev, N):
    """The default branch picks 8 rays per block for small batches and 32 (hypotheses staged along N) from 8192 rays on:
    both against the oracle, forward and backward, with a mask and ragged last blocks."""
    from scade_b200 import functional as F_
    rng = np.random.default_rng(N)
    P, K = 48, 6
    pred = rng.uniform(0.3, 4.8, (N, P)).astype(np.float32)
    hyp = rng.uniform(0.1, 5.0, (K, N, 1)).astype(np.float32)
    hyp[1] = hyp[0]                                                                  # duplicates: ties pick the first k
    mask = (rng.random(N) > 0.15).astype(np.float32)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    p_t, h_t = T(pred).requires_grad_(True), T(hyp).requires_grad_(True)
    loss = F_.space_carving_loss(p_t, h_t, mask=T(mask), threshold=0.02)
    loss.backward()
    np.testing.assert_allclose(float(loss.detach()), float(O.space_carving_loss(pred, hyp, False, mask, 0.02)), rtol=3e-6)
    d_pred, d_hyp = O.space_carving_loss_bwd(pred, hyp, False, mask, 0.02)
    np.testing.assert_allclose(p_t.grad.cpu().numpy(), d_pred, rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(h_t.grad.cpu().numpy(), d_hyp, rtol=1e-4, atol=1e-10)
    assert float(h_t.grad[1].abs().max()) == 0.0                                     # the duplicate never wins
