"""GPU parity at the sizes that are benchmarked: every precision mode of the field network against the fp32 oracle

  * at the BASELINE metric shape (128 coarse + 128 importance samples, two 8x256 nets) on the first 512 rays of the
    very ray batch bench.py times (synthetic.make_ray_batch(4096, seed=50)), and
  * on a BASELINE config 3 train step (8x256 nets, 64 coarse + 128 importance, K = 20 hypotheses, 1024 rays):
    losses and per-tensor gradients against oracle.train_loss_and_grads (RS:954-985).

Two kinds of comparison, both with written tolerances (PARITY_TOL below, also tabulated in DESIGN.md section 4):

  end to end       the whole render_rays chain against the oracle's.  The chain 2^8*pi encoding -> 8-layer MLP ->
                   inverse-CDF resampling amplifies any rounding of the coarse pass into O(1e-3) shifts of a few fine
                   sample positions, so fine-pass quantities are stated as mean errors / PSNR.
  teacher-forced   the oracle's fine pass evaluated at the merged sample positions z_vals the CUDA path produced
                   (oracle.render_rays(z_fine=...)): isolates the fine network, its compositing, the second resampling
                   and all gradients from the discontinuous resampling step, so max-abs tolerances apply.

The measured errors are printed (pytest -s) and recorded in profiles/; tolerances are ~2-3x the measured values.
"""
import numpy as np
import pytest
import torch

from oracle import scade_oracle as O
from scade_b200 import synthetic as syn
from tests.golden.generate_goldens import net_pair

pytestmark = pytest.mark.gpu

# (max abs, mean abs) per output and precision mode.  fp32: summation-order noise of two correct fp32 implementations.
# tc_f16: single-pass fp16 operands (11-bit mantissa).  tc_f16x3: fp16 hi/lo split operands, three tensor-core passes.
PARITY_TOL = {
    # measured on B200 (profiles/r02_parity_report.md): rgb0 1.5e-5/2.9e-7, raw 6.6e-5/7.0e-6, rgb_map 1.3e-6/2.1e-7, PSNR 75.6 dB
    "fp32": {
        "coarse": {"rgb0": (5e-5, 1e-6), "depth0": (3e-4, 6e-6), "acc0": (5e-5, 1e-6), "weights0": (5e-5, 1e-7)},
        "teacher": {"raw": (2e-4, 2e-5), "rgb_map": (5e-6, 1e-6), "depth_map": (5e-6, 1e-6), "acc_map": (2e-6, 5e-7),
                    "weights": (2e-6, 5e-8), "pred_hyp": (0.1, 5e-5)},
        "e2e": {"rgb_psnr": 70.0, "depth_mean": 3e-5, "pred_hyp_mean": 1e-4},
    },
    # measured: rgb0 5.8e-5/1.3e-6, raw 3.9e-4/5.9e-5, rgb_map 4.0e-6/7.5e-7, depth_map 7.9e-6/2.6e-6, PSNR 75.9 dB (the fp32 path
    # itself: 75.6 dB -- the end-to-end figure is the workload's discontinuity floor, not the arithmetic).  The operand split
    # alone would give raw 7e-5 (oracle.nerf_forward_f16(split=True)); the rest is the tensor core's fp32 accumulator, which
    # truncates (round-toward-zero) at each of the 48 accumulation steps of a 256-wide layer.
    "tc_f16x3": {
        "coarse": {"rgb0": (1.5e-4, 4e-6), "depth0": (1e-3, 3e-5), "acc0": (2e-4, 5e-6), "weights0": (2e-4, 2e-7)},
        "teacher": {"raw": (1e-3, 1.5e-4), "rgb_map": (1.2e-5, 2e-6), "depth_map": (2.5e-5, 8e-6), "acc_map": (2e-6, 5e-7),
                    "weights": (5e-6, 1e-7), "pred_hyp": (0.1, 5e-5)},
        "e2e": {"rgb_psnr": 70.0, "depth_mean": 6e-5, "pred_hyp_mean": 1e-4},
    },
    # measured: rgb0 1.7e-2/3.9e-4, raw 7.1e-2/8.5e-3, rgb_map 1.5e-3/1.5e-4, depth_map 2.3e-3/4.4e-4, PSNR 49.4 dB
    "tc_f16": {
        "coarse": {"rgb0": (4e-2, 1.2e-3), "depth0": (0.2, 8e-3), "acc0": (4e-2, 1.5e-3), "weights0": (4e-2, 6e-5)},
        "teacher": {"raw": (0.2, 2.5e-2), "rgb_map": (5e-3, 5e-4), "depth_map": (8e-3, 1.5e-3), "acc_map": (1e-4, 1e-6),
                    "weights": (2e-3, 2e-5), "pred_hyp": (1.5, 1.5e-3)},
        "e2e": {"rgb_psnr": 44.0, "depth_mean": 5e-3, "pred_hyp_mean": 6e-3},
    },
}

# train step: relative tolerances on the losses, on d_scale / d_shift (sums of signed terms) and per-tensor gradient error
# max|g - g_ref| / max|g_ref|.  Measured on B200: fp32 loss 1.3e-7, sc 3.1e-6, d_scale 9.1e-4, coarse 3.9e-5, fine 5.7e-4;
# tc_f16 loss 5.8e-5, sc 1.0e-4, d_scale 9.5e-4, coarse 3.2e-2 (pts_linears.0.weight), fine 1.4e-2 (alpha_linear.weight).
TRAIN_TOL = {
    "fp32": {"loss": 2e-6, "sc": 2e-5, "ss": 3e-3, "coarse": 2e-4, "fine": 2e-3},
    "tc_f16": {"loss": 5e-4, "sc": 5e-4, "ss": 5e-3, "coarse": 8e-2, "fine": 4e-2},
}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from scade_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def npy(t):
    return t.detach().cpu().numpy()


def psnr(a, b):
    return -10.0 * np.log10(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2) + 1e-30)


def make_kwargs(dev, precision, perturb, Nc, Nf, requires_grad=False):
    from scade_b200 import nerf_helpers as NH
    from scade_b200.render import NetworkQuery
    pc, pf = net_pair(8, 256)
    bb_center, bb_scale = syn.bounding_box()

    def mk(p):
        net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision=precision)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
        return net.to(dev).requires_grad_(requires_grad)
    qf = NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision=precision)
    kwargs = dict(network_fn=mk(pc), network_query_fn=qf, N_samples=Nc, embedded_cam=torch.tensor((), device=dev), retraw=True,
                  perturb=perturb, N_importance=Nf, network_fine=mk(pf), raw_noise_std=0.0)
    return kwargs, (pc, pf, bb_center, bb_scale)


def _check(report, group, key, got, ref, tol):
    d = np.abs(np.asarray(got, np.float64) - np.asarray(ref, np.float64))
    assert np.isfinite(d).all(), f"{group}/{key}: non-finite difference"
    report.append(f"  {group:8s} {key:10s} max {d.max():.3e} (tol {tol[0]:.1e})   mean {d.mean():.3e} (tol {tol[1]:.1e})")
    return d.max() <= tol[0] and d.mean() <= tol[1]


@pytest.mark.parametrize("precision", ["fp32", "tc_f16x3", "tc_f16"])
def test_render_metric_shape_vs_oracle(dev, precision):
    """BASELINE metric shape, the benchmarked ray batch: coarse pass, teacher-forced fine pass and the end-to-end chain."""
    from scade_b200 import render as R_
    n, Nc, Nf = 512, 128, 128
    kwargs, (pc, pf, bb_center, bb_scale) = make_kwargs(dev, precision, 0.0, Nc, Nf)
    rb = syn.make_ray_batch(4096, seed=50)[:n]
    with torch.no_grad():
        ret = R_.render_rays(T(rb, dev), True, **kwargs)
    ret = {k: npy(v) for k, v in ret.items()}
    ref = O.render_rays(rb, pc, pf, bb_center, bb_scale, Nc, Nf)
    ref_tf = O.render_rays(rb, pc, pf, bb_center, bb_scale, Nc, Nf, z_fine=ret["z_vals"], retraw=True)
    tol = PARITY_TOL[precision]
    report, ok = [f"render parity, {precision}, {n} rays x ({Nc}c+{Nf}f):"], True
    np.testing.assert_array_equal(ret["z_vals0"], ref["z_vals0"])                    # sample placement is bit-exact in every mode
    for k, t in tol["coarse"].items():
        ok &= _check(report, "coarse", k, ret[k], ref[k], t)
    for k, t in tol["teacher"].items():
        ok &= _check(report, "teacher", k, ret[k], ref_tf[k], t)
    e = tol["e2e"]
    p = psnr(ret["rgb_map"], ref["rgb_map"])
    dm = np.abs(ret["depth_map"] - ref["depth_map"]).mean()
    hm = np.abs(ret["pred_hyp"] - ref["pred_hyp"]).mean()
    report.append(f"  e2e      rgb PSNR {p:.1f} dB (>= {e['rgb_psnr']})   depth mean {dm:.3e} (tol {e['depth_mean']:.1e})   "
                  f"pred_hyp mean {hm:.3e} (tol {e['pred_hyp_mean']:.1e})")
    ok &= p >= e["rgb_psnr"] and dm <= e["depth_mean"] and hm <= e["pred_hyp_mean"]
    print("\n" + "\n".join(report))
    assert ok, "\n".join(report)


@pytest.mark.parametrize("precision", ["fp32", "tc_f16"])
def test_train_step_c3_vs_oracle(dev, precision):
    """BASELINE config 3 (8x256 nets, 64c+128f, K=20) at 1024 rays: loss values and every gradient tensor against
    oracle.train_loss_and_grads, the oracle's fine pass teacher-forced to the CUDA path's merged sample positions."""
    from scade_b200 import nerf_helpers as NH
    from scade_b200 import render as R_
    n, Nc, Nf, K = 1024, 64, 128, 20
    kwargs, (pc, pf, bb_center, bb_scale) = make_kwargs(dev, precision, 1.0, Nc, Nf, requires_grad=True)
    kwargs["retraw"] = False
    rb = syn.make_ray_batch(n, seed=80)
    t_rand, u_c, u_f = syn.make_uniforms(n, Nc, Nf, seed=81)
    target_s, target_h = syn.make_train_targets(n, K=K, seed=82)
    scale = torch.tensor([1.1], device=dev, requires_grad=True)
    shift = torch.tensor([-0.05], device=dev, requires_grad=True)
    th = T(target_h, dev) * scale + shift                                            # RS:954
    ret = R_.render_rays(T(rb, dev), True, cached_u=T(u_f, dev), t_rand=T(t_rand, dev), u_coarse=T(u_c, dev), **kwargs)
    img_loss = NH.img2mse(ret["rgb_map"], T(target_s, dev))                          # RS:968
    sc = NH.compute_space_carving_loss(ret["pred_hyp"], th, is_joint=False, norm_p=2, threshold=0.0)     # RS:974
    img_loss0 = NH.img2mse(ret["rgb0"], T(target_s, dev))                            # RS:981
    loss = img_loss + 0.007 * sc + img_loss0                                         # RS:976,983
    loss.backward()                                                                  # RS:985
    torch.cuda.synchronize()
    losses, gc, gf, d_scale, d_shift, _ = O.train_loss_and_grads(
        rb, pc, pf, bb_center, bb_scale, Nc, Nf, target_s, target_h, t_rand, u_c, u_f, scale=1.1, shift=-0.05,
        z_fine=npy(ret["z_vals"]))
    tol = TRAIN_TOL[precision]
    report, ok = [f"train-step parity, {precision}, {n} rays x ({Nc}c+{Nf}f), K={K}:"], True

    def rel_scalar(name, got, ref, t):
        r = abs(float(got) - float(ref)) / (abs(float(ref)) + 1e-30)
        report.append(f"  {name:22s} got {float(got):.7g} ref {float(ref):.7g} rel {r:.2e} (tol {t:.1e})")
        return r <= t
    ok &= rel_scalar("img_loss0", img_loss0.item(), losses["img_loss0"], tol["loss"])
    ok &= rel_scalar("img_loss", img_loss.item(), losses["img_loss"], tol["loss"])
    ok &= rel_scalar("space_carving", sc.item(), losses["space_carving"], tol["sc"])
    ok &= rel_scalar("loss", loss.item(), losses["loss"], tol["loss"])
    ok &= rel_scalar("d_scale", scale.grad.item(), d_scale, tol["ss"])
    ok &= rel_scalar("d_shift", shift.grad.item(), d_shift, tol["ss"])
    for which, net, refg in (("coarse", kwargs["network_fn"], gc), ("fine", kwargs["network_fine"], gf)):
        worst, worst_name = 0.0, ""
        for name, p in net.named_parameters():
            r = refg[name]
            err = float(np.abs(npy(p.grad) - r).max() / (np.abs(r).max() + 1e-30))
            if err > worst:
                worst, worst_name = err, name
        report.append(f"  {which:6s} net: worst gradient tensor {worst_name} rel err {worst:.2e} (tol {tol[which]:.1e})")
        ok &= worst <= tol[which]
    print("\n" + "\n".join(report))
    assert ok, "\n".join(report)


@pytest.mark.parametrize("P", [700, 33000])
def test_mlp_x3_vs_oracle(dev, P):
    """The tight tensor-core mode on identical inputs (NeRF.forward, H:223-247): against the fp32 oracle (float64-accumulated)
    and against the CPU emulation of the same hi/lo operand split.  P = 700 leaves a ragged last 256-point step and fewer
    steps than SM pairs; 33000 gives every SM pair several steps (ring wrap-around, accumulator ping-pong across steps).
    Tolerance: max |raw - oracle| <= 1e-3, mean <= 1.5e-4 -- two orders below single-pass fp16 (7e-2 / 8e-3); the operand
    split alone gives 7e-5 in emulation, the rest is the tensor core's truncating fp32 accumulator."""
    from scade_b200 import functional as F_, nerf_helpers as NH
    params = syn.make_nerf_params(seed=10, D=8, W=256, bias_scale=0.05, alpha_bias=0.5, weight_gain=1.3)
    net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16x3")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    net = net.to(dev).requires_grad_(False)
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, (P, 60)).astype(np.float32)
    with torch.no_grad():
        out = npy(net(T(x, dev)))
    ref = O.nerf_forward(params, x, dtype=np.float64).astype(np.float32)
    d = np.abs(out - ref)
    print(f"\nmlp x3, P={P}: max |raw - fp64 oracle| = {d.max():.3e}, mean {d.mean():.3e}")
    assert np.isfinite(out).all()
    assert d.max() <= 1e-3 and d.mean() <= 1.5e-4, (d.max(), d.mean())
    if P <= 1000:
        emu = O.nerf_forward_f16(params, x, split=True)
        de = np.abs(out - emu)
        print(f"mlp x3, P={P}: max |raw - hi/lo emulation| = {de.max():.3e}")
        assert de.max() <= 1e-3
    # rays mode (positional encoding in the prologue) agrees with the embedded mode on the same points
    if P <= 1000:
        bb_center, bb_scale = syn.bounding_box()
        rb = syn.make_ray_batch(5, seed=3)
        z = np.linspace(0.1, 5.0, 140, dtype=np.float32)[None, :].repeat(5, 0)
        raw = npy(F_.mlp_forward_rays(net.handle(), T(rb, dev), T(z, dev), bb_center, bb_scale, "tc_f16x3"))
        pts = rb[:, None, 0:3] + rb[:, None, 3:6] * z[:, :, None]
        xin = O.network_inputs(pts, rb[:, 8:11], bb_center, bb_scale)
        ref_r = O.nerf_forward(params, xin.reshape(-1, 60), dtype=np.float64).astype(np.float32).reshape(5, 140, 4)
        dr = np.abs(raw - ref_r)
        print(f"mlp x3 rays mode: max |raw - oracle| = {dr.max():.3e}")
        assert dr.max() <= 1e-3


@pytest.mark.parametrize("S,N", [(64, 37), (128, 19), (256, 5), (256, 300), (128, 1), (32, 77),
                                 (192, 1), (192, 5), (192, 4096), (192, 1237), (96, 50), (320, 9), (512, 3), (224, 611)])
def test_fused_compositing_is_bit_identical(dev, S, N):
    """north_star: the cumprod alpha-composite fused into the kernel of the last GEMM.  scade_mlp_forward_rays_composite
    (one kernel: raw never reaches memory) against scade_mlp_forward_rays + scade_raw2outputs (RS:659-660, RS:511-562):
    every output bit-identical, for 1, 2, 4 and 8 warps per ray (whole rays per 256-point step), for the sample counts that
    straddle tiles and steps (192 = the reference config's 64 + 128, 96, 224, 320, 512: the kernel's chain mode), ragged last
    tiles and ranges, and retraw on / off."""
    from scade_b200 import functional as F_, nerf_helpers as NH
    params = syn.make_nerf_params(seed=11, D=8, W=256, bias_scale=0.05, alpha_bias=0.5, weight_gain=1.3)
    net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    net = net.to(dev).requires_grad_(False)
    h = net.handle()
    assert F_.composite_fusable(h, "tc_f16", S) and not F_.composite_fusable(h, "tc_f16", 200) and not F_.composite_fusable(h, "fp32", S)
    bb_center, bb_scale = syn.bounding_box()
    rb = T(syn.make_ray_batch(N, seed=60 + S), dev)
    rng = np.random.default_rng(S + N)
    z = np.sort(rng.uniform(0.1, 5.0, (N, S)).astype(np.float32), -1)
    z_t = T(z, dev)
    with torch.no_grad():
        raw = F_.mlp_forward_rays(h, rb, z_t, bb_center, bb_scale, "tc_f16")
        ref = F_.raw2outputs(raw, z_t, rb[:, 3:6].contiguous())                     # (rgb, disp, acc, weights, depth)
        for retraw in (False, True):
            rgb, disp, acc, w, depth, raw2 = F_.mlp_forward_rays_composite(h, rb, z_t, bb_center, bb_scale, "tc_f16", retraw=retraw)
            torch.cuda.synchronize()
            for name, a, b in (("weights", w, ref[3]), ("rgb_map", rgb, ref[0]), ("depth_map", depth, ref[4]), ("acc_map", acc, ref[2])):
                np.testing.assert_array_equal(npy(a), npy(b), err_msg=f"{name} S={S} N={N} retraw={retraw}")
            np.testing.assert_array_equal(npy(disp), npy(ref[1]))                   # NaN where acc == 0 on both sides (RS:559)
            if retraw:
                np.testing.assert_array_equal(npy(raw2), npy(raw))
    # and against the oracle's compositing of the same raw values (fp32 tolerance)
    o = O.raw2outputs(npy(raw), z, npy(rb)[:, 3:6])
    np.testing.assert_allclose(npy(w), o[3], rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(npy(rgb), o[0], rtol=2e-5, atol=2e-6)
