"""Generate golden fixtures by EXECUTING THE UNMODIFIED REFERENCE (mikacuy/scade) on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/generate_goldens.py

The reference has no tests or golden vectors of its own, so these files are the pin
for ``oracle/scade_oracle.py`` and, through it, for the CUDA path.  Inputs come from
``scade_b200.synthetic`` (numpy, seeded) so the tests can rebuild them on the GPU box where
/root/reference does not exist; only the reference's OUTPUTS (and tiny inputs) are stored.

Import recipe: SURVEY §8(c) -- stub the third-party modules the hot path never touches.
"""
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
import torchvision  # noqa: F401  (must be imported before the stubs)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("SCADE_REFERENCE", "/root/reference")


def import_reference():
    for name in ["configargparse", "skimage", "skimage.metrics", "skimage.io", "lpips", "imageio", "pandas"]:
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__path__ = []
        sys.modules[name] = m
    if not hasattr(sys.modules["skimage.metrics"], "structural_similarity"):
        sys.modules["skimage.metrics"].structural_similarity = None
    if not hasattr(sys.modules["lpips"], "LPIPS"):
        sys.modules["lpips"].LPIPS = None
    sys.path.insert(0, REF)
    import run_scade_scannet as R
    import model.run_nerf_helpers as H
    R.device = torch.device("cpu")
    return R, H


from scade_b200 import synthetic as syn  # noqa: E402


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def build_ref_nerf(H, params, D, W, input_ch=57, input_ch_views=3):
    net = H.NeRF(D=D, W=W, input_ch=input_ch, output_ch=5, skips=[4], input_ch_views=input_ch_views,
                 input_ch_cam=0, use_viewdirs=True)
    net.load_state_dict({k: t(v) for k, v in params.items()})
    return net


def make_query_fn(R, H, bb_center, bb_scale, multires=9, multires_views=0):
    embed_fn, _ = H.get_embedder(multires, 0)
    embeddirs_fn, _ = H.get_embedder(multires_views, 0)
    return lambda inputs, viewdirs, embedded_cam, network_fn: R.run_network(
        inputs, viewdirs, embedded_cam, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn,
        bb_center=t(bb_center), bb_scale=float(bb_scale), netchunk=1024 * 64)


class InjectRand:
    """Feed explicit uniforms to the reference's torch.rand_like (RS:570) and torch.rand (H:350)."""

    def __init__(self, t_rand, u_coarse):
        self.t_rand, self.u_coarse = t(t_rand), t(u_coarse)

    def __enter__(self):
        self._rand, self._rand_like = torch.rand, torch.rand_like
        torch.rand_like = lambda x, *a, **k: self.t_rand.clone()
        torch.rand = lambda *a, **k: self.u_coarse.clone()
        return self

    def __exit__(self, *exc):
        torch.rand, torch.rand_like = self._rand, self._rand_like


RENDER_CASES = {
    # name: (n_rays, Nc, Nf, D, W, perturb)
    "render_det_64c128f": (96, 64, 128, 8, 256, 0.0),
    "render_perturb_64c128f": (96, 64, 128, 8, 256, 1.0),
    "render_det_128c128f": (48, 128, 128, 8, 256, 0.0),
    "render_det_small_net": (64, 32, 64, 4, 128, 0.0),
}


def net_pair(D, W):
    pc = syn.make_nerf_params(seed=10, D=D, W=W, bias_scale=0.05, alpha_bias=0.5, weight_gain=1.3)
    pf = syn.make_nerf_params(seed=11, D=D, W=W, bias_scale=0.05, alpha_bias=0.5, weight_gain=1.3)
    return pc, pf


def main():
    R, H = import_reference()
    torch.set_grad_enabled(False)
    out = {}

    # ---- a5 embedder ------------------------------------------------------------------
    x = np.random.default_rng(0).uniform(-1, 1, (64, 3)).astype(np.float32)
    x[0] = [0.5, -0.25, 0.125]
    emb9, dim9 = H.get_embedder(9, 0)
    emb0, dim0 = H.get_embedder(0, 0)
    np.savez_compressed(os.path.join(HERE, "embed.npz"), x=x, emb9=emb9(t(x)).numpy(), emb0=emb0(t(x)).numpy(),
                        dim9=dim9, dim0=dim0)

    # ---- a6 NeRF forward --------------------------------------------------------------
    nf = {}
    for tag, (D, W) in {"d8w256": (8, 256), "d4w128": (4, 128)}.items():
        params = syn.make_nerf_params(seed=3, D=D, W=W, bias_scale=0.1, alpha_bias=0.3)
        net = build_ref_nerf(H, params, D, W)
        xin = np.random.default_rng(4).uniform(-1, 1, (200, 60)).astype(np.float32)
        nf[f"{tag}_x"] = xin
        nf[f"{tag}_out"] = net(t(xin)).numpy()
    sp_in = np.array([-1, 0, .5, 1.99, 2.0, 2.01, 3], np.float32)
    nf["softplus_in"] = sp_in
    nf["softplus_out"] = torch.nn.functional.softplus(t(sp_in), beta=10).numpy()
    np.savez_compressed(os.path.join(HERE, "nerf_forward.npz"), **nf)

    # ---- a4 run_network ---------------------------------------------------------------
    bb_center, bb_scale = syn.bounding_box()
    rb = syn.make_ray_batch(32, seed=5)
    zz = np.linspace(0.1, 5.0, 16, dtype=np.float32)[None, :].repeat(32, 0)
    pts = rb[:, None, 0:3] + rb[:, None, 3:6] * zz[:, :, None]
    params = syn.make_nerf_params(seed=3, D=8, W=256, bias_scale=0.1, alpha_bias=0.3)
    net = build_ref_nerf(H, params, 8, 256)
    qf = make_query_fn(R, H, bb_center, bb_scale)
    raw = qf(t(pts), t(rb[:, 8:11]), torch.tensor(()), net).numpy()
    np.savez_compressed(os.path.join(HERE, "run_network.npz"), z=zz, raw=raw)

    # ---- a7/a8 raw2outputs (BASELINE config 1: 64 rays x 64 samples) ---------------------
    rng = np.random.default_rng(0)
    raw = rng.standard_normal((64, 64, 4)).astype(np.float32)
    raw[..., 3] = np.abs(raw[..., 3]) * 3.0          # softplus'd sigma is non-negative
    raw[5, :, 3] = 0.0                               # an empty ray: acc == 0 -> disp nan (RS:559)
    z = np.linspace(0.1, 5.0, 64, dtype=np.float32)[None, :].repeat(64, 0)
    rb = syn.make_ray_batch(64, seed=6)
    res = R.raw2outputs(t(raw), t(z), t(rb[:, 3:6]))
    r2o = dict(raw=raw, z=z, rgb_map=res[0].numpy(), disp_map=res[1].numpy(), acc_map=res[2].numpy(),
               weights=res[3].numpy(), depth_map=res[4].numpy())
    # with sigma noise through the reference's pytest hook (RS:548-552): noise = np.random.rand * std
    resn = R.raw2outputs(t(raw), t(z), t(rb[:, 3:6]), raw_noise_std=0.5, pytest=True)
    np.random.seed(0)
    r2o["noise"] = (np.random.rand(64, 64) * 0.5).astype(np.float32)
    r2o["n_rgb_map"], r2o["n_weights"], r2o["n_depth_map"] = resn[0].numpy(), resn[3].numpy(), resn[4].numpy()
    # KAT of SURVEY Appendix B
    kraw = np.array([[[0, 1, -1, .5], [2, 0, 0, 1], [0, 0, 3, 0], [1, 1, 1, 2]]], np.float32)
    kres = R.raw2outputs(t(kraw), t(np.array([[1, 2, 3, 4]], np.float32)), t(np.array([[0, 0, 2]], np.float32)))
    r2o["kat_raw"] = kraw
    for n, v in zip(["rgb", "disp", "acc", "w", "depth"], kres):
        r2o["kat_" + n] = v.numpy()
    # backward of raw2outputs through the reference autograd
    with torch.enable_grad():
        rt = t(raw[:16, :24].copy()).requires_grad_(True)
        zt, dt = t(z[:16, :24].copy()), t(rb[:16, 3:6])
        o = R.raw2outputs(rt, zt, dt)
        gr = np.random.default_rng(7)
        gs = [gr.standard_normal(tuple(v.shape)).astype(np.float32) for v in o]
        # ray 5 has acc == 0 -> nan disp; keep its disp gradient out of the sum
        gs[1][5] = 0.0
        gs[1][:] *= 0.01
        tot = sum((v * t(g)).sum() for i, (v, g) in enumerate(zip(o, gs)) if not (i == 1))
        disp_ok = torch.where(torch.isnan(o[1]), torch.zeros_like(o[1]), o[1])
        tot = tot + (disp_ok * t(gs[1])).sum()
        tot.backward()
    r2o["bwd_d_raw"] = rt.grad.numpy()
    for n, g in zip(["rgb", "disp", "acc", "w", "depth"], gs):
        r2o["bwd_g_" + n] = g
    np.savez_compressed(os.path.join(HERE, "raw2outputs.npz"), **r2o)

    # ---- a9 perturb -------------------------------------------------------------------
    zc = np.linspace(0.1, 5.0, 64, dtype=np.float32)[None, :].repeat(8, 0)
    pz = R.perturb_z_vals(t(zc), pytest=True).numpy()
    np.random.seed(0)
    tr = np.random.rand(8, 64).astype(np.float32)
    kat = R.perturb_z_vals(t(np.linspace(0, 1, 4, dtype=np.float32)), pytest=True).numpy()
    np.savez_compressed(os.path.join(HERE, "perturb.npz"), z=zc, t_rand=tr, out=pz, kat=kat)

    # ---- a10 sample_pdf ---------------------------------------------------------------
    sp = {}
    rng = np.random.default_rng(8)
    N, S = 40, 64
    zs = np.sort(rng.uniform(0.1, 5.0, (N, S)).astype(np.float32), -1)
    bins = 0.5 * (zs[:, 1:] + zs[:, :-1])
    w = rng.random((N, S - 2), dtype=np.float32) ** 4
    w[0] = 0.0                                  # all-zero weights -> uniform pdf
    w[1] = 0.0; w[1, 17] = 1.0                  # one-hot -> denom < 1e-5 branch on the flat parts
    w[2] = 0.0; w[2, 0] = 1.0
    w[3] = 0.0; w[3, -1] = 1.0
    u = rng.random((N, 48), dtype=np.float32)
    u[4, 0], u[4, 1] = 0.0, 1.0                 # endpoints (u == cdf[-1] -> below = above = last)
    sp.update(bins=bins, w=w, u=u)
    sp["det"] = H.sample_pdf(t(bins), t(w), 48, det=True).numpy()
    s_u, u_back = H.sample_pdf_return_u(t(bins), t(w), 48, det=False, load_u=t(u))
    sp["with_u"] = s_u.numpy()
    s_det_u, u_det = H.sample_pdf_return_u(t(bins), t(w), 33, det=True)
    sp["det33"], sp["u_det33"] = s_det_u.numpy(), u_det.numpy()
    uj = rng.random((48,), dtype=np.float32)
    sj, _ = H.sample_pdf_joint_return_u(t(bins), t(w), 48, load_u=t(uj)[None, :].repeat(N, 1))
    sp["u_joint"], sp["joint"] = uj, sj.numpy()
    kb = np.linspace(0, 1, 5, dtype=np.float32)[None]
    kw = np.array([[1, 2, 1, 0]], np.float32)
    sp["kat_det"] = H.sample_pdf(t(kb), t(kw), 6, det=True).numpy()
    sp["kat_u"] = H.sample_pdf_return_u(t(kb), t(kw), 4, load_u=t(np.array([[.1, .9, .5, .25]], np.float32)))[0].numpy()
    with torch.enable_grad():
        wt = t(w.copy()).requires_grad_(True)
        s, _ = H.sample_pdf_return_u(t(bins), wt, 48, load_u=t(u))
        gs = rng.standard_normal((N, 48)).astype(np.float32)
        (s * t(gs)).sum().backward()
    sp["bwd_g"], sp["bwd_d_w"] = gs, wt.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "sample_pdf.npz"), **sp)

    # ---- a12 space carving ------------------------------------------------------------
    sc = {}
    rng = np.random.default_rng(9)
    N, P, K = 50, 32, 20
    pred = rng.uniform(0.1, 5.0, (N, P)).astype(np.float32)
    hyp = rng.uniform(0.1, 5.0, (K, N, 1)).astype(np.float32)
    hyp_full = rng.uniform(0.1, 5.0, (K, N, P)).astype(np.float32)
    mask = (rng.random(N) > 0.3).astype(np.float32)
    pred[0, 0] = hyp[3, 0, 0]                   # exact hit: zero distance, zero gradient
    sc.update(pred=pred, hyp=hyp, hyp_full=hyp_full, mask=mask)
    cases = {"default": {}, "joint": dict(is_joint=True), "thr": dict(threshold=0.6), "mask": dict(mask=t(mask)),
             "joint_mask_thr": dict(is_joint=True, mask=t(mask), threshold=0.3)}
    for name, kw_ in cases.items():
        with torch.enable_grad():
            pt = t(pred.copy()).requires_grad_(True)
            ht = t(hyp.copy()).requires_grad_(True)
            L = H.compute_space_carving_loss(pt, ht, **kw_)
            L.backward()
        sc[name + "_loss"] = L.detach().numpy()
        sc[name + "_d_pred"], sc[name + "_d_hyp"] = pt.grad.numpy(), ht.grad.numpy()
    with torch.enable_grad():
        pt = t(pred.copy()).requires_grad_(True)
        ht = t(hyp_full.copy()).requires_grad_(True)
        L = H.compute_space_carving_loss(pt, ht)
        L.backward()
    sc["full_loss"], sc["full_d_pred"], sc["full_d_hyp"] = L.detach().numpy(), pt.grad.numpy(), ht.grad.numpy()
    kp = t(np.array([[1, 2], [3, 5]], np.float32))
    kh = t(np.array([[[1.5], [2.0]], [[0], [4.5]], [[2], [9]]], np.float32))
    sc["kat"] = np.array([H.compute_space_carving_loss(kp, kh).item(),
                          H.compute_space_carving_loss(kp, kh, is_joint=True).item(),
                          H.compute_space_carving_loss(kp, kh, threshold=0.6).item(),
                          H.compute_space_carving_loss(kp, kh, mask=t(np.array([1, 0], np.float32))).item()], np.float32)
    np.savez_compressed(os.path.join(HERE, "space_carving.npz"), **sc)

    # ---- a3 get_rays ------------------------------------------------------------------
    c2w = np.array([[0.8, -0.6, 0.0, 0.3], [0.6, 0.8, 0.0, -0.2], [0.0, 0.0, 1.0, 0.1]], np.float32)
    ro, rd = H.get_rays(6, 8, t(np.array([10.0, 11.0, 4.0, 3.0], np.float32)), t(c2w))
    kro, krd = H.get_rays(2, 3, t(np.array([100.0, 100.0, 1.5, 1.0], np.float32)), torch.eye(4)[:3])
    np.savez_compressed(os.path.join(HERE, "get_rays.npz"), c2w=c2w, rays_o=ro.numpy(), rays_d=rd.numpy(), kat_d=krd.numpy())

    # ---- a1 render_rays ---------------------------------------------------------------
    keys = ["rgb_map", "disp_map", "acc_map", "depth_map", "z_vals", "weights", "pred_hyp", "u", "rgb0", "disp0",
            "acc0", "depth0", "z_vals0", "weights0", "z_std"]
    for name, (n, Nc, Nf, D, W, perturb) in RENDER_CASES.items():
        pc, pf = net_pair(D, W)
        netc, netf = build_ref_nerf(H, pc, D, W), build_ref_nerf(H, pf, D, W)
        rb = syn.make_ray_batch(n, seed=20)
        t_rand, u_c, u_f = syn.make_uniforms(n, Nc, Nf, seed=21)
        kwargs = dict(network_fn=netc, network_query_fn=qf, N_samples=Nc, embedded_cam=torch.tensor(()),
                      retraw=True, perturb=perturb, N_importance=Nf, network_fine=netf, raw_noise_std=0.0)
        if perturb > 0:
            with InjectRand(t_rand, u_c):
                ret = R.render_rays(t(rb), True, cached_u=t(u_f), **kwargs)
        else:
            ret = R.render_rays(t(rb), True, **kwargs)
        d = {k: ret[k].numpy() for k in keys}
        d["raw_head"] = ret["raw"][:8].numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "acc mean", float(ret["acc_map"].mean()), "depth mean", float(ret["depth_map"].mean()))

    # ---- a13 train step: losses and gradients through the reference autograd --------------
    for tag, (n, Nc, Nf, D, W) in {"train_small_net": (48, 32, 64, 4, 64), "train_d8w256": (64, 64, 128, 8, 256)}.items():
        pc, pf = net_pair(D, W)
        netc, netf = build_ref_nerf(H, pc, D, W), build_ref_nerf(H, pf, D, W)
        rb = syn.make_ray_batch(n, seed=30)
        t_rand, u_c, u_f = syn.make_uniforms(n, Nc, Nf, seed=31)
        target_s, target_h = syn.make_train_targets(n, K=20, seed=32)
        with torch.enable_grad():
            scale = torch.tensor([1.1], requires_grad=True)
            shift = torch.tensor([-0.05], requires_grad=True)
            th = t(target_h) * scale + shift                                               # RS:954
            kwargs = dict(network_fn=netc, network_query_fn=qf, N_samples=Nc, embedded_cam=torch.tensor(()),
                          retraw=True, perturb=1.0, N_importance=Nf, network_fine=netf, raw_noise_std=0.0)
            with InjectRand(t_rand, u_c):
                ret = R.render_rays(t(rb), True, cached_u=t(u_f), **kwargs)
            img_loss = H.img2mse(ret["rgb_map"], t(target_s))                              # RS:968
            scl = H.compute_space_carving_loss(ret["pred_hyp"], th, is_joint=False, norm_p=2, threshold=0.0)
            img_loss0 = H.img2mse(ret["rgb0"], t(target_s))                                # RS:981
            loss = img_loss + 0.007 * scl + img_loss0                                      # RS:976,983
            loss.backward()                                                                # RS:985
        d = dict(loss=loss.item(), img_loss=img_loss.item(), img_loss0=img_loss0.item(), space_carving=scl.item(),
                 d_scale=scale.grad.numpy(), d_shift=shift.grad.numpy(), rgb_map=ret["rgb_map"].detach().numpy(),
                 pred_hyp=ret["pred_hyp"].detach().numpy())
        for pref, net in (("c.", netc), ("f.", netf)):
            for k, p in net.named_parameters():
                g = p.grad.numpy()
                if D * W <= 512:
                    d[pref + k] = g
                else:   # big net: norms + a strided subsample keep the fixture small
                    d[pref + k + ".l2"] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
                    d[pref + k + ".sum"] = np.float64(g.astype(np.float64).sum())
                    d[pref + k + ".sub"] = g.reshape(-1)[::97].copy()
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **d)
        print(tag, "loss", loss.item(), "sc", scl.item())

    # ---- same train step with the reference run in float64: removes fp32 chaos so that the oracle's
    #      analytic backward can be pinned tightly (tests compare against oracle(dtype=float64)) ----
    n, Nc, Nf, D, W = 48, 32, 64, 4, 64
    pc, pf = net_pair(D, W)
    netc, netf = build_ref_nerf(H, pc, D, W).double(), build_ref_nerf(H, pf, D, W).double()
    rb = syn.make_ray_batch(n, seed=30)
    t_rand, u_c, u_f = syn.make_uniforms(n, Nc, Nf, seed=31)
    target_s, target_h = syn.make_train_targets(n, K=20, seed=32)
    embed_fn, _ = H.get_embedder(9, 0)
    embeddirs_fn, _ = H.get_embedder(0, 0)
    qf64 = lambda inputs, viewdirs, embedded_cam, network_fn: R.run_network(
        inputs, viewdirs, embedded_cam, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn,
        bb_center=t(bb_center).double(), bb_scale=float(bb_scale), netchunk=1024 * 64)
    with torch.enable_grad():
        scale = torch.tensor([1.1], requires_grad=True, dtype=torch.float64)
        shift = torch.tensor([-0.05], requires_grad=True, dtype=torch.float64)
        th = t(target_h).double() * scale + shift
        kwargs = dict(network_fn=netc, network_query_fn=qf64, N_samples=Nc, embedded_cam=torch.tensor(()),
                      retraw=True, perturb=1.0, N_importance=Nf, network_fine=netf, raw_noise_std=0.0)
        with InjectRand(t_rand.astype(np.float64), u_c.astype(np.float64)):
            ret = R.render_rays(t(rb).double(), True, cached_u=t(u_f).double(), **kwargs)
        img_loss = H.img2mse(ret["rgb_map"], t(target_s).double())
        scl = H.compute_space_carving_loss(ret["pred_hyp"], th, is_joint=False, norm_p=2, threshold=0.0)
        img_loss0 = H.img2mse(ret["rgb0"], t(target_s).double())
        loss = img_loss + 0.007 * scl + img_loss0
        loss.backward()
    d = dict(loss=loss.item(), img_loss=img_loss.item(), img_loss0=img_loss0.item(), space_carving=scl.item(),
             d_scale=scale.grad.numpy(), d_shift=shift.grad.numpy(), rgb_map=ret["rgb_map"].detach().numpy(),
             pred_hyp=ret["pred_hyp"].detach().numpy(), weights=ret["weights"].detach().numpy())
    for pref, net in (("c.", netc), ("f.", netf)):
        for k, p in net.named_parameters():
            d[pref + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "train_small_net_f64.npz"), **d)
    print("train f64 loss", loss.item())

    sizes = {f: os.path.getsize(os.path.join(HERE, f)) for f in sorted(os.listdir(HERE)) if f.endswith(".npz")}
    print(sizes, sum(sizes.values()))


if __name__ == "__main__":
    main()
