"""CPU ORACLE for the SCADE per-ray hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module; nothing under ``scade_b200/`` does, and the product
path raises when its CUDA library is missing instead of falling back here.

This is a plain numpy restatement of the reference algorithm (mikacuy/scade @ 23139b1), each
function citing the reference lines it follows (RS = run_scade_scannet.py,
H = model/run_nerf_helpers.py).  The reference ships no tests or golden vectors
("parity unpinned" by the reference itself); the oracle is therefore pinned
against outputs of the reference executed in the build container:
``tests/golden/generate_goldens.py`` imports the unmodified reference, runs it on inputs
from ``scade_b200.synthetic`` and commits the results under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them, together with the
known-answer vectors of SURVEY Appendix B.

All arithmetic runs in ``dtype`` (float32 by default, like the reference; float64 is used by
tests to decide which of two fp32 results is nearer the truth).  Backward functions implement
the analytic gradients of SURVEY Appendix A and are pinned against the reference's autograd.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def linspace(start, end, steps, dtype=F32):
    """torch.linspace's scalar formula (ATen RangeFactories: start + step*i below the midpoint,
    end - step*(steps-1-i) from it on), each product/sum rounded to ``dtype``.  The reference calls
    torch.linspace at RS:640 and H:347; ATen's AVX path and nvcc's FMA contraction move single values
    by one ulp between platforms, so parity on u / t_vals is stated to 1 ulp, not bit-exact."""
    if steps == 1:
        return np.array([start], dtype)
    step = dtype((dtype(end) - dtype(start)) / dtype(steps - 1))
    i = np.arange(steps)
    lo = (dtype(start) + (step * i.astype(dtype)).astype(dtype)).astype(dtype)
    hi = (dtype(end) - (step * (steps - 1 - i).astype(dtype)).astype(dtype)).astype(dtype)
    return np.where(i < steps // 2, lo, hi).astype(dtype)


# --------------------------------------------------------------------------------------
# a5  Embedder  (H:142-189)
# --------------------------------------------------------------------------------------
def embed(x, multires, dtype=F32):
    """get_embedder(multires, 0): [x, sin(x*pi*2^k), cos(x*pi*2^k)]_{k<multires}  (H:151-166).

    Product order is (x * pi) * freq with pi rounded to the tensor dtype (H:165);
    multires == 0 -> identity on 3 channels (num_freqs = 0).
    """
    x = np.asarray(x, dtype)
    outs = [x]
    xp = x * dtype(np.pi)
    for k in range(multires):
        arg = xp * dtype(2.0 ** k)
        outs.append(np.sin(arg).astype(dtype))
        outs.append(np.cos(arg).astype(dtype))
    return np.concatenate(outs, -1)


# --------------------------------------------------------------------------------------
# a6  NeRF MLP  (H:193-247)
# --------------------------------------------------------------------------------------
def softplus_beta10(x):
    """F.softplus(x, beta=10) with torch's default threshold=20 (H:242)."""
    x = np.asarray(x)
    bx = x * x.dtype.type(10.0)
    safe = np.minimum(bx, x.dtype.type(20.0))
    return np.where(bx > 20.0, x, np.log1p(np.exp(safe)) / x.dtype.type(10.0)).astype(x.dtype)


def sigmoid(x):
    x = np.asarray(x)
    one = x.dtype.type(1.0)
    return (one / (one + np.exp(-x))).astype(x.dtype)


def _num_pts_layers(params):
    return len([k for k in params if k.startswith("pts_linears.") and k.endswith(".weight")])


def nerf_forward(params, x, input_ch=57, skips=(4,), dtype=F32, return_acts=False):
    """NeRF(use_viewdirs=True).forward (H:223-247).  x: [P, input_ch + input_ch_views]."""
    x = np.asarray(x, dtype)
    p = {k: np.asarray(v, dtype) for k, v in params.items()}
    D = _num_pts_layers(p)
    input_pts, input_views = x[:, :input_ch], x[:, input_ch:]
    h = input_pts
    acts = {"input_pts": input_pts, "input_views": input_views, "lin_in": [], "pre": []}
    for i in range(D):
        acts["lin_in"].append(h)
        z = h @ p[f"pts_linears.{i}.weight"].T + p[f"pts_linears.{i}.bias"]       # H:227
        acts["pre"].append(z)
        h = np.maximum(z, 0)                                                        # H:228
        if i in skips:
            h = np.concatenate([input_pts, h], -1)                                  # H:230
    acts["h_last"] = h
    alpha = h @ p["alpha_linear.weight"].T + p["alpha_linear.bias"]                 # H:233
    feature = h @ p["feature_linear.weight"].T + p["feature_linear.bias"]           # H:234
    hv_in = np.concatenate([feature, input_views], -1)                              # H:235
    zv = hv_in @ p["views_linears.0.weight"].T + p["views_linears.0.bias"]          # H:238
    hv = np.maximum(zv, 0)                                                          # H:239
    rgb = hv @ p["rgb_linear.weight"].T + p["rgb_linear.bias"]                      # H:241
    out = np.concatenate([rgb, softplus_beta10(alpha)], -1).astype(dtype)           # H:242
    if return_acts:
        acts.update(alpha=alpha, hv_in=hv_in, zv=zv, hv=hv)
        return out, acts
    return out


def nerf_backward(params, x, d_out, input_ch=57, skips=(4,), dtype=F32):
    """Gradient of nerf_forward w.r.t. every parameter given d_out [P,4] (autograd of H:223-247).
    Returns dict name -> grad (same keys as params).  No input gradient (RS:711 detaches z)."""
    p = {k: np.asarray(v, dtype) for k, v in params.items()}
    _, a = nerf_forward(params, x, input_ch, skips, dtype, return_acts=True)
    d_out = np.asarray(d_out, dtype)
    D = _num_pts_layers(p)
    g = {}
    d_rgb, d_sigma = d_out[:, :3], d_out[:, 3:4]
    # softplus'(x) = sigmoid(10 x) below the threshold, 1 above
    d_alpha = d_sigma * np.where(a["alpha"] * 10.0 > 20.0, 1.0, sigmoid(a["alpha"] * dtype(10.0))).astype(dtype)
    g["rgb_linear.weight"] = d_rgb.T @ a["hv"]
    g["rgb_linear.bias"] = d_rgb.sum(0)
    d_hv = d_rgb @ p["rgb_linear.weight"]
    d_zv = d_hv * (a["zv"] > 0)
    g["views_linears.0.weight"] = d_zv.T @ a["hv_in"]
    g["views_linears.0.bias"] = d_zv.sum(0)
    W = p["feature_linear.weight"].shape[0]
    d_feature = (d_zv @ p["views_linears.0.weight"])[:, :W]
    g["feature_linear.weight"] = d_feature.T @ a["h_last"]
    g["feature_linear.bias"] = d_feature.sum(0)
    g["alpha_linear.weight"] = d_alpha.T @ a["h_last"]
    g["alpha_linear.bias"] = d_alpha.sum(0)
    d_h = d_feature @ p["feature_linear.weight"] + d_alpha @ p["alpha_linear.weight"]
    for i in reversed(range(D)):
        if i in skips:
            d_h = d_h[:, input_ch:]            # concat put input_pts first (H:230); it has no grad
        d_z = d_h * (a["pre"][i] > 0)
        g[f"pts_linears.{i}.weight"] = d_z.T @ a["lin_in"][i]
        g[f"pts_linears.{i}.bias"] = d_z.sum(0)
        if i > 0:
            d_h = d_z @ p[f"pts_linears.{i}.weight"]
    return {k: v.astype(dtype) for k, v in g.items()}


# --------------------------------------------------------------------------------------
# a4  run_network  (RS:48-63)
# --------------------------------------------------------------------------------------
def network_inputs(pts, viewdirs, bb_center, bb_scale, multires=9, multires_views=0, dtype=F32):
    """The [P, 57+3] matrix run_network feeds the net: (pts - bb_center) * bb_scale -> embed,
    concat embedded dirs broadcast over samples (RS:51-59)."""
    pts = np.asarray(pts, dtype)
    flat = (pts.reshape(-1, 3) - np.asarray(bb_center, dtype)) * dtype(bb_scale)    # RS:52
    emb = embed(flat, multires, dtype)                                               # RS:53
    dirs = np.broadcast_to(np.asarray(viewdirs, dtype)[:, None, :], pts.shape).reshape(-1, 3)  # RS:56-57
    emb_d = embed(dirs, multires_views, dtype)                                       # RS:58
    return np.concatenate([emb, emb_d], -1)                                          # RS:59


def run_network(pts, viewdirs, params, bb_center, bb_scale, multires=9, multires_views=0,
                skips=(4,), dtype=F32):
    x = network_inputs(pts, viewdirs, bb_center, bb_scale, multires, multires_views, dtype)
    input_ch = 3 + 6 * multires
    out = nerf_forward(params, x, input_ch, skips, dtype)
    return out.reshape(pts.shape[:-1] + (4,))                                        # RS:62


# --------------------------------------------------------------------------------------
# a7/a8  compositing  (RS:511-562)
# --------------------------------------------------------------------------------------
def compute_weights(raw, z_vals, rays_d, noise=0.0, dtype=F32):
    """RS:511-522: dists (last = 1e10) * |rays_d|, alpha = 1-exp(-relu(sigma+noise)*dists),
    w = alpha * exclusive_cumprod(1 - alpha + 1e-10)."""
    raw = np.asarray(raw, dtype)
    z = np.asarray(z_vals, dtype)
    dists = z[..., 1:] - z[..., :-1]                                                 # RS:514
    dists = np.concatenate([dists, np.full_like(dists[..., :1], 1e10)], -1)          # RS:515
    dists = dists * np.linalg.norm(np.asarray(rays_d, dtype)[..., None, :], axis=-1).astype(dtype)  # RS:516
    sig = np.maximum(raw[..., 3] + np.asarray(noise, dtype), 0)
    with np.errstate(over="ignore"):
        alpha = (dtype(1.0) - np.exp(-sig * dists)).astype(dtype)                    # RS:512,518
    t = np.concatenate([np.ones((alpha.shape[0], 1), dtype), dtype(1.0) - alpha + dtype(1e-10)], -1)
    trans = np.cumprod(t, -1, dtype=dtype)[:, :-1]                                   # RS:520
    return (alpha * trans).astype(dtype)


def raw2outputs(raw, z_vals, rays_d, noise=0.0, dtype=F32):
    """RS:530-562 -> (rgb_map, disp_map, acc_map, weights, depth_map).  ``noise`` is the already
    drawn sigma noise tensor (RS:545-552 draws it with randn * raw_noise_std)."""
    raw = np.asarray(raw, dtype)
    z = np.asarray(z_vals, dtype)
    rgb = sigmoid(raw[..., :3])                                                      # RS:543
    w = compute_weights(raw, z, rays_d, noise, dtype)                                # RS:554
    rgb_map = np.sum(w[..., None] * rgb, -2, dtype=dtype)                            # RS:556
    depth_map = np.sum(w * z, -1, dtype=dtype)                                       # RS:558
    acc_map = np.sum(w, -1, dtype=dtype)                                             # RS:560
    with np.errstate(divide="ignore", invalid="ignore"):
        q = depth_map / acc_map
        # torch.max propagates nan (acc == 0); np.maximum does as well
        disp_map = (dtype(1.0) / np.maximum(dtype(1e-10), q)).astype(dtype)          # RS:559
    return rgb_map, disp_map, acc_map, w, depth_map


def raw2outputs_bwd(raw, z_vals, rays_d, d_rgb_map, d_disp, d_acc, d_weights, d_depth,
                    noise=0.0, dtype=F32):
    """d raw [N,S,4] from the gradients of raw2outputs' five outputs (autograd of RS:511-562;
    formulas of SURVEY Appendix A).  z_vals and rays_d carry no gradient on the reference path."""
    raw = np.asarray(raw, dtype)
    z = np.asarray(z_vals, dtype)
    N, S = z.shape
    zero = np.zeros((N,), dtype)
    d_rgb_map = np.zeros((N, 3), dtype) if d_rgb_map is None else np.asarray(d_rgb_map, dtype)
    d_disp = zero if d_disp is None else np.asarray(d_disp, dtype)
    d_acc = zero if d_acc is None else np.asarray(d_acc, dtype)
    d_depth = zero if d_depth is None else np.asarray(d_depth, dtype)
    d_w_in = np.zeros((N, S), dtype) if d_weights is None else np.asarray(d_weights, dtype)

    rgb = sigmoid(raw[..., :3])
    dists = np.concatenate([z[..., 1:] - z[..., :-1], np.full((N, 1), 1e10, dtype)], -1)
    dists = dists * np.linalg.norm(np.asarray(rays_d, dtype)[..., None, :], axis=-1).astype(dtype)
    pre = raw[..., 3] + np.asarray(noise, dtype)
    sig = np.maximum(pre, 0)
    with np.errstate(over="ignore"):
        e = np.exp(-sig * dists).astype(dtype)
    alpha = dtype(1.0) - e
    tfac = dtype(1.0) - alpha + dtype(1e-10)
    trans = np.cumprod(np.concatenate([np.ones((N, 1), dtype), tfac], -1), -1, dtype=dtype)[:, :-1]
    w = alpha * trans
    depth = np.sum(w * z, -1)
    acc = np.sum(w, -1)
    # disp = 1/max(1e-10, depth/acc)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = depth / acc
        live = q > 1e-10
        d_q = np.where(live, -d_disp / (q * q), 0).astype(dtype)
        d_depth_t = d_depth + np.where(live, d_q / acc, 0)
        d_acc_t = d_acc + np.where(live, -d_q * depth / (acc * acc), 0)
    g = d_w_in + (d_rgb_map[:, None, :] * rgb).sum(-1) + d_depth_t[:, None] * z + d_acc_t[:, None]
    # dL/dalpha_i = g_i T_i - (sum_{j>i} g_j w_j) / (1 - alpha_i + 1e-10)
    gw = g * w
    suffix = np.cumsum(gw[:, ::-1], -1)[:, ::-1] - gw
    d_alpha = g * trans - suffix / tfac
    d_sig = d_alpha * dists * e * (pre > 0)
    d_raw = np.zeros_like(raw)
    d_raw[..., 3] = d_sig
    d_raw[..., :3] = w[..., None] * rgb * (dtype(1.0) - rgb) * d_rgb_map[:, None, :]
    return d_raw.astype(dtype)


# --------------------------------------------------------------------------------------
# a9  perturb_z_vals  (RS:564-579)  and the coarse z schedule  (RS:640-655)
# --------------------------------------------------------------------------------------
def coarse_z_vals(near, far, n_samples, lindisp=False, dtype=F32):
    t32 = linspace(0.0, 1.0, n_samples, F32)            # RS:640 (torch.linspace is fp32 whatever the rays are)
    t, omt = t32.astype(dtype), (F32(1.0) - t32).astype(dtype)   # (1.-t_vals) is formed in fp32 too
    near = np.asarray(near, dtype).reshape(-1, 1)
    far = np.asarray(far, dtype).reshape(-1, 1)
    if not lindisp:
        return (near * omt + far * t).astype(dtype)                                  # RS:648
    return (dtype(1.0) / (dtype(1.0) / near * omt + dtype(1.0) / far * t)).astype(dtype)   # RS:651


def perturb_z_vals(z_vals, t_rand, dtype=F32):
    z = np.asarray(z_vals, dtype)
    mids = dtype(0.5) * (z[..., 1:] + z[..., :-1])                                   # RS:566
    upper = np.concatenate([mids, z[..., -1:]], -1)                                  # RS:567
    lower = np.concatenate([z[..., :1], mids], -1)                                   # RS:568
    return (lower + (upper - lower) * np.asarray(t_rand, dtype)).astype(dtype)       # RS:578


# --------------------------------------------------------------------------------------
# a10  sample_pdf family  (H:337-538)
# --------------------------------------------------------------------------------------
def _cdf(weights, dtype):
    w = np.asarray(weights, dtype) + dtype(1e-5)                                     # H:339
    pdf = w / np.sum(w, -1, keepdims=True, dtype=dtype)                              # H:340
    cdf = np.cumsum(pdf, -1, dtype=dtype)                                            # H:342
    return w, pdf, np.concatenate([np.zeros_like(cdf[..., :1]), cdf], -1)            # H:343


def sample_pdf(bins, weights, n_samples, det=False, u=None, dtype=F32, return_aux=False):
    """sample_pdf / sample_pdf_return_u (H:337-436).  ``u`` plays load_u (H:393,412-414);
    det -> linspace(0,1,n) including both endpoints (H:347).  Returns (samples, u)."""
    bins = np.asarray(bins, dtype)
    _, _, cdf = _cdf(weights, dtype)
    N = cdf.shape[0]
    if u is None:
        if not det:
            raise ValueError("oracle needs explicit uniforms when det=False")
        u = np.broadcast_to(linspace(0.0, 1.0, n_samples, dtype), (N, n_samples))       # H:347
    u = np.ascontiguousarray(np.asarray(u, dtype))
    if u.ndim == 1:                       # sample_pdf_joint: one row shared by all rays (H:452-453)
        u = np.ascontiguousarray(np.broadcast_to(u, (N, u.shape[0])))
    inds = np.stack([np.searchsorted(cdf[r], u[r], side="right") for r in range(N)], 0)  # H:366
    below = np.maximum(0, inds - 1)                                                  # H:368
    above = np.minimum(cdf.shape[-1] - 1, inds)                                      # H:369
    rows = np.arange(N)[:, None]
    cdf_lo, cdf_hi = cdf[rows, below], cdf[rows, above]                              # H:375
    b_lo, b_hi = bins[rows, below], bins[rows, above]                                # H:376
    denom = cdf_hi - cdf_lo                                                          # H:378
    denom = np.where(denom < 1e-5, np.ones_like(denom), denom)                       # H:379
    t = (u - cdf_lo) / denom                                                         # H:380
    samples = (b_lo + t * (b_hi - b_lo)).astype(dtype)                               # H:381
    if return_aux:
        return samples, u, dict(below=below, above=above, cdf=cdf)
    return samples, u


def sample_pdf_bwd(bins, weights, u, d_samples, dtype=F32):
    """d weights from d samples (autograd of H:339-381 w.r.t. weights only: bins are grad-free on
    the reference path since z_samples is detached, RS:711).  SURVEY Appendix A."""
    bins = np.asarray(bins, dtype)
    w, pdf, cdf = _cdf(weights, dtype)
    u = np.asarray(u, dtype)
    if u.ndim == 1:
        u = np.broadcast_to(u, (cdf.shape[0], u.shape[0]))
    d_s = np.asarray(d_samples, dtype)
    N, B = cdf.shape
    inds = np.stack([np.searchsorted(cdf[r], u[r], side="right") for r in range(N)], 0)
    below = np.maximum(0, inds - 1)
    above = np.minimum(B - 1, inds)
    rows = np.arange(N)[:, None]
    c_lo, c_hi = cdf[rows, below], cdf[rows, above]
    db = bins[rows, above] - bins[rows, below]
    den = c_hi - c_lo
    clamped = den < 1e-5
    den_s = np.where(clamped, np.ones_like(den), den)
    # s = b_lo + (u - c_lo)/den * db
    g_lo = np.where(clamped, -db, db * (u - c_hi) / (den_s * den_s)) * d_s
    g_hi = np.where(clamped, 0.0, -db * (u - c_lo) / (den_s * den_s)) * d_s
    d_cdf = np.zeros((N, B), dtype)
    np.add.at(d_cdf, (np.broadcast_to(rows, below.shape), below), g_lo.astype(dtype))
    np.add.at(d_cdf, (np.broadcast_to(rows, above.shape), above), g_hi.astype(dtype))
    # cdf[j] = sum_{m<j} pdf[m]  ->  d pdf[m] = sum_{j>m} d cdf[j]
    d_pdf = np.cumsum(d_cdf[:, ::-1], -1)[:, ::-1][:, 1:]
    tot = np.sum(w, -1, keepdims=True)
    d_w = (d_pdf - np.sum(d_pdf * pdf, -1, keepdims=True)) / tot
    return d_w.astype(dtype)


# --------------------------------------------------------------------------------------
# a12  space-carving loss  (H:93-128)
# --------------------------------------------------------------------------------------
def _sc_distances(pred, hyp, mask, threshold, dtype):
    pred = np.asarray(pred, dtype)
    hyp = np.asarray(hyp, dtype)
    if hyp.shape[-1] == 1:
        hyp = np.broadcast_to(hyp, hyp.shape[:2] + (pred.shape[1],))                 # H:97-99
    diff = pred[None] - hyp
    dist = np.abs(diff)                                                              # H:106 (norm over a singleton dim)
    if mask is not None:
        dist = dist * np.asarray(mask, dtype)[None, :, None]                         # H:108-110
    if threshold > 0:
        dist = np.where(dist < threshold, dtype(0.0), dist)                          # H:112-113
    return diff, dist.astype(dtype)


def space_carving_loss(pred, hyp, is_joint=False, mask=None, threshold=0.0, dtype=F32):
    _, dist = _sc_distances(pred, hyp, mask, threshold, dtype)
    if is_joint:
        qm = np.mean(dist, 1, dtype=dtype)                                           # H:117
        return dtype(np.mean(np.min(qm, 0), dtype=dtype))                            # H:118-119
    best = np.min(dist, 0)                                                           # H:124
    return dtype(np.mean(np.mean(best, -1, dtype=dtype), dtype=dtype))               # H:125-126


def space_carving_loss_bwd(pred, hyp, is_joint=False, mask=None, threshold=0.0, dtype=F32):
    """(d_pred [N,P], d_hyp same shape as hyp) for an upstream gradient of 1 (autograd of H:93-128).
    Ties in min pick the FIRST k (torch.min(dim) semantics); abs has zero gradient at 0."""
    pred = np.asarray(pred, dtype)
    hyp_in = np.asarray(hyp, dtype)
    diff, dist = _sc_distances(pred, hyp_in, mask, threshold, dtype)
    K, N, P = dist.shape
    m = np.ones((N,), dtype) if mask is None else np.asarray(mask, dtype)
    sgn = np.sign(diff) * m[None, :, None]
    if threshold > 0:
        sgn = np.where(np.abs(diff) * m[None, :, None] < threshold, 0.0, sgn)
    g = np.zeros((K, N, P), dtype)
    if is_joint:
        qm = np.mean(dist, 1)
        kstar = np.argmin(qm, 0)                      # [P]
        for p in range(P):
            g[kstar[p], :, p] = sgn[kstar[p], :, p] / (N * P)
    else:
        kstar = np.argmin(dist, 0)                    # [N,P]
        nn, pp = np.meshgrid(np.arange(N), np.arange(P), indexing="ij")
        g[kstar, nn, pp] = sgn[kstar, nn, pp] / (N * P)
    d_pred = g.sum(0)
    d_hyp = -g
    if hyp_in.shape[-1] == 1:
        d_hyp = d_hyp.sum(-1, keepdims=True)
    return d_pred.astype(dtype), d_hyp.astype(dtype)


def img2mse(x, y, dtype=F32):
    """H:11"""
    d = np.asarray(x, dtype) - np.asarray(y, dtype)
    return dtype(np.mean(d * d, dtype=dtype))


def mse2psnr(x):
    """H:12"""
    return F32(-10.0 * np.log(x) / np.log(10.0))


# --------------------------------------------------------------------------------------
# a3  get_rays  (H:285-305)
# --------------------------------------------------------------------------------------
def get_rays(H, W, intrinsic, c2w, dtype=F32):
    fx, fy, cx, cy = [dtype(v) for v in intrinsic]
    c2w = np.asarray(c2w, dtype)
    j, i = np.meshgrid(np.arange(H, dtype=dtype), np.arange(W, dtype=dtype), indexing="ij")   # H:289-291
    dirs = np.stack([((i + dtype(0.5)) - cx) / fx, (dtype(H) - (j + dtype(0.5)) - cy) / fy,
                     -np.ones_like(i)], -1)                                          # H:295
    rays_d = np.sum(dirs[..., None, :] * c2w[:3, :3], -1).astype(dtype)              # H:297
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape).astype(dtype)                # H:304
    return rays_o, rays_d


def make_ray_batch(rays_o, rays_d, near, far, dtype=F32):
    """render()'s assembly of the [N,11] batch (RS:123-141)."""
    rays_o = np.asarray(rays_o, dtype).reshape(-1, 3)
    rays_d = np.asarray(rays_d, dtype).reshape(-1, 3)
    viewdirs = rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)               # RS:129
    nearv = dtype(near) * np.ones_like(rays_d[:, :1])
    farv = dtype(far) * np.ones_like(rays_d[:, :1])
    return np.concatenate([rays_o, rays_d, nearv, farv, viewdirs], -1).astype(dtype)


# --------------------------------------------------------------------------------------
# a1  render_rays  (RS:581-751), N_importance > 0 branch (the only live one)
# --------------------------------------------------------------------------------------
def render_rays(ray_batch, params_coarse, params_fine, bb_center, bb_scale, N_samples, N_importance,
                perturb=0.0, t_rand=None, u_coarse=None, u_fine=None, lindisp=False,
                multires=9, multires_views=0, skips=(4,), is_joint=False, retraw=False,
                dtype=F32, mlp=None, z_fine=None):
    """Returns the same dict as the reference (RS:733-744).  Random draws are explicit:
    t_rand [N,Nc] (RS:570), u_coarse [N,Nimp] (H:350 in the RS:705 call), u_fine = cached_u
    (RS:726).  ``mlp`` optionally replaces run_network (used to emulate operand rounding).
    ``z_fine`` [N,Nc+Nimp] (test infrastructure) teacher-forces the merged sample positions of RS:713:
    the fine pass is then evaluated at exactly the positions another implementation chose, which
    removes the (discontinuous, noise-amplifying) inverse-CDF resampling from a comparison of the
    fine network, its compositing and its gradients."""
    rb = np.asarray(ray_batch, dtype)
    rays_o, rays_d = rb[:, 0:3], rb[:, 3:6]                                          # RS:628
    viewdirs = rb[:, 8:11]                                                           # RS:632
    near, far = rb[:, 6:7], rb[:, 7:8]                                               # RS:638-639
    if params_fine is None:
        params_fine = params_coarse                                                  # RS:716
    net = mlp or (lambda pts, p: run_network(pts, viewdirs, p, bb_center, bb_scale, multires,
                                             multires_views, skips, dtype))
    z_vals = coarse_z_vals(near, far, N_samples, lindisp, dtype)                     # RS:640-651
    det = not (perturb > 0.0)
    if not det:
        z_vals = perturb_z_vals(z_vals, t_rand, dtype)                               # RS:653-655
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]               # RS:657
    raw = net(pts, params_coarse)                                                    # RS:659
    rgb0, disp0, acc0, w0, depth0 = raw2outputs(raw, z_vals, rays_d, 0.0, dtype)     # RS:660
    z0 = z_vals
    mid = dtype(0.5) * (z_vals[:, 1:] + z_vals[:, :-1])                              # RS:702
    z_samples, _ = sample_pdf(mid, w0[:, 1:-1], N_importance, det=det,
                              u=None if det else u_coarse, dtype=dtype)              # RS:705
    z_vals = np.sort(np.concatenate([z_vals, z_samples], -1), -1)                    # RS:713
    if z_fine is not None:
        z_vals = np.asarray(z_fine, dtype)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]               # RS:714
    raw = net(pts, params_fine)                                                      # RS:718
    rgb, disp, acc, w, depth = raw2outputs(raw, z_vals, rays_d, 0.0, dtype)          # RS:720
    mid = dtype(0.5) * (z_vals[:, 1:] + z_vals[:, :-1])                              # RS:723
    pred_hyp, u = sample_pdf(mid, w[:, 1:-1], N_importance, det=det,
                             u=None if det else u_fine, dtype=dtype)                 # RS:726
    ret = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth, "z_vals": z_vals,
           "weights": w, "pred_hyp": pred_hyp, "u": np.asarray(u, dtype),            # RS:733-734
           "rgb0": rgb0, "disp0": disp0, "acc0": acc0, "depth0": depth0, "z_vals0": z0,
           "weights0": w0,                                                           # RS:738-743
           "z_std": np.std(pred_hyp, -1).astype(dtype)}                              # RS:744 (unbiased=False)
    if retraw:
        ret["raw"] = raw                                                             # RS:736
    return ret


def train_loss_and_grads(ray_batch, params_coarse, params_fine, bb_center, bb_scale, N_samples,
                         N_importance, target_s, target_h, t_rand, u_coarse, u_fine,
                         space_carving_weight=0.007, scale=1.0, shift=0.0, mask=None,
                         threshold=0.0, multires=9, multires_views=0, skips=(4,), dtype=F32, z_fine=None):
    """One training step's loss and gradients (RS:954-985): loss = mse(rgb) + w_sc * space_carving
    + mse(rgb0).  Gradient reach follows the reference's autograd graph: the fine net gets grads from
    the fine MSE and from space carving through sample_pdf_return_u; the coarse net only from the
    coarse MSE (z_samples is detached, RS:711).  Returns (losses dict, grads_coarse, grads_fine,
    d_scale, d_shift)."""
    rb = np.asarray(ray_batch, dtype)
    rays_d, viewdirs = rb[:, 3:6], rb[:, 8:11]
    input_ch = 3 + 6 * multires
    out = render_rays(rb, params_coarse, params_fine, bb_center, bb_scale, N_samples, N_importance,
                      perturb=1.0, t_rand=t_rand, u_coarse=u_coarse, u_fine=u_fine,
                      multires=multires, multires_views=multires_views, skips=skips, retraw=True,
                      dtype=dtype, z_fine=z_fine)
    target_s = np.asarray(target_s, dtype)
    h_raw = np.asarray(target_h, dtype)
    h = h_raw * dtype(scale) + dtype(shift)                                          # RS:954
    N = rb.shape[0]
    img_loss = img2mse(out["rgb_map"], target_s, dtype)                              # RS:968
    img_loss0 = img2mse(out["rgb0"], target_s, dtype)                                # RS:981
    sc = space_carving_loss(out["pred_hyp"], h, False, mask, threshold, dtype)       # RS:974
    loss = img_loss + dtype(space_carving_weight) * sc + img_loss0                   # RS:976,983
    # ---- backward ----
    d_rgb = (dtype(2.0) / dtype(N * 3)) * (out["rgb_map"] - target_s)
    d_rgb0 = (dtype(2.0) / dtype(N * 3)) * (out["rgb0"] - target_s)
    d_pred, d_h = space_carving_loss_bwd(out["pred_hyp"], h, False, mask, threshold, dtype)
    d_pred = d_pred * dtype(space_carving_weight)
    d_h = d_h * dtype(space_carving_weight)
    d_scale = dtype(np.sum(d_h * h_raw))
    d_shift = dtype(np.sum(d_h))
    zf = out["z_vals"]
    mid = dtype(0.5) * (zf[:, 1:] + zf[:, :-1])
    d_wmid = sample_pdf_bwd(mid, out["weights"][:, 1:-1], out["u"], d_pred, dtype)
    d_w = np.zeros_like(out["weights"])
    d_w[:, 1:-1] = d_wmid
    d_raw_f = raw2outputs_bwd(out["raw"], zf, rays_d, d_rgb, None, None, d_w, None, 0.0, dtype)
    pts_f = rb[:, None, 0:3] + rays_d[:, None, :] * zf[:, :, None]
    x_f = network_inputs(pts_f, viewdirs, bb_center, bb_scale, multires, multires_views, dtype)
    g_fine = nerf_backward(params_fine, x_f, d_raw_f.reshape(-1, 4), input_ch, skips, dtype)
    z0 = out["z_vals0"]
    pts_c = rb[:, None, 0:3] + rays_d[:, None, :] * z0[:, :, None]
    x_c = network_inputs(pts_c, viewdirs, bb_center, bb_scale, multires, multires_views, dtype)
    raw_c = nerf_forward(params_coarse, x_c, input_ch, skips, dtype).reshape(N, -1, 4)
    d_raw_c = raw2outputs_bwd(raw_c, z0, rays_d, d_rgb0, None, None, None, None, 0.0, dtype)
    g_coarse = nerf_backward(params_coarse, x_c, d_raw_c.reshape(-1, 4), input_ch, skips, dtype)
    losses = {"loss": loss, "img_loss": img_loss, "img_loss0": img_loss0, "space_carving": sc}
    return losses, g_coarse, g_fine, d_scale, d_shift, out


# --------------------------------------------------------------------------------------
# operand-rounding emulation (documents the tensor-core mode's tolerance)
# --------------------------------------------------------------------------------------
def nerf_forward_f16(params, x, input_ch=57, skips=(4,), split=False):
    """The MLP as the tcgen05 kernels compute it: activations and weights of the wide layers rounded
    to fp16 before an fp32-accumulated product; bias/ReLU in fp32; alpha and rgb heads as fp32 dot
    products of the UNROUNDED fp32 activations.  ``split=True`` emulates the tight mode
    (SCADE_PREC_TC_F16X3): every operand is an fp16 (hi, lo) pair with lo = fp16(x - hi), and the product
    is hi*hi + lo*hi + hi*lo (three tensor-core passes; the lo*lo term, ~2^-22 relative, is dropped).
    Used only to set/explain test tolerances."""
    r16 = lambda a: np.asarray(a, np.float32).astype(np.float16).astype(np.float32)

    def mm(a, w):
        a, w = np.asarray(a, np.float32), np.asarray(w, np.float32)
        ah, wh = r16(a).astype(np.float64), r16(w).astype(np.float64)
        if not split:
            return (ah @ wh.T).astype(np.float32)
        al = r16(a - ah.astype(np.float32)).astype(np.float64)
        wl = r16(w - wh.astype(np.float32)).astype(np.float64)
        return (ah @ wh.T + al @ wh.T + ah @ wl.T).astype(np.float32)
    p = {k: np.asarray(v, np.float32) for k, v in params.items()}
    D = _num_pts_layers(p)
    x = np.asarray(x, np.float32)
    input_pts, input_views = x[:, :input_ch], x[:, input_ch:]
    h = input_pts
    for i in range(D):
        z = mm(h, p[f"pts_linears.{i}.weight"])
        h = np.maximum(z + p[f"pts_linears.{i}.bias"], 0)
        if i in skips:
            h = np.concatenate([input_pts, h], -1)
    alpha = h @ p["alpha_linear.weight"].T + p["alpha_linear.bias"]
    feature = mm(h, p["feature_linear.weight"]) + p["feature_linear.bias"]
    hv_in = np.concatenate([feature, input_views], -1)
    zv = mm(hv_in, p["views_linears.0.weight"])
    hv = np.maximum(zv + p["views_linears.0.bias"], 0)
    rgb = hv @ p["rgb_linear.weight"].T + p["rgb_linear.bias"]
    return np.concatenate([rgb, softplus_beta10(alpha)], -1).astype(np.float32)


# --------------------------------------------------------------------------------------
# teacher-forced backward for the tensor-core training path (test infrastructure)
# --------------------------------------------------------------------------------------
def nerf_backward_teacher_forced(params, emb, h, feature, hv, alpha_pre, inactive, inactive_v, d_out, input_ch=57,
                                 input_views=3, skips=(4,)):
    """Autograd of NeRF.forward (H:223-247) evaluated in float64 on GIVEN activations and ReLU masks.

    The tensor-core forward rounds operands to fp16, which flips the sign of a few pre-activations that sit within
    ~1e-3 of zero; a gradient check against `nerf_backward` would then measure those flips, not the backward kernels.
    Here the activations (emb [P,>=input_ch+input_views], h[l] [P,W] post-ReLU outputs of pts_linears[l], feature [P,W],
    hv [P,W/2], alpha_pre [P]) and the masks (inactive[l], inactive_v: True where the pre-activation was negative) are
    the ones the kernel stashed, so the result isolates dgrad / wgrad arithmetic.  Returns (grads, dz) with dz the
    per-layer pre-activation gradients {"v", "feature", 0..D-1}."""
    f8 = np.float64
    p = {k: np.asarray(v, f8) for k, v in params.items()}
    D = _num_pts_layers(p)
    emb, feature, hv = np.asarray(emb, f8), np.asarray(feature, f8), np.asarray(hv, f8)
    h = [np.asarray(x, f8) for x in h]
    d_out = np.asarray(d_out, f8)
    d_rgb, d_sigma = d_out[:, :3], d_out[:, 3:4]
    al = np.asarray(alpha_pre, f8).reshape(-1, 1)
    d_alpha = d_sigma * np.where(al * 10.0 > 20.0, 1.0, 1.0 / (1.0 + np.exp(-al * 10.0)))       # softplus'(beta=10), H:242
    g, dz = {}, {}
    W = feature.shape[1]
    input_pts, views = emb[:, :input_ch], emb[:, input_ch:input_ch + input_views]
    g["rgb_linear.weight"] = d_rgb.T @ hv                                                        # H:241
    g["rgb_linear.bias"] = d_rgb.sum(0)
    dz["v"] = (d_rgb @ p["rgb_linear.weight"]) * (~np.asarray(inactive_v, bool))                 # H:239
    g["views_linears.0.weight"] = dz["v"].T @ np.concatenate([feature, views], -1)               # H:235-238
    g["views_linears.0.bias"] = dz["v"].sum(0)
    dz["feature"] = (dz["v"] @ p["views_linears.0.weight"])[:, :W]
    g["feature_linear.weight"] = dz["feature"].T @ h[D - 1]                                      # H:234
    g["feature_linear.bias"] = dz["feature"].sum(0)
    g["alpha_linear.weight"] = d_alpha.T @ h[D - 1]                                              # H:233
    g["alpha_linear.bias"] = d_alpha.sum(0)
    d_h = dz["feature"] @ p["feature_linear.weight"] + d_alpha @ p["alpha_linear.weight"]
    for i in reversed(range(D)):
        dz[i] = d_h * (~np.asarray(inactive[i], bool))                                           # H:228
        x_in = input_pts if i == 0 else (np.concatenate([input_pts, h[i - 1]], -1) if (i - 1) in skips else h[i - 1])
        g[f"pts_linears.{i}.weight"] = dz[i].T @ x_in                                            # H:227, H:230
        g[f"pts_linears.{i}.bias"] = dz[i].sum(0)
        if i > 0:
            d_h = dz[i] @ p[f"pts_linears.{i}.weight"]
            if (i - 1) in skips:
                d_h = d_h[:, input_ch:]
    return g, dz


# --------------------------------------------------------------------------------------
# §8(f) rows: training sampler and video post-processing (restated for the next-row kernels)
# --------------------------------------------------------------------------------------
def train_batch(H, W, img_i, images, depths, valid_depths, poses, intrinsics, all_hypothesis, select_inds, cached_u=None,
                mask_corners=False, near=0.0, far=1.0, dtype=F32):
    """get_ray_batch_from_one_image_hypothesis_idx (RS:772-827) for GIVEN flat pixel indices (the reference draws them with
    np.random.choice, H:281) + the [N,11] ray batch render() assembles from batch_rays (RS:123-141)."""
    rays_o, rays_d = get_rays(H, W, intrinsics[img_i], poses[img_i][:3, :4], dtype)          # RS:784
    rr, cc = np.asarray(select_inds) // W, np.asarray(select_inds) % W                        # coords[select_inds], H:282
    o, d = rays_o[rr, cc], rays_d[rr, cc]
    out = {"batch_rays": np.stack([o, d], 0), "target_s": images[img_i][rr, cc], "target_d": depths[img_i][rr, cc],
           "target_vd": valid_depths[img_i][rr, cc], "target_h": all_hypothesis[img_i][:, rr, cc],       # RS:786-791
           "ray_batch": make_ray_batch(o, d, near, far, dtype)}
    if cached_u is not None:
        out["cached_u"] = cached_u[img_i][rr, cc]                                             # RS:805-806
    if mask_corners:                                                                           # RS:810-821
        m = np.ones((H, W), dtype)
        m[:20, :20] = 0; m[:20, -20:] = 0; m[-20:, :20] = 0; m[-20:, -20:] = 0
        out["mask"] = m[rr, cc]
    return out


def to8b(x):
    """H:13"""
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def to16b(x):
    """H:14"""
    return ((2 ** 16 - 1) * np.clip(x, 0, 1)).astype(np.uint16)


def depth_std(z_vals, weights, depth_map, dtype=F32):
    """RS:257-258"""
    z_vals, weights, depth_map = (np.asarray(a, dtype) for a in (z_vals, weights, depth_map))
    var = (((z_vals - depth_map[..., None]) ** 2) * weights).sum(-1)
    return np.sqrt(np.clip(var, 0.0, 1.0)).astype(dtype)


def video_frame(rgb, depth_map, z_vals, weights, depth_scale, lut_depth=None, lut_std=None):
    """The frame render_video writes (RS:252-259): [to8b(rgb) in BGR | colour-mapped to8b(depth / far) | colour-mapped
    to8b(depth_std)], side by side.  lut_*: [256,3] uint8 BGR tables (cv2.applyColorMap's) or None for grey."""
    rgb8 = to8b(np.asarray(rgb, F32))[..., ::-1]                                              # cv2.COLOR_RGB2BGR
    d8 = to8b((np.asarray(depth_map, F32) / F32(depth_scale)))
    s8 = to8b(depth_std(z_vals, weights, depth_map))
    pan = lambda q, lut: (lut[q] if lut is not None else np.repeat(q[..., None], 3, -1))
    return np.concatenate([rgb8, pan(d8, lut_depth), pan(s8, lut_std)], 1)
