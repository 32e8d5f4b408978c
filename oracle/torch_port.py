"""CPU BASELINE PORT (torch ops) of the SCADE render path -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Only bench.py's ``cpu_baseline`` / ``--impl reference`` legs and tests/ may import this.  The reference
is a PyTorch program, so its CPU cost is that of ATen's multi-threaded kernels (MKL SGEMM, vectorised
elementwise ops, 1.7k op calls per render_rays).  The numpy oracle is single-threaded outside
BLAS and would understate the reference; this file restates the same op sequence with torch CPU ops so that
the timed baseline does the work the reference does, the way it does it (``kind: "port"``; the unmodified
reference cannot travel to the GPU box: /root/reference does not exist there).
tests/test_oracle_golden.py::test_torch_port_matches_oracle pins it to the numpy oracle.

Reference lines followed: RS = run_scade_scannet.py, H = model/run_nerf_helpers.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def embed(x, multires):                                       # H:142-172
    outs = [x]
    for k in range(multires):
        arg = x * math.pi * (2.0 ** k)
        outs += [torch.sin(arg), torch.cos(arg)]
    return torch.cat(outs, -1)


def nerf_forward(p, x, input_ch=57, skips=(4,)):               # H:223-247
    D = len([k for k in p if k.startswith("pts_linears.") and k.endswith(".weight")])
    input_pts, input_views = x[:, :input_ch], x[:, input_ch:]
    h = input_pts
    for i in range(D):
        h = F.relu(F.linear(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
        if i in skips:
            h = torch.cat([input_pts, h], -1)
    alpha = F.linear(h, p["alpha_linear.weight"], p["alpha_linear.bias"])
    feature = F.linear(h, p["feature_linear.weight"], p["feature_linear.bias"])
    h = torch.cat([feature, input_views], -1)
    h = F.relu(F.linear(h, p["views_linears.0.weight"], p["views_linears.0.bias"]))
    rgb = F.linear(h, p["rgb_linear.weight"], p["rgb_linear.bias"])
    return torch.cat([rgb, F.softplus(alpha, beta=10)], -1)


def run_network(pts, viewdirs, p, bb_center, bb_scale, multires=9, netchunk=1024 * 64):    # RS:48-63
    flat = (pts.reshape(-1, 3) - bb_center) * bb_scale
    emb = embed(flat, multires)
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
    emb = torch.cat([emb, dirs], -1)
    out = torch.cat([nerf_forward(p, emb[i:i + netchunk], 3 + 6 * multires) for i in range(0, emb.shape[0], netchunk)], 0)
    return out.reshape(*pts.shape[:-1], 4)


def raw2outputs(raw, z, rays_d):                               # RS:511-562
    dists = z[..., 1:] - z[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1) * torch.norm(rays_d[..., None, :], dim=-1)
    alpha = 1. - torch.exp(-F.relu(raw[..., 3]) * dists)
    w = alpha * torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1)), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    rgb = torch.sigmoid(raw[..., :3])
    rgb_map = torch.sum(w[..., None] * rgb, -2)
    depth = torch.sum(w * z, -1)
    acc = torch.sum(w, -1)
    disp = 1. / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    return rgb_map, disp, acc, w, depth


def sample_pdf(bins, weights, n, u=None):                      # H:337-436
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        u = torch.linspace(0., 1., steps=n).expand(list(cdf.shape[:-1]) + [n])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b_lo, b_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return b_lo + (u - cdf_lo) / denom * (b_hi - b_lo), u


def render_rays(ray_batch, pc, pf, bb_center, bb_scale, n_samples, n_importance, t_rand=None, u_coarse=None,
                u_fine=None, multires=9):                      # RS:581-751
    rays_o, rays_d, viewdirs = ray_batch[:, 0:3], ray_batch[:, 3:6], ray_batch[:, 8:11]
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    t = torch.linspace(0., 1., steps=n_samples)
    z = near * (1. - t) + far * t
    if t_rand is not None:
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper, lower = torch.cat([mids, z[..., -1:]], -1), torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    raw = run_network(pts, viewdirs, pc, bb_center, bb_scale, multires)
    rgb0, disp0, acc0, w0, depth0 = raw2outputs(raw, z, rays_d)
    z0 = z
    mid = .5 * (z[..., 1:] + z[..., :-1])
    z_samples, _ = sample_pdf(mid, w0[..., 1:-1], n_importance, u_coarse)
    z, _ = torch.sort(torch.cat([z, z_samples.detach()], -1), -1)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    raw = run_network(pts, viewdirs, pf, bb_center, bb_scale, multires)
    rgb, disp, acc, w, depth = raw2outputs(raw, z, rays_d)
    mid = .5 * (z[..., 1:] + z[..., :-1])
    hyp, u = sample_pdf(mid, w[..., 1:-1], n_importance, u_fine)
    return {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth, "z_vals": z, "weights": w,
            "pred_hyp": hyp, "u": u, "rgb0": rgb0, "disp0": disp0, "acc0": acc0, "depth0": depth0, "z_vals0": z0,
            "weights0": w0, "z_std": torch.std(hyp, dim=-1, unbiased=False)}
