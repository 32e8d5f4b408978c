"""Deterministic synthetic inputs for the SCADE per-ray hot path (numpy only).

There is no dataset and no checkpoint in the build/bench environment, so every
test, golden fixture and benchmark draws its inputs from this factory
(SURVEY.md §8(d)).  Only numpy is used so that the golden generator (which
imports the reference), the oracle, the CUDA tests and bench.py all see
bit-identical inputs regardless of which torch build is around.

Shapes follow the reference:
  * NeRF parameters keep the reference ``state_dict`` names and (out, in) layout
    (model/run_nerf_helpers.py:206-221).
  * rays come from a pinhole camera exactly as ``get_rays`` builds them
    (model/run_nerf_helpers.py:285-305): ``rays_d`` un-normalised, ``viewdirs``
    normalised (run_scade_scannet.py:129-130).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

# Reference defaults (run_scade_scannet.py:1100-1147).
DEFAULT_MULTIRES = 9
DEFAULT_MULTIRES_VIEWS = 0
CAM_H, CAM_W = 480, 640
CAM_INTRINSIC = (585.0, 585.0, 320.0, 240.0)  # fx, fy, cx, cy
NEAR, FAR = 0.1, 5.0


def embed_dim(multires: int) -> int:
    """Output width of get_embedder(multires) (model/run_nerf_helpers.py:174-189)."""
    return 3 + 3 * 2 * multires


def nerf_layer_shapes(D=8, W=256, input_ch=57, input_ch_views=3, skips=(4,)):
    """(name, out, in) for every Linear of NeRF(use_viewdirs=True), in state_dict order.

    Mirrors model/run_nerf_helpers.py:206-219: layer i+1 takes W+input_ch inputs
    when i is in ``skips`` (the concat puts input_pts FIRST, :230).
    """
    shapes = [("pts_linears.0", W, input_ch)]
    for i in range(D - 1):
        fan_in = W + input_ch if i in skips else W
        shapes.append((f"pts_linears.{i + 1}", W, fan_in))
    shapes.append(("views_linears.0", W // 2, input_ch_views + W))
    shapes.append(("feature_linear", W, W))
    shapes.append(("alpha_linear", 1, W))
    shapes.append(("rgb_linear", 3, W // 2))
    return shapes


def make_nerf_params(seed=0, D=8, W=256, input_ch=57, input_ch_views=3, skips=(4,),
                     bias_scale=0.0, alpha_bias=0.0, weight_gain=1.0):
    """Xavier-uniform weights like DenseLayer.reset_parameters (run_nerf_helpers.py:136-139).

    gain = sqrt(2) for relu layers, 1 for linear heads.  The reference zeroes biases;
    ``bias_scale`` > 0 draws U(-s, s) biases instead so tests exercise the bias path,
    and ``alpha_bias`` shifts the density head (SURVEY §8(d) "boost" weights).
    Returns an OrderedDict name -> float32 array with reference state_dict keys.
    """
    rng = np.random.default_rng(seed)
    params = OrderedDict()
    for name, fan_out, fan_in in nerf_layer_shapes(D, W, input_ch, input_ch_views, skips):
        relu = name.startswith("pts_linears") or name.startswith("views_linears")
        gain = math.sqrt(2.0) if relu else 1.0
        bound = weight_gain * gain * math.sqrt(6.0 / (fan_in + fan_out))
        params[name + ".weight"] = rng.uniform(-bound, bound, size=(fan_out, fan_in)).astype(np.float32)
        if bias_scale > 0:
            b = rng.uniform(-bias_scale, bias_scale, size=(fan_out,)).astype(np.float32)
        else:
            b = np.zeros((fan_out,), np.float32)
        if name == "alpha_linear":
            b = b + np.float32(alpha_bias)
        params[name + ".bias"] = b.astype(np.float32)
    return params


def get_rays_np(H, W, intrinsic, c2w, coords=None):
    """numpy twin of get_rays (model/run_nerf_helpers.py:285-305); fp32 arithmetic."""
    fx, fy, cx, cy = [np.float32(v) for v in intrinsic]
    c2w = np.asarray(c2w, np.float32)
    if coords is None:
        j, i = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    else:
        coords = np.asarray(coords)
        i, j = coords[:, 1].astype(np.float32), coords[:, 0].astype(np.float32)
    dirs = np.stack([((i + np.float32(0.5)) - cx) / fx,
                     (np.float32(H) - (j + np.float32(0.5)) - cy) / fy,
                     -np.ones_like(i)], -1).astype(np.float32)
    rays_d = np.sum(dirs[..., None, :] * c2w[:3, :3], -1).astype(np.float32)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape).astype(np.float32)
    return rays_o, rays_d


def make_ray_batch(n_rays, seed=0, near=NEAR, far=FAR, c2w=None, H=CAM_H, W=CAM_W,
                   intrinsic=CAM_INTRINSIC):
    """[N, 11] ray batch = (o3, d3, near, far, viewdir3) as render() assembles it
    (run_scade_scannet.py:129-141).  Pixels are drawn without replacement with a seeded
    generator (the reference uses np.random.choice, run_nerf_helpers.py:279-283)."""
    if c2w is None:
        c2w = np.eye(4, dtype=np.float32)[:3]
    rng = np.random.default_rng(seed)
    flat = rng.choice(H * W, size=n_rays, replace=n_rays > H * W)
    coords = np.stack([flat // W, flat % W], -1)
    rays_o, rays_d = get_rays_np(H, W, intrinsic, c2w, coords)
    viewdirs = rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)
    nearv = np.full((n_rays, 1), near, np.float32)
    farv = np.full((n_rays, 1), far, np.float32)
    return np.concatenate([rays_o, rays_d, nearv, farv, viewdirs.astype(np.float32)], -1).astype(np.float32)


def make_uniforms(n_rays, n_coarse, n_importance, seed=1):
    """Explicit RNG draws injected into both oracle and kernels (SURVEY §7 hard part 6):
    t_rand [N,Nc] for perturb_z_vals, u_coarse / u_fine [N,Nimp] for the two sample_pdf calls."""
    rng = np.random.default_rng(seed)
    t_rand = rng.random((n_rays, n_coarse), dtype=np.float32)
    u_coarse = rng.random((n_rays, n_importance), dtype=np.float32)
    u_fine = rng.random((n_rays, n_importance), dtype=np.float32)
    return t_rand, u_coarse, u_fine


def make_train_targets(n_rays, K=20, seed=2, near=NEAR, far=FAR):
    """target_s [N,3] ~ U[0,1), target_h [K,N,1] ~ U[near,far] (data/load_scene.py:348 clips to that)."""
    rng = np.random.default_rng(seed)
    target_s = rng.random((n_rays, 3), dtype=np.float32)
    target_h = (near + (far - near) * rng.random((K, n_rays, 1), dtype=np.float32)).astype(np.float32)
    return target_s, target_h


def bounding_box(far=FAR):
    """bb_center = 0, bb_scale = 2/(2*far): same rule as run_scade_scannet.py:1243-1244."""
    return np.zeros(3, np.float32), np.float32(2.0 / (2.0 * far))


def spiral_poses(n_frames=120, radius=0.3, turns=2.0):
    """Harness-side camera path for BASELINE config 5 (the reference has no spiral generator;
    its video poses come from transforms_video.json)."""
    poses = []
    for f in range(n_frames):
        a = 2.0 * math.pi * turns * f / n_frames
        c2w = np.eye(4, dtype=np.float32)
        c2w[0, 3] = radius * math.cos(a)
        c2w[1, 3] = radius * math.sin(a)
        c2w[2, 3] = 0.1 * radius * math.sin(0.5 * a)
        poses.append(c2w[:3])
    return np.stack(poses, 0)


def make_train_scene(n_img=2, H=48, W=64, K=3, n_u=0, seed=40, near=NEAR, far=FAR, depth_channels=1):
    """A tiny synthetic training scene in the reference's array layouts (data/load_scene.py:243-360): images [n,H,W,3],
    depths [n,H,W,C], valid_depths [n,H,W] bool, poses [n,4,4], intrinsics [n,4], all_hypothesis [n,K,H,W,1] clipped to
    [near, far], optional cached_u [n,H,W,n_u]."""
    rng = np.random.default_rng(seed)
    images = rng.random((n_img, H, W, 3)).astype(np.float32)
    depths = rng.uniform(near, far, (n_img, H, W, depth_channels)).astype(np.float32)
    valid = rng.random((n_img, H, W)) < 0.7
    poses = np.tile(np.eye(4, dtype=np.float32), (n_img, 1, 1))
    for i in range(n_img):
        a = 0.3 * (i + 1)
        poses[i, :3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
        poses[i, :3, 3] = rng.uniform(-0.5, 0.5, 3).astype(np.float32)
    intr = np.tile(np.array([W * 0.9, W * 0.91, W / 2 - 0.5, H / 2 + 0.25], np.float32), (n_img, 1))
    hyp = np.clip(rng.uniform(near - 0.5, far + 0.5, (n_img, K, H, W, 1)), near, far).astype(np.float32)
    cu = rng.random((n_img, H, W, n_u)).astype(np.float32) if n_u else None
    return dict(images=images, depths=depths, valid_depths=valid, poses=poses, intrinsics=intr, all_hypothesis=hyp, cached_u=cu)
