"""Training ray-batch construction: the reference's per-step sampler with the same names, arguments
and return values, built by ONE CUDA kernel for the selected pixels only.

    select_coordinates                               model/run_nerf_helpers.py:279-283
    get_ray_batch_from_one_image                     run_scade_scannet.py:753-770
    get_ray_batch_from_one_image_hypothesis_idx      run_scade_scannet.py:772-827

The pixel choice stays ``np.random.choice`` on the host (H:281) so that a run seeded like the reference (np.random.seed(0),
RS:831) visits the same pixels; everything downstream of the indices happens on the device.  The returned ``batch_rays``
carries the finished [N,11] ray batch as ``batch_rays.scade_ray_batch`` (near / far from ``args`` when present) so that
``render(..., rays=batch_rays)`` does not have to re-assemble it.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def select_indices(H, W, N_rand):
    """The flat pixel indices select_coordinates draws (H:281): np.random.choice(H*W, N_rand, replace=False)."""
    return np.random.choice(H * W, size=[N_rand], replace=False)


def select_coordinates(coords, N_rand):
    """H:279-283 (kept for drop-in use; the fused sampler below only needs select_indices)."""
    coords = torch.reshape(coords, [-1, 2])
    select_inds = np.random.choice(coords.shape[0], size=[N_rand], replace=False)
    return coords[select_inds].long()


def gather_train_batch(H, W, intrinsic, pose, select_inds, image, depth=None, valid_depth=None, hypotheses=None, cached_u=None,
                       mask_corners=False, near=0.0, far=1.0, want_mask=False):
    """scade_gather_train_batch.  image [H,W,3]; depth [H,W,C]; valid_depth [H,W] bool; hypotheses [K,H,W(,1)];
    cached_u [H,W,Nu]; select_inds: int64 numpy array or tensor of flat pixel indices.  Returns a dict."""
    dev = image.device
    if not image.is_cuda:
        raise _lib.ScadeError("scade_b200 kernels need CUDA tensors; there is no CPU path")
    if not torch.is_tensor(select_inds):
        select_inds = torch.from_numpy(np.ascontiguousarray(select_inds, dtype=np.int64))
    sel = select_inds.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
    N = sel.numel()
    f32 = lambda t: None if t is None else _lib.f32(t, dev)
    image = f32(image)
    depth = f32(depth)
    if depth is not None and depth.dim() == 2:
        depth = depth.unsqueeze(-1)
    Cd = 0 if depth is None else depth.shape[-1]
    valid8 = None if valid_depth is None else valid_depth.to(device=dev, dtype=torch.bool).contiguous().view(torch.uint8)
    hyp16 = hypotheses is not None and hypotheses.dtype == torch.float16       # a slice of checkpoint.HypothesisStore
    hyp = hypotheses.to(dev).contiguous() if hyp16 else f32(hypotheses)
    K = 0
    if hyp is not None:
        hyp = hyp.reshape(hyp.shape[0], H, W)
        K = hyp.shape[0]
    cu = f32(cached_u)
    Nu = 0 if cu is None else cu.shape[-1]
    out = {
        "ray_batch": torch.empty((N, 11), dtype=torch.float32, device=dev),
        "batch_rays": torch.empty((2, N, 3), dtype=torch.float32, device=dev),
        "target_s": torch.empty((N, 3), dtype=torch.float32, device=dev),
        "target_d": None if depth is None else torch.empty((N, Cd), dtype=torch.float32, device=dev),
        "target_vd": None if valid8 is None else torch.empty((N,), dtype=torch.uint8, device=dev),
        "target_h": None if hyp is None else torch.empty((K, N, 1), dtype=torch.float32, device=dev),
        "mask": torch.empty((N,), dtype=torch.float32, device=dev) if (mask_corners or want_mask) else None,
        "cached_u": None if cu is None else torch.empty((N, Nu), dtype=torch.float32, device=dev),
    }
    intr = _lib.host_floats([float(v) for v in torch.as_tensor(intrinsic).detach().cpu().reshape(-1)[:4]])
    c2w = _lib.host_floats([float(v) for v in torch.as_tensor(pose).detach().cpu().reshape(-1, 4)[:3].reshape(-1)])
    fn = _lib.load().scade_gather_train_batch_h16 if hyp16 else _lib.load().scade_gather_train_batch
    check(fn(
        int(H), int(W), intr, c2w, ptr(sel), N, float(near), float(far), ptr(image), ptr(depth), int(Cd), ptr(valid8), ptr(hyp),
        int(K), ptr(cu), int(Nu), int(bool(mask_corners)), ptr(out["ray_batch"]), ptr(out["batch_rays"]), ptr(out["target_s"]),
        ptr(out["target_d"]), ptr(out["target_vd"]), ptr(out["target_h"]), ptr(out["mask"]), ptr(out["cached_u"]), stream_ptr()),
        "scade_gather_train_batch")
    if out["target_vd"] is not None:
        out["target_vd"] = out["target_vd"].view(torch.bool)
    return out


def _attach(batch_rays, ray_batch, near, far):
    batch_rays.scade_ray_batch = (ray_batch, float(near), float(far))
    return batch_rays


def get_ray_batch_from_one_image(H, W, i_train, images, depths, valid_depths, poses, intrinsics, args):
    """RS:753-770 -> (batch_rays [2,N,3], target_s, target_d, target_vd, img_i)."""
    img_i = np.random.choice(i_train)
    near, far = getattr(args, "near", 0.0), getattr(args, "far", 1.0)
    o = gather_train_batch(H, W, intrinsics[img_i, :], poses[img_i], select_indices(H, W, args.N_rand), images[img_i],
                           depths[img_i], valid_depths[img_i], near=near, far=far)
    return _attach(o["batch_rays"], o["ray_batch"], near, far), o["target_s"], o["target_d"], o["target_vd"], img_i


def get_ray_batch_from_one_image_hypothesis_idx(H, W, img_i, images, depths, valid_depths, poses, intrinsics, all_hypothesis, args,
                                                space_carving_idx=None, cached_u=None):
    """RS:772-827 -> (batch_rays, target_s, target_d, target_vd, img_i, target_h [K,N,1], space_carving_mask, curr_cached_u).
    ``space_carving_idx`` (RS:793-802) is never passed by the reference's training loop (RS:951-952 passes None) and is not
    supported."""
    if space_carving_idx is not None:
        raise NotImplementedError("space_carving_idx is dead code on the reference's training path (RS:951-952 passes None)")
    near, far = getattr(args, "near", 0.0), getattr(args, "far", 1.0)
    o = gather_train_batch(H, W, intrinsics[img_i, :], poses[img_i], select_indices(H, W, args.N_rand), images[img_i],
                           depths[img_i], valid_depths[img_i], all_hypothesis[img_i],
                           None if cached_u is None else cached_u[img_i], mask_corners=bool(getattr(args, "mask_corners", False)),
                           near=near, far=far)
    return (_attach(o["batch_rays"], o["ray_batch"], near, far), o["target_s"], o["target_d"], o["target_vd"], img_i, o["target_h"],
            o["mask"] if getattr(args, "mask_corners", False) else None, o["cached_u"])
