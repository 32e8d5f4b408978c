"""Checkpoint and depth-hypothesis formats of the reference.

    checkpoint .tar      run_scade_scannet.py:1004-1019 (save), 412-420 + 476-487 (load)
    depth hypotheses     data/load_scene.py:319-348  (<img_id>_<k>.npy, clipped to [near, far])

The reference wraps both networks in nn.DataParallel (RS:438,455), so its state-dict keys carry a ``module.`` prefix; our
NeRF modules are not wrapped.  ``load_state`` accepts both spellings, ``save_checkpoint`` writes the reference's (prefixed)
spelling, so files move between the two programs unchanged.
"""
from __future__ import annotations

import os

import numpy as np
import torch


def _strip(sd):
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}


def _unwrap(net):
    return net.module if hasattr(net, "module") else net


def load_state(net, state_dict):
    """model.load_state_dict(ckpt['network_fn_state_dict']) (RS:483) for wrapped or unwrapped modules and keys."""
    _unwrap(net).load_state_dict(_strip(state_dict))


def find_latest(ckpt_dir, expname):
    """load_checkpoint's search (RS:412-420): the lexicographically last '*000.tar' under ckpt_dir/expname, or None."""
    path = os.path.join(ckpt_dir, expname)
    if not os.path.isdir(path):
        return None
    ckpts = [os.path.join(path, f) for f in sorted(os.listdir(path)) if "000.tar" in f]
    return ckpts[-1] if ckpts else None


def warm_packed_weights(*nets):
    """Checkpoint load -> packed weight-stream cache: builds the tensor-core stream (scade_mlp_pack) of every CUDA network whose
    precision is tc_f16 / tc_f16x3 right away, so the first render after a load does not pay for it.  Returns the streams."""
    from . import functional as F_
    out = []
    for net in nets:
        net = _unwrap(net)
        prec = F_.PRECISIONS.get(getattr(net, "precision", "fp32"))
        if prec in F_.TC_PRECISIONS and next(net.parameters()).is_cuda and net.handle().tc_supported():
            out.append(net.handle().packed(prec))
    return out


def load_checkpoint(path, network_fn, network_fine=None, optimizer=None, map_location=None, warm=True):
    """Returns (global_step, extras) after loading both networks (RS:476-487).  The reference leaves the optimizer state
    alone (RS:480 is commented out); pass `optimizer` to restore it as well.  extras: depth_scales / depth_shifts /
    embedded_cam when the file has them (RS:1014-1017).  warm: re-pack the tensor-core weight streams now (warm_packed_weights)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    load_state(network_fn, ckpt["network_fn_state_dict"])
    if network_fine is not None and "network_fine_state_dict" in ckpt:
        load_state(network_fine, ckpt["network_fine_state_dict"])
    if warm:
        warm_packed_weights(*[n for n in (network_fn, network_fine) if n is not None])
    if optimizer is not None and "optimizer_state_dict" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    extras = {k: ckpt[k] for k in ("depth_scales", "depth_shifts", "embedded_cam") if k in ckpt}
    return int(ckpt.get("global_step", 0)), extras


def save_checkpoint(path, global_step, network_fn, network_fine=None, optimizer=None, depth_scales=None, depth_shifts=None):
    """RS:1004-1019: same keys, state-dict keys prefixed with ``module.`` like the reference's DataParallel modules."""
    pref = lambda net: {"module." + k: v.detach().clone() for k, v in _unwrap(net).state_dict().items()}
    d = {"global_step": int(global_step), "network_fn_state_dict": pref(network_fn)}
    if optimizer is not None:
        d["optimizer_state_dict"] = optimizer.state_dict()
    if network_fine is not None:
        d["network_fine_state_dict"] = pref(network_fine)
    if depth_shifts is not None:
        d["depth_shifts"] = depth_shifts
    if depth_scales is not None:
        d["depth_scales"] = depth_scales
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(d, path)
    return path


def load_depth_hypotheses(leres_dir, img_ids, num_hypothesis, near, far, device=None, pin=True):
    """data/load_scene.py:319-348: [n_img, K, H, W, 1] float32, clipped to [near, far].  Read into pinned host memory and, if
    `device` is given, moved to the GPU once: the training sampler (scade_b200.sampler) gathers K values per selected pixel
    from the resident tensor every step (4 K H W bytes per image: 24.6 MB for 20 x 480 x 640)."""
    first = np.load(os.path.join(leres_dir, f"{img_ids[0]}_0.npy"))
    H, W = first.shape[:2]
    out = torch.empty((len(img_ids), num_hypothesis, H, W, 1), dtype=torch.float32,
                      pin_memory=bool(pin and torch.cuda.is_available()))
    for i, img_id in enumerate(img_ids):
        for j in range(num_hypothesis):
            d = np.load(os.path.join(leres_dir, f"{img_id}_{j}.npy")).astype(np.float32)
            out[i, j, :, :, 0] = torch.from_numpy(np.clip(d.reshape(H, W), near, far))
    return out.to(device, non_blocking=True) if device is not None else out


class HypothesisStore:
    """The K depth hypotheses of every training image as ONE resident fp16 tensor [n_img, K, H, W], on the
    GPU or in pinned host memory: half the bytes of the reference's fp32 array (load_scene.py:319-348; 20 hypotheses of a
    640 x 480 image: 12.3 MB instead of 24.6 MB).  ``store[img_i]`` is the [K, H, W] half tensor the training sampler takes as
    `all_hypothesis[img_i]` (scade_b200.sampler gathers the N_rand pixels with scade_gather_train_batch_h16; the target_h it
    returns holds the exact fp32 value of each stored half).  Depths live in [near, far] ~ [0.1, 10] m: fp16 keeps 11
    significant bits (<= 2.4 mm at 5 m), below the noise of the monocular hypotheses."""

    def __init__(self, data_f16):
        if data_f16.dtype != torch.float16 or data_f16.dim() != 4:
            raise ValueError("HypothesisStore holds a [n_img, K, H, W] float16 tensor")
        self.data = data_f16

    @classmethod
    def from_float(cls, hyp, near, far, device=None):
        """hyp: [n_img, K, H, W(, 1)] float32 (numpy or tensor, host or device).  With a CUDA `device` the clip + conversion runs
        on the GPU (scade_pack_hypotheses_f16), image by image, so the fp32 copy is never resident as a whole."""
        from . import _lib
        hyp = torch.as_tensor(hyp)
        if hyp.dim() == 5:
            hyp = hyp[..., 0]
        n_img = hyp.shape[0]
        if device is not None and torch.device(device).type == "cuda":
            out = torch.empty(hyp.shape, dtype=torch.float16, device=device)
            for i in range(n_img):
                src = hyp[i].to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
                _lib.check(_lib.load().scade_pack_hypotheses_f16(_lib.ptr(src), src.numel(), float(near), float(far), _lib.ptr(out[i]),
                                                                 _lib.stream_ptr()), "scade_pack_hypotheses_f16")
            return cls(out)
        out = hyp.float().clamp(float(near), float(far)).to(torch.float16)
        return cls(out.pin_memory() if torch.cuda.is_available() else out)

    @classmethod
    def from_files(cls, leres_dir, img_ids, num_hypothesis, near, far, device=None):
        """<img_id>_<k>.npy files (data/load_scene.py:319-348) -> store, one image at a time."""
        first = np.load(os.path.join(leres_dir, f"{img_ids[0]}_0.npy"))
        H, W = first.shape[:2]
        parts = []
        for img_id in img_ids:
            one = np.stack([np.load(os.path.join(leres_dir, f"{img_id}_{j}.npy")).astype(np.float32).reshape(H, W)
                            for j in range(num_hypothesis)], 0)
            parts.append(cls.from_float(one[None], near, far, device).data)
        return cls(torch.cat(parts, 0))

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, img_i):
        return self.data[img_i]

    @property
    def nbytes(self):
        return self.data.numel() * 2
