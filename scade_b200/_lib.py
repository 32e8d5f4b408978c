"""ctypes binding of libscade_b200.so (include/scade_b200.h).

The library is the product: if it cannot be built or loaded this module raises, and nothing in
the package falls back to a CPU or eager-PyTorch path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

from . import build as _build

MAX_PARAM_TENSORS = 40
PREC_FP32 = 0
PREC_TC_F16 = 1
PREC_TC_F16X3 = 2


class NetDesc(Structure):
    _fields_ = [("D", c_int32), ("W", c_int32), ("multires", c_int32), ("multires_views", c_int32), ("skip", c_int32)]


class Net(Structure):
    _fields_ = [("desc", NetDesc), ("params", c_void_p * MAX_PARAM_TENSORS), ("packed_f16", c_void_p)]


class RenderCfg(Structure):
    _fields_ = [("N_samples", c_int32), ("N_importance", c_int32), ("lindisp", c_int32), ("precision", c_int32),
                ("is_joint", c_int32), ("ray_stride", c_int32), ("bb_center", c_float * 3), ("bb_scale", c_float)]


RENDER_OUT_FIELDS = ["rgb_map", "disp_map", "acc_map", "depth_map", "z_vals", "weights", "pred_hyp", "u", "raw",
                     "rgb0", "disp0", "acc0", "depth0", "z_vals0", "weights0", "z_std"]


class RenderOut(Structure):
    _fields_ = [(k, c_void_p) for k in RENDER_OUT_FIELDS]


class ScadeError(RuntimeError):
    pass


_P = c_void_p
_SIGNATURES = {
    "scade_version": (c_int, []),
    "scade_last_error_string": (c_char_p, []),
    "scade_kernel_launch_count": (ctypes.c_uint64, []),
    "scade_mlp_packed_bytes": (c_size_t, [POINTER(NetDesc)]),
    "scade_mlp_pack_f16": (c_int, [POINTER(Net), _P, _P]),
    "scade_mlp_packed_bytes_for": (c_size_t, [POINTER(NetDesc), c_int]),
    "scade_mlp_pack": (c_int, [POINTER(Net), c_int, _P, _P]),
    "scade_mlp_workspace_bytes": (c_size_t, [POINTER(NetDesc), c_int64, c_int, c_int]),
    "scade_mlp_forward_rays": (c_int, [POINTER(Net), c_int, _P, c_int, _P, c_int64, c_int, POINTER(c_float), c_float,
                                       _P, _P, c_size_t, c_int, _P]),
    "scade_mlp_forward_rays_composite_supported": (c_int, [POINTER(NetDesc), c_int, c_int]),
    "scade_mlp_forward_rays_composite": (c_int, [POINTER(Net), c_int, _P, c_int, _P, c_int64, c_int, POINTER(c_float), c_float,
                                                 _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "scade_mlp_forward_embedded": (c_int, [POINTER(Net), c_int, _P, c_int64, _P, _P, c_size_t, c_int, _P]),
    "scade_mlp_backward": (c_int, [POINTER(Net), c_int, _P, c_int64, POINTER(c_void_p), _P, c_size_t, _P]),
    "scade_mlp_tc_stash_layout": (c_int, [POINTER(NetDesc), c_int64, POINTER(c_int64), c_int]),
    "scade_mlp_composite_plan": (c_int, [c_int, c_int64, c_int, POINTER(c_int), POINTER(c_int)]),
    "scade_embed": (c_int, [_P, c_int64, c_int, _P, _P]),
    "scade_get_rays": (c_int, [c_int, c_int, POINTER(c_float), POINTER(c_float), c_int, c_int, _P, _P, _P]),
    "scade_make_ray_batch": (c_int, [_P, _P, c_int64, c_float, c_float, _P, _P]),
    "scade_camera_ray_batch": (c_int, [c_int, c_int, POINTER(c_float), POINTER(c_float), c_int, c_int, c_int64,
                                       c_int64, c_float, c_float, _P, _P]),
    "scade_coarse_z_vals": (c_int, [_P, c_int, c_int64, c_int, c_int, _P, _P, _P]),
    "scade_perturb_z_vals": (c_int, [_P, _P, c_int64, c_int, _P, _P]),
    "scade_raw2outputs": (c_int, [_P, _P, _P, c_int, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P]),
    "scade_raw2outputs_backward": (c_int, [_P, _P, _P, c_int, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "scade_sample_pdf": (c_int, [_P, _P, c_int64, c_int, c_int, _P, c_int, _P, _P, _P]),
    "scade_sample_pdf_backward": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    "scade_resample_from_z": (c_int, [_P, _P, c_int64, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P]),
    "scade_composite_resample": (c_int, [_P, _P, _P, c_int, c_int64, c_int, _P, _P, _P, _P, _P, c_int, _P, c_int, _P, _P, _P, _P, _P]),
    "scade_resample_from_z_backward": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P, c_int, _P]),
    "scade_sort_merge": (c_int, [_P, c_int, _P, c_int, c_int64, _P, _P]),
    "scade_space_carving_workspace_bytes": (c_size_t, [c_int, c_int64, c_int]),
    "scade_space_carving_loss": (c_int, [_P, _P, c_int, _P, c_int, c_int64, c_int, c_int, c_float, c_float, _P, _P, _P,
                                         _P, c_size_t, _P]),
    "scade_space_carving_loss_affine": (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, c_int, c_float, c_float, c_int64, _P, _P, _P, _P,
                                                c_int, _P]),
    "scade_space_carving_joint_accumulate": (c_int, [_P, _P, c_int, _P, c_int, c_int64, c_int, c_float, _P, _P]),
    "scade_space_carving_joint_finish": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int64, c_int64, c_int, c_float, c_float, _P, _P, _P,
                                                 _P, _P]),
    "scade_img2mse": (c_int, [_P, _P, c_int64, c_int64, c_float, _P, _P, _P]),
    "scade_gather_train_batch": (c_int, [c_int, c_int, POINTER(c_float), POINTER(c_float), _P, c_int64, c_float, c_float, _P, _P,
                                         c_int, _P, _P, c_int, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "scade_gather_train_batch_h16": (c_int, [c_int, c_int, POINTER(c_float), POINTER(c_float), _P, c_int64, c_float, c_float, _P, _P,
                                             c_int, _P, _P, c_int, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "scade_pack_hypotheses_f16": (c_int, [_P, c_int64, c_float, c_float, _P, _P]),
    "scade_video_frame": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P, _P]),
    "scade_adam_step": (c_int, [_P, _P, _P, _P, c_int64, c_double, c_double, c_double, c_double, c_int64, _P]),
    "scade_adam_step_graph": (c_int, [_P, _P, _P, _P, c_int64, _P, c_double, c_double, c_double, _P, _P]),
    "scade_render_rays_workspace_bytes": (c_size_t, [POINTER(RenderCfg), POINTER(NetDesc), POINTER(NetDesc), c_int64]),
    "scade_render_rays_forward": (c_int, [POINTER(RenderCfg), _P, c_int64, POINTER(Net), POINTER(Net), _P, _P, _P,
                                          POINTER(RenderOut), _P, c_size_t, _P]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Build (if stale) and dlopen the library; declare every prototype of the header."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return list(_SIGNATURES)


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().scade_last_error_string().decode("utf-8", "replace")
        raise ScadeError(f"{what or 'scade call'} failed with status {status}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be CUDA and contiguous (dtype is the caller's business)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ScadeError("scade_b200 kernels need CUDA tensors; there is no CPU path")
    if not t.is_contiguous():
        raise ScadeError("tensor must be contiguous")
    return c_void_p(t.data_ptr())


def f32(t, device=None):
    """Contiguous fp32 CUDA view/copy of t."""
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if device is not None and t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def host_floats(values):
    arr = (c_float * len(values))(*[float(v) for v in values])
    return arr
