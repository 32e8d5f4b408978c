"""Video / eval post-processing on the device.

    depth_std (RS:257-258), the three-panel video frame of render_video (RS:249-260), to16b(depth) (H:14, RS:403)

``video_frame`` returns the finished [H, 3W, 3] uint8 BGR frame (ready for cv2.imwrite, RS:260) as a CUDA tensor; the only
thing left on the host is the file write.  Colour tables are cv2's own (COLORMAP_TURBO / COLORMAP_VIRIDIS) when cv2 is
importable, otherwise grey.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, f32, ptr, stream_ptr

_LUTS = {}


def colormap_lut(name, device):
    """[256,3] uint8 BGR table of cv2.applyColorMap(., COLORMAP_<name>) on `device`, or None without cv2."""
    key = (name, str(device))
    if key not in _LUTS:
        try:
            import cv2
            lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), getattr(cv2, "COLORMAP_" + name)).reshape(256, 3)
            _LUTS[key] = torch.from_numpy(np.ascontiguousarray(lut)).to(device)
        except Exception:
            _LUTS[key] = None
    return _LUTS[key]


def video_frame(rgb, depth_map, z_vals, weights, depth_scale, lut_depth="TURBO", lut_std="VIRIDIS", want_frame=True,
                want_std=True, want_depth16=False):
    """rgb [H,W,3], depth_map [H,W], z_vals / weights [H,W,S] -> dict(frame uint8 [H,3W,3] BGR, depth_std [H,W],
    depth16 int32 view of uint16 values [H,W])."""
    rgb, depth_map, z_vals, weights = f32(rgb), f32(depth_map), f32(z_vals), f32(weights)
    H, W = depth_map.shape
    S = z_vals.shape[-1]
    dev = depth_map.device
    ld = colormap_lut(lut_depth, dev) if isinstance(lut_depth, str) else lut_depth
    ls = colormap_lut(lut_std, dev) if isinstance(lut_std, str) else lut_std
    frame = torch.empty((H, 3 * W, 3), dtype=torch.uint8, device=dev) if want_frame else None
    std = torch.empty((H, W), dtype=torch.float32, device=dev) if want_std else None
    d16 = torch.empty((H, W), dtype=torch.int16, device=dev) if want_depth16 else None
    check(_lib.load().scade_video_frame(ptr(rgb), ptr(depth_map), ptr(z_vals), ptr(weights), H, W, S, float(depth_scale), ptr(ld),
                                        ptr(ls), ptr(frame), ptr(std), ptr(d16), stream_ptr()), "scade_video_frame")
    out = {"frame": frame, "depth_std": std}
    if d16 is not None:
        out["depth16"] = d16.to(torch.int32) & 0xFFFF          # uint16 values (torch has no uint16 arithmetic)
    return out


def depth_std(z_vals, weights, depth_map):
    """RS:257-258: sqrt(clamp(sum((z - depth)^2 * w), 0, 1)) for maps of any leading shape."""
    lead = depth_map.shape
    z2, w2 = f32(z_vals).reshape(-1, z_vals.shape[-1]), f32(weights).reshape(-1, z_vals.shape[-1])
    d2 = f32(depth_map).reshape(1, -1)
    std = torch.empty(d2.shape, dtype=torch.float32, device=d2.device)
    check(_lib.load().scade_video_frame(None, ptr(d2), ptr(z2), ptr(w2), 1, d2.shape[1], z2.shape[-1], 1.0, None, None, None,
                                        ptr(std), None, stream_ptr()), "scade_video_frame")
    return std.reshape(lead)
