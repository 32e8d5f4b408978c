"""Shared comparison helpers for the parity tests."""
import numpy as np


def rays_close(a, b, tol, frac=0.97, loose=None, name=""):
    """Per-ray comparison that tolerates the reference's own discontinuities.

    sample_pdf (H:378-379) switches ``denom`` to 1 when a cdf step is < 1e-5; an empty bin's step is
    1e-5/sum(w + 1e-5), i.e. within a few fp32 ulps of that threshold, so which side it falls on is
    decided by summation order.  A uniform draw landing in such a bin (probability ~1e-5 per bin and
    sample) moves one z sample by up to a bin width and with it that ray's fine-pass outputs.  Two
    correct fp32 implementations therefore agree tightly on almost every ray and differ visibly on a
    few: require ``frac`` of the rays within ``tol`` (max abs over the ray's values) and, if ``loose``
    is given, all rays within ``loose``.
    """
    a = np.asarray(a, np.float64).reshape(len(a), -1)
    b = np.asarray(b, np.float64).reshape(len(b), -1)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = np.abs(a - b).max(-1)
    ok = (err <= tol).mean()
    assert ok >= frac, f"{name}: only {ok:.3f} of rays within {tol} (median {np.median(err):.3g}, max {err.max():.3g})"
    if loose is not None:
        assert np.nanmax(err) <= loose, f"{name}: max ray error {err.max():.3g} > {loose}"
    return err


# fp32 noise floor of the end-to-end render (oracle fp32 vs the reference fp32 vs an fp64 evaluation, see
# DESIGN.md "Parity"): (mean abs error, max abs error) that two CORRECT fp32 implementations stay within.
# The chain 2^8*pi encoding -> 8-layer MLP -> inverse-CDF resampling amplifies summation-order noise, most of
# all on z samples that fall where the density is ~0 (d sample / d cdf = bin width / pdf).
RENDER_FP32_TOL = {
    "rgb0": (5e-6, 2e-4), "weights0": (1e-6, 2e-4), "depth0": (2e-5, 1e-3), "disp0": (1e-4, 5e-3), "acc0": (2e-6, 5e-5),
    "z_vals": (5e-5, 0.1), "rgb_map": (3e-5, 5e-3), "depth_map": (5e-5, 5e-3), "disp_map": (2e-4, 2e-2),
    "weights": (3e-6, 5e-3), "pred_hyp": (3e-4, 0.3), "acc_map": (2e-6, 5e-5), "z_std": (5e-4, 2e-2),
}


def mean_close(a, b, mean_tol, max_tol, name=""):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = np.abs(a - b)
    assert np.isfinite(err).all(), f"{name}: non-finite difference"
    assert err.mean() <= mean_tol, f"{name}: mean abs err {err.mean():.3g} > {mean_tol}"
    assert err.max() <= max_tol, f"{name}: max abs err {err.max():.3g} > {max_tol}"
    return err


# ---- tensor-core training stash (include/scade_b200.h: scade_mlp_tc_stash_layout) ----------------------------
def stash_swizzle(x):
    """float [T*128, chunks*64] -> uint8 image [T][chunks][128 rows][128 B] in the K-major SWIZZLE_128B layout
    (16-byte piece j of row r at r*128 + ((j ^ (r & 7)) << 4)); the inverse of stash_unswizzle."""
    n, c = x.shape
    T, chunks = n // 128, c // 64
    tiles = np.ascontiguousarray(x.astype(np.float16).reshape(T, 128, chunks, 64).transpose(0, 2, 1, 3))   # [T, chunks, 128, 64]
    pieces = tiles.view(np.uint8).reshape(T, chunks, 128, 8, 16)
    out = np.empty_like(pieces)
    r = np.arange(128)[:, None]
    j = np.arange(8)[None, :]
    out[:, :, r, j ^ (r & 7), :] = pieces[:, :, r, j, :]
    return out.reshape(-1)


def stash_unswizzle(buf, off, T, chunks, bf16=False):
    """uint8 workspace -> float32 [T*128, chunks*64] of the chunk region at byte offset `off` (fp16 activations, or
    bf16 gradients with bf16=True)."""
    raw = np.asarray(buf[off:off + T * chunks * 16384]).reshape(T, chunks, 128, 8, 16)
    r = np.arange(128)[:, None]
    j = np.arange(8)[None, :]
    out = np.ascontiguousarray(raw[:, :, r, j ^ (r & 7), :]).reshape(T, chunks, 128, 128)
    if bf16:
        halfs = (out.view(np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)
    else:
        halfs = out.view(np.float16).astype(np.float32)
    return halfs.transpose(0, 2, 1, 3).reshape(T * 128, chunks * 64)


_SIGN_ELEM = [4 * (k & 7) + (k >> 3) for k in range(32)]      # bit (31 - k) of a mask word <- element _SIGN_ELEM[k]


def sign_mask_words(neg):
    """bool [..., 32 n] (True = negative pre-activation) -> uint32 [..., n] in the kernel's sign_mask32 bit order."""
    neg = np.asarray(neg, bool)
    g = neg.reshape(neg.shape[:-1] + (neg.shape[-1] // 32, 32))
    w = np.zeros(g.shape[:-1], np.uint32)
    for k in range(32):
        w |= g[..., _SIGN_ELEM[k]].astype(np.uint32) << np.uint32(31 - k)
    return w


def decode_sign_mask(words):
    """uint32 [..., n] -> bool [..., 32 n] (True = sign bit set = inactive ReLU)."""
    words = np.asarray(words, np.uint32)
    out = np.zeros(words.shape + (32,), bool)
    for k in range(32):
        out[..., _SIGN_ELEM[k]] = (words >> np.uint32(31 - k)) & np.uint32(1)
    return out.reshape(words.shape[:-1] + (words.shape[-1] * 32,))
