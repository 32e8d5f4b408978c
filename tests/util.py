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
