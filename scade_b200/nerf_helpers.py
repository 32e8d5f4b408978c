"""Host-side mirror of the reference's ``model/run_nerf_helpers.py`` for the hot path.

Same names, arguments and return conventions as the reference so that its training / test /
video loops can import these instead; every call lands in a CUDA kernel of
libscade_b200.so.  CPU tensors raise -- there is no fallback.

    NeRF, DenseLayer            H:131-139, H:193-247
    Embedder / get_embedder     H:142-189
    sample_pdf*                 H:337-538
    compute_space_carving_loss  H:93-128
    get_rays                    H:285-305
    img2mse, mse2psnr           H:11-12
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import functional as F_


def img2mse(x, y):
    """H:11"""
    return F_.img2mse(x, y)


def mse2psnr(x):
    """H:12"""
    return -10.0 * torch.log(x) / math.log(10.0)


to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)                 # H:13
to16b = lambda x: ((2 ** 16 - 1) * np.clip(x, 0, 1)).astype(np.uint16)     # H:14


class DenseLayer(nn.Linear):
    """Parameter container with the reference's initialisation (H:131-139): Xavier-uniform with the
    gain of the following activation, zero bias.  forward() is never used on the hot path -- the
    owning NeRF module runs all layers in one fused call."""

    def __init__(self, in_dim, out_dim, activation="relu", *args, **kwargs):
        self.activation = activation
        super().__init__(in_dim, out_dim, *args, **kwargs)

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight, gain=nn.init.calculate_gain(self.activation))
        if self.bias is not None:
            nn.init.zeros_(self.bias)


class Embedder:
    """Positional encoding object with the reference's kwargs interface (H:142-172)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        if not (kwargs.get("include_input", True) and kwargs.get("log_sampling", True) and kwargs.get("input_dims", 3) == 3):
            raise NotImplementedError("only the configuration get_embedder() builds is supported (H:178-185)")
        self.num_freqs = int(kwargs["num_freqs"])
        self.out_dim = 3 + 6 * self.num_freqs

    def embed(self, inputs):
        return F_.embed(inputs, self.num_freqs)


class _EmbedFn:
    """Callable returned by get_embedder; carries ``multires`` so run_network can fuse it."""

    def __init__(self, multires):
        self.multires = multires
        self.embedder = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                                 log_sampling=True, periodic_fns=[torch.sin, torch.cos])

    def __call__(self, x):
        return self.embedder.embed(x)


def get_embedder(multires, i=0):
    """H:174-189 -> (embed_fn, out_dim)."""
    if i == -1:
        return nn.Identity(), 3
    fn = _EmbedFn(multires)
    return fn, fn.embedder.out_dim


class NeRF(nn.Module):
    """Same constructor, parameters and state_dict keys as the reference NeRF (H:193-221); forward()
    (H:223-247) is one call into the CUDA library.

    ``precision``: "fp32" (FFMA GEMMs, the reference's arithmetic) or "tc_f16" (tcgen05 tensor cores).
    """

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, input_ch_cam=0, output_ch=4, skips=[4],
                 use_viewdirs=False, precision="fp32"):
        super().__init__()
        if not use_viewdirs:
            raise NotImplementedError("the SCADE hot path always runs use_viewdirs=True (RS:1141)")
        if input_ch_cam != 0:
            raise NotImplementedError("input_ch_cam > 0 is dead code in the reference (SURVEY App. C)")
        if len(skips) > 1:
            raise NotImplementedError("one skip connection (reference: skips=[4])")
        if (input_ch - 3) % 6 or (input_ch_views - 3) % 6:
            raise ValueError("input_ch / input_ch_views must come from get_embedder (3 + 6*multires)")
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views, self.input_ch_cam = input_ch, input_ch_views, input_ch_cam
        self.skips, self.use_viewdirs, self.precision = list(skips), use_viewdirs, precision
        self.pts_linears = nn.ModuleList(
            [DenseLayer(input_ch, W, activation="relu")]
            + [DenseLayer(W + input_ch if i in self.skips else W, W, activation="relu") for i in range(D - 1)])
        self.views_linears = nn.ModuleList([DenseLayer(input_ch_views + input_ch_cam + W, W // 2, activation="relu")])
        self.feature_linear = DenseLayer(W, W, activation="linear")
        self.alpha_linear = DenseLayer(W, 1, activation="linear")
        self.rgb_linear = DenseLayer(W // 2, 3, activation="linear")
        self._handle = None
        self._handle_key = None

    def ordered_parameters(self):
        ps = []
        for layer in self.pts_linears:
            ps += [layer.weight, layer.bias]
        for layer in (self.views_linears[0], self.feature_linear, self.alpha_linear, self.rgb_linear):
            ps += [layer.weight, layer.bias]
        return ps

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() move the parameter storage: drop the cached C-side view
        self._handle = None
        return super()._apply(fn, *args, **kwargs)

    def handle(self):
        """C-side view of the parameters (cached: this sits on the per-call path of render_rays).  The cache is checked
        against the identity and storage of the first and last parameter; `_apply` and load_state_dict paths that move storage
        reset it."""
        h = self._handle
        if h is not None:
            first, last = self.pts_linears[0].weight, self.rgb_linear.bias
            if h.params[0] is first and h.params[-1] is last and self._handle_key == (first.data_ptr(), last.data_ptr()):
                return h
        ps = self.ordered_parameters()
        skip = self.skips[0] if self.skips else -1
        self._handle = F_.NetHandle(ps, self.D, self.W, (self.input_ch - 3) // 6, (self.input_ch_views - 3) // 6, skip)
        self._handle_key = (ps[0].data_ptr(), ps[-1].data_ptr())
        return self._handle

    def forward(self, x):
        return F_.mlp_forward_embedded(self.handle(), x, self.precision)


def get_rays(H, W, intrinsic, c2w, coords=None):
    """H:285-305 (full image only on the hot path, RS:108)."""
    if coords is not None:
        raise NotImplementedError("per-coordinate ray generation belongs to the training sampler (RS:772-827)")
    return F_.get_rays(int(H), int(W), intrinsic, c2w)


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """H:337-383"""
    return sample_pdf_return_u(bins, weights, N_samples, det=det, pytest=pytest)[0]


def _draw_u(shape, det, pytest, device):
    if pytest:  # H:352-361: numpy's seed-0 stream replaces torch's
        np.random.seed(0)
        if det:
            return torch.tensor(np.broadcast_to(np.linspace(0.0, 1.0, shape[-1]), shape).copy(), dtype=torch.float32,
                                device=device)
        return torch.tensor(np.random.rand(*shape), dtype=torch.float32, device=device)
    if det:
        return None             # kernel forms linspace(0,1,N) itself (H:347)
    return torch.rand(shape, device=device)


def sample_pdf_return_u(bins, weights, N_samples, det=False, pytest=False, load_u=None):
    """H:385-436 -> (samples, u)"""
    u = load_u if load_u is not None else _draw_u(tuple(bins.shape[:-1]) + (N_samples,), det, pytest, bins.device)
    return F_.sample_pdf(bins, weights, N_samples, u=u, joint=False)


def sample_pdf_joint(bins, weights, N_samples, det=False, pytest=False):
    """H:439-486"""
    return sample_pdf_joint_return_u(bins, weights, N_samples, det=det, pytest=pytest)[0]


def sample_pdf_joint_return_u(bins, weights, N_samples, det=False, pytest=False, load_u=None):
    """H:488-538: one row of uniforms shared by every ray (H:502-503)."""
    if load_u is not None:
        return F_.sample_pdf(bins, weights, N_samples, u=load_u, joint=False)
    if pytest:
        u = _draw_u(tuple(bins.shape[:-1]) + (N_samples,), det, True, bins.device)
        return F_.sample_pdf(bins, weights, N_samples, u=u, joint=False)
    if det:
        return F_.sample_pdf(bins, weights, N_samples, u=None)
    return F_.sample_pdf(bins, weights, N_samples, u=torch.rand(N_samples, device=bins.device), joint=True)


def compute_space_carving_loss(pred_depth, target_hypothesis, is_joint=False, mask=None, norm_p=2, threshold=0.0):
    """H:93-128.  norm_p is accepted and irrelevant: the norm runs over a singleton axis (H:106)."""
    return F_.space_carving_loss(pred_depth, target_hypothesis, is_joint=is_joint, mask=mask, threshold=threshold)
