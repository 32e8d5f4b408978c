"""Tensor-level wrappers over the C ABI (include/scade_b200.h) with autograd support.

Every function here launches hand-written CUDA kernels from libscade_b200.so on the current
torch stream; torch is used for device memory, streams and the autograd tape only.  There is no
CPU or eager fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
import math
from ctypes import byref, c_void_p

import torch

from . import _lib
from ._lib import PREC_FP32, PREC_TC_F16, PREC_TC_F16X3, check, f32, ptr, stream_ptr

# "fp32": FFMA GEMMs (the reference's arithmetic).  "tc_f16": tcgen05, fp16 operands (fast mode, forward + backward).
# "tc_f16x3": tcgen05 at an fp32-level tolerance, fp16 (hi, lo) operand pairs, three MMA passes (tight mode, forward only).
PRECISIONS = {"fp32": PREC_FP32, "tc_f16": PREC_TC_F16, "tc_f16x3": PREC_TC_F16X3, PREC_FP32: PREC_FP32,
              PREC_TC_F16: PREC_TC_F16, PREC_TC_F16X3: PREC_TC_F16X3}
TC_PRECISIONS = (PREC_TC_F16, PREC_TC_F16X3)


def _L():
    return _lib.load()


_weights_epoch = 0


def mark_weights_changed():
    """Parameters were updated behind autograd's back (a raw kernel such as scade_adam_step does not bump tensor version
    counters): every NetHandle re-packs its fp16 tile image on next use."""
    global _weights_epoch
    _weights_epoch += 1


def _bytes(n, device):
    return torch.empty(max(int(n), 256), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------------------------
# network handle
# ----------------------------------------------------------------------------------------------
class NetHandle:
    """C-side view (scade_net) of a NeRF module's parameters, in the reference state_dict order
    (model/run_nerf_helpers.py:206-219).  Keeps the fp16 tile image for the tensor-core path in sync
    with the fp32 master weights (re-packed whenever a parameter's version counter moved, i.e. after
    optimizer.step() or load_state_dict())."""

    def __init__(self, params, D, W, multires, multires_views, skip):
        self.params = list(params)
        if len(self.params) != 2 * D + 8:
            raise ValueError(f"expected {2 * D + 8} parameter tensors, got {len(self.params)}")
        self.desc = _lib.NetDesc(D, W, multires, multires_views, skip)
        self._packed = {}                # precision -> packed weight stream (uint8 tensor)
        self._packed_key = {}            # precision -> key the stream was packed for
        self._struct_key = None          # (data pointers) the cached scade_net structs were built from
        self._structs = {}

    def tc_supported(self):
        return _L().scade_mlp_packed_bytes(byref(self.desc)) > 0

    def struct(self, precision):
        """scade_net for this call.  The struct (24 pointers) is cached per precision and rebuilt only when a parameter's
        storage moved; the fp16 stream is re-packed by packed() when a version counter moved."""
        ptrs = tuple(p.data_ptr() for p in self.params)
        if ptrs != self._struct_key:
            for p in self.params:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.ScadeError("network parameters must be contiguous fp32 CUDA tensors")
            self._struct_key, self._structs = ptrs, {}
        net = self._structs.get(precision)
        if net is None:
            net = _lib.Net()
            net.desc = self.desc
            for i, ptr_ in enumerate(ptrs):
                net.params[i] = ptr_
            net.packed_f16 = None
            self._structs[precision] = net
        if precision in TC_PRECISIONS:
            net.packed_f16 = self.packed(precision).data_ptr()
        return net

    def invalidate_packed(self):
        """Force a re-pack of every tensor-core weight stream on next use (e.g. right before a CUDA-graph capture, so that
        the pack launches are recorded inside the graph)."""
        self._packed_key = {}

    def packed(self, precision=PREC_TC_F16):
        """The packed fp16 weight stream of a tensor-core precision, re-packed (scade_mlp_pack) whenever a parameter's version
        counter, its storage or the global weights epoch moved."""
        key = (_weights_epoch, self._struct_key if self._struct_key is not None else tuple(p.data_ptr() for p in self.params)) \
            + tuple(p._version for p in self.params)
        buf = self._packed.get(precision)
        if buf is None or key != self._packed_key.get(precision):
            nbytes = _L().scade_mlp_packed_bytes_for(byref(self.desc), precision)
            if nbytes == 0:
                raise _lib.ScadeError("this network shape is not supported by the tensor-core (tc_f16 / tc_f16x3) paths")
            if buf is None or buf.numel() != nbytes or buf.device != self.params[0].device:
                buf = self._packed[precision] = torch.zeros(nbytes, dtype=torch.uint8, device=self.params[0].device)   # (alignment padding stays 0)
            net = self.struct(PREC_FP32)
            check(_L().scade_mlp_pack(byref(net), precision, ptr(buf), stream_ptr()), "scade_mlp_pack")
            self._packed_key[precision] = key
        return buf

    def workspace_bytes(self, P, precision, save):
        return _L().scade_mlp_workspace_bytes(byref(self.desc), int(P), int(precision), int(save))


def tc_train_enabled():
    """SCADE_TC_TRAIN=0 sends training through the fp32 FFMA GEMMs even when the network's precision is tc_f16."""
    import os
    return os.environ.get("SCADE_TC_TRAIN", "1") != "0"


def stash_layout(handle, P):
    """Offsets of the tensor-core training stash inside the forward workspace (scade_mlp_tc_stash_layout)."""
    buf = (ctypes.c_int64 * 64)()
    n = _L().scade_mlp_tc_stash_layout(byref(handle.desc), int(P), buf, 64)
    v = list(buf[:n])
    names = ["T", "D", "total", "emb", "feat", "hv", "dzv", "dzf", "maskv", "alpha", "gs"]
    out = dict(zip(names, v[:11]))
    out["h"], out["dz"], out["maskh"] = v[11:19], v[19:27], v[27:35]
    return out


_direct_grad_refs = {}         # id(param) -> weakref of parameters whose owner opted in to in-place gradient accumulation


def enable_direct_grads(params, on=True):
    """Opt in (scade_b200.optim.FlatParams does) to having the backward kernels accumulate straight into ``p.grad`` instead of
    returning gradients through autograd.  Explicit because it bypasses autograd's own accumulation: tensor hooks on the
    parameters do not fire and torch.autograd.grad() would still write .grad."""
    import weakref
    for p in params:
        if on:
            _direct_grad_refs[id(p)] = weakref.ref(p, lambda _r, k=id(p): _direct_grad_refs.pop(k, None))
        else:
            _direct_grad_refs.pop(id(p), None)


def _direct_grad(p):
    r = _direct_grad_refs.get(id(p))
    return r is not None and r() is p              # (an id can be recycled after the registered tensor died)


def _mlp_backward(handle, precision, d_out, P, ws, device):
    """scade_mlp_backward ACCUMULATES.  For parameters that opted in (enable_direct_grads: FlatParams views), all of which
    require grad and own a contiguous fp32 .grad, the kernels add straight into .grad and autograd gets None; otherwise one
    zeroed flat buffer is carved into per-parameter gradients and returned through autograd.  `handle.grad_ready_hook`, if
    set, is called once the backward kernels of this network are enqueued (scade_b200.dist launches that network's gradient
    bucket all-reduce from it, overlapping the rest of the backward)."""
    params = handle.params
    direct = all(_direct_grad(p) and p.requires_grad and p.grad is not None and p.grad.is_contiguous()
                 and p.grad.dtype == torch.float32 and p.grad.is_cuda for p in params)
    if direct:
        grads = [p.grad for p in params]
    else:
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
        grads, off = [], 0
        for p, n in zip(params, sizes):
            grads.append(flat[off:off + p.numel()].view(p.shape))
            off += n
    net = handle.struct(precision)
    arr = (c_void_p * len(grads))(*[g.data_ptr() for g in grads])
    check(_L().scade_mlp_backward(byref(net), precision, ptr(d_out), P, arr, ptr(ws), ws.numel(), stream_ptr()),
          "scade_mlp_backward")
    hook = getattr(handle, "grad_ready_hook", None)
    if hook is not None and direct:
        hook(handle)
    return [None] * len(params) if direct else grads


class _MLPRaysFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays, z_vals, handle, precision, bb_center, bb_scale, need_grad, *params):
        N, S = z_vals.shape
        raw = torch.empty((N, S, 4), dtype=torch.float32, device=z_vals.device)
        ws = _bytes(handle.workspace_bytes(N * S, precision, need_grad), z_vals.device)
        net = handle.struct(precision)
        check(_L().scade_mlp_forward_rays(byref(net), precision, ptr(rays), rays.shape[1], ptr(z_vals), N, S,
                                          _lib.host_floats(bb_center), float(bb_scale), ptr(raw), ptr(ws), ws.numel(),
                                          int(need_grad), stream_ptr()), "scade_mlp_forward_rays")
        ctx.handle, ctx.precision, ctx.P = handle, precision, N * S
        ctx.ws = ws if need_grad else None
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        if ctx.ws is None:
            raise _lib.ScadeError("backward through a forward that did not stash activations")
        grads = _mlp_backward(ctx.handle, ctx.precision, f32(d_raw), ctx.P, ctx.ws, d_raw.device)
        ctx.ws = None
        return (None, None, None, None, None, None, None, *grads)


class _MLPEmbeddedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, handle, precision, need_grad, *params):
        P = x.shape[0]
        out = torch.empty((P, 4), dtype=torch.float32, device=x.device)
        ws = _bytes(handle.workspace_bytes(P, precision, need_grad), x.device)
        net = handle.struct(precision)
        check(_L().scade_mlp_forward_embedded(byref(net), precision, ptr(x), P, ptr(out), ptr(ws), ws.numel(),
                                              int(need_grad), stream_ptr()), "scade_mlp_forward_embedded")
        ctx.handle, ctx.precision, ctx.P = handle, precision, P
        ctx.ws = ws if need_grad else None
        return out

    @staticmethod
    def backward(ctx, d_out):
        if ctx.ws is None:
            raise _lib.ScadeError("backward through a forward that did not stash activations")
        grads = _mlp_backward(ctx.handle, ctx.precision, f32(d_out), ctx.P, ctx.ws, d_out.device)
        ctx.ws = None
        return (None, None, None, None, *grads)


def _train_precision(handle, precision, need_grad):
    """Arithmetic of a forward that must be differentiable: the tight mode (tc_f16x3) is forward-only, so gradients at an fp32
    tolerance come from the fp32 FFMA path (any D / W / multires); tc_f16 trains on the tensor cores when the shape allows."""
    if not need_grad:
        return precision
    if precision == PREC_TC_F16X3:
        return PREC_FP32
    if precision == PREC_TC_F16 and not (tc_train_enabled() and handle.tc_supported()):
        return PREC_FP32
    return precision


def mlp_forward_rays(handle, rays, z_vals, bb_center, bb_scale, precision=PREC_FP32):
    """run_network fused with pts = o + d*z (RS:48-63, 657): rays [N,>=11], z [N,S] -> raw [N,S,4]."""
    precision = PRECISIONS[precision]
    rays, z_vals = f32(rays), f32(z_vals)
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in handle.params)
    precision = _train_precision(handle, precision, need_grad)
    if rays.shape[0] == 0:
        return torch.empty((0, z_vals.shape[1], 4), dtype=torch.float32, device=z_vals.device)
    return _MLPRaysFn.apply(rays, z_vals, handle, precision, [float(c) for c in bb_center], float(bb_scale),
                            need_grad, *handle.params)


def composite_fusable(handle, precision, S):
    """True when scade_mlp_forward_rays_composite handles this (network, precision, samples-per-ray) combination."""
    return bool(_L().scade_mlp_forward_rays_composite_supported(byref(handle.desc), PRECISIONS[precision], int(S)))


def composite_buffers(handle, N, S, precision=PREC_TC_F16, retraw=False, device=None):
    """Outputs + workspace of one mlp_forward_rays_composite call, as ONE allocation each (pass as `out=` to reuse them)."""
    precision = PRECISIONS[precision]
    n_raw = N * S * 4 if retraw else 0
    flat = torch.empty(n_raw + N * S + 6 * N, dtype=torch.float32, device=device)      # raw first: it needs 16-byte alignment
    o = n_raw
    buf = {"raw": flat[:n_raw].view(N, S, 4) if retraw else None}
    for key, shape in (("weights", (N, S)), ("rgb", (N, 3)), ("disp", (N,)), ("acc", (N,)), ("depth", (N,))):
        n = math.prod(shape)
        buf[key] = flat[o:o + n].view(shape)
        o += n
    buf["ws"] = torch.empty(handle.workspace_bytes(N * S, precision, 0), dtype=torch.uint8, device=device)
    return buf


def mlp_forward_rays_composite(handle, rays, z_vals, bb_center, bb_scale, precision=PREC_TC_F16, retraw=False, out=None):
    """run_network + raw2outputs (RS:659-660 / RS:718-720) as one kernel (no autograd): the alpha compositing runs on a
    compositor warp of the tensor-core kernel, fed by the epilogue of its last layer.  Returns (rgb_map, disp_map, acc_map,
    weights, depth_map, raw or None); `out` = composite_buffers(...) of the same shape writes into existing buffers."""
    precision = PRECISIONS[precision]
    rays, z_vals = f32(rays), f32(z_vals)
    N, S = z_vals.shape
    b = out if out is not None else composite_buffers(handle, N, S, precision, retraw, z_vals.device)
    if b["weights"].shape != (N, S) or (retraw and b["raw"] is None):
        raise ValueError("mlp_forward_rays_composite: `out` was made for another shape")
    raw = b["raw"] if retraw else None
    net = handle.struct(precision)
    ws = b["ws"]
    check(_L().scade_mlp_forward_rays_composite(byref(net), precision, ptr(rays), rays.shape[1], ptr(z_vals), N, S,
                                                _lib.host_floats([float(c) for c in bb_center]), float(bb_scale), ptr(raw),
                                                ptr(b["weights"]), ptr(b["rgb"]), ptr(b["disp"]), ptr(b["acc"]), ptr(b["depth"]),
                                                ptr(ws), ws.numel(), stream_ptr()),
          "scade_mlp_forward_rays_composite")
    return b["rgb"], b["disp"], b["acc"], b["weights"], b["depth"], raw


def mlp_forward_embedded(handle, x, precision=PREC_FP32):
    """NeRF.forward (H:223-247) on embedded inputs [..., in_ch + in_views] -> [..., 4]."""
    precision = PRECISIONS[precision]
    lead = x.shape[:-1]
    x2 = f32(x).reshape(-1, x.shape[-1])
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in handle.params)
    precision = _train_precision(handle, precision, need_grad)
    if x2.shape[0] == 0:
        return torch.empty((*lead, 4), dtype=torch.float32, device=x2.device)
    return _MLPEmbeddedFn.apply(x2, handle, precision, need_grad, *handle.params).reshape(*lead, 4)


def embed(x, multires):
    """Embedder.embed (H:171-172)."""
    lead = x.shape[:-1]
    x2 = f32(x).reshape(-1, 3)
    out = torch.empty((x2.shape[0], 3 + 6 * multires), dtype=torch.float32, device=x2.device)
    check(_L().scade_embed(ptr(x2), x2.shape[0], multires, ptr(out), stream_ptr()), "scade_embed")
    return out.reshape(*lead, -1)


# ----------------------------------------------------------------------------------------------
# rays
# ----------------------------------------------------------------------------------------------
def get_rays(H, W, intrinsic, c2w, col0=0, ncols=None, device=None):
    """get_rays (H:285-305) -> rays_o, rays_d [H, ncols, 3]."""
    ncols = W if ncols is None else ncols
    device = device or (c2w.device if torch.is_tensor(c2w) and c2w.is_cuda else torch.device("cuda"))
    intr = _lib.host_floats([float(v) for v in intrinsic][:4])
    pose = _lib.host_floats([float(v) for v in torch.as_tensor(c2w).detach().cpu().reshape(-1)[:12]])
    rays_o = torch.empty((H, ncols, 3), dtype=torch.float32, device=device)
    rays_d = torch.empty_like(rays_o)
    check(_L().scade_get_rays(H, W, intr, pose, col0, ncols, ptr(rays_o), ptr(rays_d), stream_ptr()), "scade_get_rays")
    return rays_o, rays_d


def make_ray_batch(rays_o, rays_d, near, far):
    """render()'s batch assembly (RS:123-141): [N,11] = (o, d, near, far, d/|d|)."""
    rays_o, rays_d = f32(rays_o).reshape(-1, 3), f32(rays_d).reshape(-1, 3)
    out = torch.empty((rays_o.shape[0], 11), dtype=torch.float32, device=rays_o.device)
    check(_L().scade_make_ray_batch(ptr(rays_o), ptr(rays_d), rays_o.shape[0], float(near), float(far), ptr(out),
                                    stream_ptr()), "scade_make_ray_batch")
    return out


def camera_ray_batch(H, W, intrinsic, c2w, near, far, pix0=0, n=None, col0=0, ncols=None, device=None):
    ncols = W if ncols is None else ncols
    n = H * ncols - pix0 if n is None else n
    device = device or torch.device("cuda")
    intr = _lib.host_floats([float(v) for v in intrinsic][:4])
    pose = _lib.host_floats([float(v) for v in torch.as_tensor(c2w).detach().cpu().reshape(-1)[:12]])
    out = torch.empty((n, 11), dtype=torch.float32, device=device)
    check(_L().scade_camera_ray_batch(H, W, intr, pose, col0, ncols, pix0, n, float(near), float(far), ptr(out),
                                      stream_ptr()), "scade_camera_ray_batch")
    return out


def coarse_z_vals(ray_batch, n_samples, lindisp=False, t_rand=None):
    ray_batch = f32(ray_batch)
    N = ray_batch.shape[0]
    z = torch.empty((N, n_samples), dtype=torch.float32, device=ray_batch.device)
    t_rand = None if t_rand is None else f32(t_rand)
    check(_L().scade_coarse_z_vals(ptr(ray_batch), ray_batch.shape[1], N, n_samples, int(bool(lindisp)), ptr(t_rand),
                                   ptr(z), stream_ptr()), "scade_coarse_z_vals")
    return z


def perturb_z_vals(z_vals, t_rand):
    z_vals, t_rand = f32(z_vals), f32(t_rand)
    out = torch.empty_like(z_vals)
    check(_L().scade_perturb_z_vals(ptr(z_vals), ptr(t_rand), z_vals.shape[0], z_vals.shape[1], ptr(out), stream_ptr()),
          "scade_perturb_z_vals")
    return out


# ----------------------------------------------------------------------------------------------
# compositing
# ----------------------------------------------------------------------------------------------
class _Raw2OutputsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z_vals, rays_d, noise):
        N, S = z_vals.shape
        dev = raw.device
        rgb = torch.empty((N, 3), dtype=torch.float32, device=dev)
        disp = torch.empty((N,), dtype=torch.float32, device=dev)
        acc = torch.empty_like(disp)
        depth = torch.empty_like(disp)
        w = torch.empty((N, S), dtype=torch.float32, device=dev)
        check(_L().scade_raw2outputs(ptr(raw), ptr(z_vals), _row_ptr(rays_d), rays_d.stride(0), ptr(noise), N, S, ptr(rgb),
                                     ptr(disp), ptr(acc), ptr(w), ptr(depth), stream_ptr()), "scade_raw2outputs")
        ctx.save_for_backward(raw, z_vals, rays_d, noise if noise is not None else torch.empty(0, device=dev))
        ctx.has_noise = noise is not None
        ctx.set_materialize_grads(False)      # unused outputs (disp / acc / depth in the train step) arrive as None, not as zero tensors
        return rgb, disp, acc, w, depth

    @staticmethod
    def backward(ctx, d_rgb, d_disp, d_acc, d_w, d_depth):
        raw, z_vals, rays_d, noise = ctx.saved_tensors
        N, S = z_vals.shape
        if all(t is None for t in (d_rgb, d_disp, d_acc, d_w, d_depth)):
            return None, None, None, None
        d_raw = torch.empty_like(raw)
        g = [None if t is None else f32(t) for t in (d_rgb, d_disp, d_acc, d_w, d_depth)]
        check(_L().scade_raw2outputs_backward(ptr(raw), ptr(z_vals), _row_ptr(rays_d), rays_d.stride(0),
                                              ptr(noise) if ctx.has_noise else None, N, S, ptr(g[0]), ptr(g[1]),
                                              ptr(g[2]), ptr(g[3]), ptr(g[4]), ptr(d_raw), stream_ptr()),
              "scade_raw2outputs_backward")
        return d_raw, None, None, None


def _row_ptr(t):
    """Device pointer of a 2-D fp32 CUDA tensor whose rows are contiguous but may be strided (a column slice such as
    ray_batch[:, 3:6]): the kernels take the row stride separately."""
    if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
        raise _lib.ScadeError("expected a 2-D fp32 CUDA tensor with unit column stride")
    return c_void_p(t.data_ptr())


def raw2outputs(raw, z_vals, rays_d, noise=None):
    """compute_weights + raw2outputs (RS:511-562) -> (rgb_map, disp_map, acc_map, weights, depth_map).  rays_d [N,3] may be a
    column slice of the ray batch (no copy is made)."""
    if not torch.is_tensor(rays_d):
        rays_d = torch.as_tensor(rays_d)
    if rays_d.dtype != torch.float32 or rays_d.dim() != 2 or rays_d.stride(1) != 1:
        rays_d = f32(rays_d).reshape(-1, rays_d.shape[-1])
    return _Raw2OutputsFn.apply(f32(raw), f32(z_vals), rays_d, None if noise is None else f32(noise))


# ----------------------------------------------------------------------------------------------
# hierarchical sampling
# ----------------------------------------------------------------------------------------------
class _SamplePdfFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bins, weights, u, n_samples, joint):
        N, B = bins.shape
        samples = torch.empty((N, n_samples), dtype=torch.float32, device=bins.device)
        u_out = torch.empty_like(samples)
        check(_L().scade_sample_pdf(ptr(bins), ptr(weights), N, B, n_samples, ptr(u), int(joint), ptr(samples),
                                    ptr(u_out), stream_ptr()), "scade_sample_pdf")
        ctx.save_for_backward(bins, weights, u_out)
        ctx.mark_non_differentiable(u_out)
        ctx.set_materialize_grads(False)
        return samples, u_out

    @staticmethod
    def backward(ctx, d_samples, _d_u):
        if d_samples is None:
            return None, None, None, None, None
        bins, weights, u = ctx.saved_tensors
        N, B = bins.shape
        d_w = torch.empty_like(weights)
        check(_L().scade_sample_pdf_backward(ptr(bins), ptr(weights), ptr(u), N, B, u.shape[1], ptr(f32(d_samples)),
                                             ptr(d_w), stream_ptr()), "scade_sample_pdf_backward")
        return None, d_w, None, None, None


def sample_pdf(bins, weights, n_samples, u=None, joint=False):
    """sample_pdf family (H:337-538).  u=None -> det (linspace); returns (samples, u_used)."""
    bins, weights = f32(bins), f32(weights)
    lead = bins.shape[:-1]
    bins2, w2 = bins.reshape(-1, bins.shape[-1]), weights.reshape(-1, weights.shape[-1])
    if u is not None:
        u = f32(u, bins.device)
        if not joint:
            u = u.reshape(-1, n_samples)
    s, uo = _SamplePdfFn.apply(bins2, w2, u, n_samples, bool(joint))
    return s.reshape(*lead, n_samples), uo.reshape(*lead, n_samples)


class _ResampleFromZFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_vals, weights, u, n_samples, joint, want_merge, want_std):
        N, S = z_vals.shape
        dev = z_vals.device
        samples = torch.empty((N, n_samples), dtype=torch.float32, device=dev)
        u_out = torch.empty_like(samples)
        merged = torch.empty((N, S + n_samples), dtype=torch.float32, device=dev) if want_merge else None
        std = torch.empty((N,), dtype=torch.float32, device=dev) if want_std else None
        check(_L().scade_resample_from_z(ptr(z_vals), ptr(weights), N, S, n_samples, ptr(u), int(joint), ptr(samples),
                                         ptr(u_out), ptr(merged), ptr(std), stream_ptr()), "scade_resample_from_z")
        ctx.save_for_backward(z_vals, weights, u_out)
        outs = (samples, u_out, merged if want_merge else torch.empty(0, device=dev),
                std if want_std else torch.empty(0, device=dev))
        ctx.mark_non_differentiable(*outs[1:])
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, d_samples, *_):
        if d_samples is None:
            return None, None, None, None, None, None, None
        z_vals, weights, u = ctx.saved_tensors
        N, S = z_vals.shape
        d_w = torch.empty_like(weights)
        check(_L().scade_resample_from_z_backward(ptr(z_vals), ptr(weights), ptr(u), N, S, u.shape[1],
                                                  ptr(f32(d_samples)), ptr(d_w), 0, stream_ptr()),
              "scade_resample_from_z_backward")
        return None, d_w, None, None, None, None, None


def resample_from_z(z_vals, weights, n_samples, u=None, joint=False, merge=False, std=False):
    """RS:702-713 / RS:723-726: mid-point bins + weights[:,1:-1] formed on the fly.
    Returns (samples, u_used, z_merged or None, z_std or None)."""
    z_vals, weights = f32(z_vals), f32(weights)
    if u is not None:
        u = f32(u, z_vals.device)
    s, uo, m, sd = _ResampleFromZFn.apply(z_vals, weights, u, n_samples, bool(joint), bool(merge), bool(std))
    return s, uo, (m if merge else None), (sd if std else None)


def sort_merge(a, b):
    """torch.sort(torch.cat([a, b], -1), -1).values (RS:713)."""
    a, b = f32(a), f32(b)
    out = torch.empty((a.shape[0], a.shape[1] + b.shape[1]), dtype=torch.float32, device=a.device)
    check(_L().scade_sort_merge(ptr(a), a.shape[1], ptr(b), b.shape[1], a.shape[0], ptr(out), stream_ptr()),
          "scade_sort_merge")
    return out


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------
class _SpaceCarvingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, hyp, mask, is_joint, threshold, want):
        N, P = pred.shape
        K = hyp.shape[0]
        full = hyp.shape[-1] != 1
        dev = pred.device
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        d_pred = torch.empty_like(pred) if want else None
        d_hyp = torch.empty_like(hyp) if want else None
        ws = _bytes(_L().scade_space_carving_workspace_bytes(K, N, P), dev)
        check(_L().scade_space_carving_loss(ptr(pred), ptr(hyp), int(full), ptr(mask), K, N, P, int(is_joint),
                                            float(threshold), 1.0, ptr(loss), ptr(d_pred), ptr(d_hyp), ptr(ws),
                                            ws.numel(), stream_ptr()), "scade_space_carving_loss")
        if want:
            ctx.save_for_backward(d_pred, d_hyp)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        d_pred, d_hyp = ctx.saved_tensors
        return d_pred * g, d_hyp * g, None, None, None, None


def space_carving_loss(pred, hyp, is_joint=False, mask=None, threshold=0.0):
    """compute_space_carving_loss (H:93-128)."""
    pred, hyp = f32(pred), f32(hyp)
    mask = None if mask is None else f32(mask)
    want = torch.is_grad_enabled() and (pred.requires_grad or hyp.requires_grad)
    return _SpaceCarvingFn.apply(pred, hyp, mask, bool(is_joint), float(threshold), want)


class _SpaceCarvingAffineFn(torch.autograd.Function):
    """compute_space_carving_loss(pred, target_h * scale + shift) (RS:954 + H:93-128, default branch) in one launch, with the
    gradients w.r.t. pred, scale and shift produced in the same pass."""

    @staticmethod
    def forward(ctx, pred, hyp_raw, scale, shift, mask, threshold, denominator, want):
        N, P = pred.shape
        K = hyp_raw.shape[0]
        dev = pred.device
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        d_pred = torch.empty_like(pred) if want else None
        d_ss = torch.empty((2,), dtype=torch.float32, device=dev) if want else None
        check(_L().scade_space_carving_loss_affine(ptr(pred), ptr(hyp_raw), ptr(scale), ptr(shift), ptr(mask), K, N, P,
                                                   float(threshold), 1.0, int(denominator), ptr(loss), ptr(d_pred),
                                                   ptr(d_ss), ptr(d_ss[1:]) if want else None, 0, stream_ptr()),
              "scade_space_carving_loss_affine")
        if want:
            ctx.save_for_backward(d_pred, d_ss)
        ctx.shapes = (scale.shape, shift.shape)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        d_pred, d_ss = ctx.saved_tensors
        ss = d_ss * g
        return d_pred * g, None, ss[0].reshape(ctx.shapes[0]), ss[1].reshape(ctx.shapes[1]), None, None, None, None


def space_carving_loss_affine(pred, hyp_raw, scale, shift, mask=None, threshold=0.0, denominator=0):
    """compute_space_carving_loss(pred, hyp_raw * scale + shift) (RS:954, RS:974; is_joint=False, hyp_raw [K,N,1], scale / shift
    single-element CUDA tensors).  denominator > 0 replaces N in the mean over rays (ray-sharded training)."""
    pred, hyp_raw = f32(pred), f32(hyp_raw)
    if hyp_raw.dim() != 3 or hyp_raw.shape[-1] != 1 or scale.numel() != 1 or shift.numel() != 1:
        raise _lib.ScadeError("space_carving_loss_affine: needs hyp_raw [K,N,1] and single-element scale / shift")
    mask = None if mask is None else f32(mask)
    want = torch.is_grad_enabled() and (pred.requires_grad or scale.requires_grad or shift.requires_grad)
    return _SpaceCarvingAffineFn.apply(pred, hyp_raw, f32(scale), f32(shift), mask, float(threshold), int(denominator), want)


class _SpaceCarvingJointShardedFn(torch.autograd.Function):
    """Joint branch (H:115-119) on a ray shard: local [K,P] distance sums -> ONE all-reduce of K*P floats -> arg-min over k of
    the GLOBAL means -> loss (global, identical on every rank) and this shard's gradients (SURVEY 8(e) "Exception")."""

    @staticmethod
    def forward(ctx, pred, hyp, mask, threshold, n_global, group, want):
        import torch.distributed as dist
        N, P = pred.shape
        K = hyp.shape[0]
        full = hyp.shape[-1] != 1
        dev = pred.device
        qsum = torch.empty((K, P), dtype=torch.float32, device=dev)
        check(_L().scade_space_carving_joint_accumulate(ptr(pred), ptr(hyp), int(full), ptr(mask), K, N, P, float(threshold),
                                                        ptr(qsum), stream_ptr()), "scade_space_carving_joint_accumulate")
        from .dist import _world
        if _world(group)[1] > 1:
            dist.all_reduce(qsum, op=dist.ReduceOp.SUM, group=group)
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        d_pred = torch.empty_like(pred) if want else None
        d_hyp = torch.empty_like(hyp) if want else None
        kstar = torch.empty((P,), dtype=torch.int32, device=dev)
        check(_L().scade_space_carving_joint_finish(ptr(pred), ptr(hyp), int(full), ptr(mask), ptr(qsum), K, N, int(n_global), P,
                                                    float(threshold), 1.0, ptr(loss), ptr(d_pred), ptr(d_hyp), ptr(kstar),
                                                    stream_ptr()), "scade_space_carving_joint_finish")
        if want:
            ctx.save_for_backward(d_pred, d_hyp)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        d_pred, d_hyp = ctx.saved_tensors
        return d_pred * g, d_hyp * g, None, None, None, None, None


def space_carving_loss_joint_sharded(pred, hyp, n_global, mask=None, threshold=0.0, group=None):
    """compute_space_carving_loss(is_joint=True) (H:115-119) when `pred` / `hyp` hold this rank's rays of a step whose rays are
    split across the ranks of `group`: returns the GLOBAL loss; backward yields this shard's part of its gradient."""
    pred, hyp = f32(pred), f32(hyp)
    mask = None if mask is None else f32(mask)
    want = torch.is_grad_enabled() and (pred.requires_grad or hyp.requires_grad)
    return _SpaceCarvingJointShardedFn.apply(pred, hyp, mask, float(threshold), int(n_global), group, want)


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, denominator, want):
        loss = torch.empty((1,), dtype=torch.float32, device=x.device)
        d_x = torch.empty_like(x) if want else None
        check(_L().scade_img2mse(ptr(x), ptr(y), x.numel(), int(denominator), 1.0, ptr(loss), ptr(d_x), stream_ptr()),
              "scade_img2mse")
        if want:
            ctx.save_for_backward(d_x)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (d_x,) = ctx.saved_tensors
        return d_x * g, None, None, None


def img2mse_head(x, y, denominator, grad_scale, loss_out):
    """img2mse (H:11) as a loss HEAD without autograd nodes: writes the loss into `loss_out` (1 float, e.g. a slot of the flat
    gradient buffer's tail) and returns grad_scale * d loss / d x for torch.autograd.backward(x, grad)."""
    x = f32(x)
    d_x = torch.empty_like(x)
    check(_L().scade_img2mse(ptr(x), ptr(f32(y).expand_as(x).contiguous()), x.numel(), int(denominator), float(grad_scale),
                             ptr(loss_out), ptr(d_x), stream_ptr()), "scade_img2mse")
    return d_x


def space_carving_affine_head(pred, hyp_raw, scale, shift, mask, threshold, denominator, grad_scale, loss_out, d_scale, d_shift,
                              accumulate):
    """scade_space_carving_loss_affine as a loss head without autograd nodes: the (unweighted) loss goes to `loss_out`,
    grad_scale * d loss / d pred is returned, grad_scale * d loss / d scale|shift are written (or accumulated) into the
    1-element tensors d_scale / d_shift."""
    pred, hyp_raw = f32(pred), f32(hyp_raw)
    N, P = pred.shape
    d_pred = torch.empty_like(pred)
    check(_L().scade_space_carving_loss_affine(ptr(pred), ptr(hyp_raw), ptr(f32(scale)), ptr(f32(shift)),
                                               ptr(None if mask is None else f32(mask)), hyp_raw.shape[0], N, P, float(threshold),
                                               float(grad_scale), int(denominator), ptr(loss_out), ptr(d_pred), ptr(d_scale),
                                               ptr(d_shift), int(bool(accumulate)), stream_ptr()),
          "scade_space_carving_loss_affine")
    return d_pred


def img2mse(x, y, denominator=0):
    """img2mse (H:11).  denominator > 0 replaces numel() in the mean (ray-sharded training)."""
    x = f32(x)
    return _MseFn.apply(x, f32(y).expand_as(x).contiguous(), denominator, torch.is_grad_enabled() and x.requires_grad)


# ----------------------------------------------------------------------------------------------
# fused render_rays forward (no autograd): one C call, one stream
# ----------------------------------------------------------------------------------------------
def render_rays_forward(ray_batch, coarse, fine, n_samples, n_importance, bb_center, bb_scale, precision=PREC_FP32,
                        lindisp=False, is_joint=False, t_rand=None, u_coarse=None, u_fine=None, retraw=False, out=None):
    """scade_render_rays_forward.  `out`: optional {name: preallocated contiguous fp32 CUDA tensor} for any of the returned
    tensors (GraphedRenderRays lays the maps it reads back in ONE buffer so that they leave in one device->host copy)."""
    precision = PRECISIONS[precision]
    ray_batch = f32(ray_batch)
    N = ray_batch.shape[0]
    dev = ray_batch.device
    fine = fine or coarse
    cfg = _lib.RenderCfg(n_samples, n_importance, int(bool(lindisp)), precision, int(bool(is_joint)), ray_batch.shape[1],
                         (ctypes.c_float * 3)(*[float(c) for c in bb_center]), float(bb_scale))
    S = n_samples + n_importance
    shapes = {"rgb_map": (N, 3), "disp_map": (N,), "acc_map": (N,), "depth_map": (N,), "z_vals": (N, S),
              "weights": (N, S), "pred_hyp": (N, n_importance), "u": (N, n_importance), "rgb0": (N, 3), "disp0": (N,),
              "acc0": (N,), "depth0": (N,), "z_vals0": (N, n_samples), "weights0": (N, n_samples), "z_std": (N,)}
    if retraw:
        shapes["raw"] = (N, S, 4)
    ret = {}
    need = []
    for k, shp in shapes.items():
        t = None if out is None else out.get(k)
        if t is not None and (tuple(t.shape) != tuple(shp) or t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous()):
            raise _lib.ScadeError(f"render_rays_forward: out[{k!r}] must be a contiguous fp32 CUDA tensor of shape {tuple(shp)}")
        ret[k] = t
        if t is None:
            need.append(k)
    if need:
        # ONE allocation for all remaining outputs (the reference's dict has 15 entries: 15 allocator calls per chunk otherwise),
        # every view 16-byte aligned
        numel = {k: int(torch.Size(shapes[k]).numel()) for k in need}
        flat = torch.empty(sum((n + 3) // 4 * 4 for n in numel.values()), dtype=torch.float32, device=dev)
        off = 0
        for k in need:
            ret[k] = flat[off:off + numel[k]].view(shapes[k])
            off += (numel[k] + 3) // 4 * 4
    out_c = _lib.RenderOut()
    for k in _lib.RENDER_OUT_FIELDS:
        setattr(out_c, k, ret[k].data_ptr() if k in ret else None)
    nc, nf = coarse.struct(precision), fine.struct(precision)
    ws = _bytes(_L().scade_render_rays_workspace_bytes(byref(cfg), byref(coarse.desc), byref(fine.desc), N), dev)
    t_rand = None if t_rand is None else f32(t_rand, dev)
    u_coarse = None if u_coarse is None else f32(u_coarse, dev)
    u_fine = None if u_fine is None else f32(u_fine, dev)
    check(_L().scade_render_rays_forward(byref(cfg), ptr(ray_batch), N, byref(nc), byref(nf), ptr(t_rand), ptr(u_coarse),
                                         ptr(u_fine), byref(out_c), ptr(ws), ws.numel(), stream_ptr()),
          "scade_render_rays_forward")
    return ret
