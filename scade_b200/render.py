"""Host-side mirror of the renderer half of the reference's ``run_scade_scannet.py`` (RS:39-233, 422-751).

Same callables, argument names, kwargs plumbing and returned dict keys as the reference, so its
train / test / video loops run unchanged on top of these.  All arithmetic happens in
libscade_b200.so; this module only sequences kernels, draws the random numbers the reference draws with
torch, and keeps the autograd tape.

    batchify, run_network            RS:39-63
    batchify_rays                    RS:66-78
    render, render_hyp               RS:80-233
    create_nerf                      RS:422-509
    compute_weights, raw2outputs     RS:511-562
    perturb_z_vals                   RS:564-579
    render_rays                      RS:581-751
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import functional as F_
from .nerf_helpers import NeRF, get_embedder


def _unwrap(net):
    """The reference wraps its nets in nn.DataParallel (RS:438,455); rays shard across processes here."""
    return net.module if isinstance(net, nn.DataParallel) else net


def batchify(fn, chunk):
    """RS:39-46"""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network(inputs, viewdirs, embedded_cam, fn, embed_fn, embeddirs_fn, bb_center, bb_scale, netchunk=1024 * 64):
    """RS:48-63 with explicit points (compatibility form: materialises the embedded [P,60] matrix like the
    reference does; render_rays never takes this route when given a NetworkQuery)."""
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    inputs_flat = (inputs_flat - torch.as_tensor(bb_center, device=inputs.device, dtype=inputs.dtype)) * bb_scale
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        embedded_dirs = embeddirs_fn(torch.reshape(input_dirs, [-1, input_dirs.shape[-1]]))
        embedded = torch.cat([embedded, embedded_dirs], -1)
    outputs_flat = batchify(_unwrap(fn), netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


class NetworkQuery:
    """What create_nerf binds as ``network_query_fn`` (RS:461-466).  Calling it with points reproduces
    run_network; render_rays recognises the object and instead fuses point generation, normalisation,
    encoding and the MLP into one kernel launch per pass."""

    def __init__(self, embed_fn, embeddirs_fn, bb_center, bb_scale, netchunk=1024 * 64, precision="fp32"):
        self.embed_fn, self.embeddirs_fn = embed_fn, embeddirs_fn
        self.bb_center = [float(c) for c in torch.as_tensor(bb_center).reshape(-1).tolist()]
        self.bb_scale = float(bb_scale)
        self.netchunk = netchunk
        self.precision = precision

    def __call__(self, inputs, viewdirs, embedded_cam, network_fn):
        return run_network(inputs, viewdirs, embedded_cam, network_fn, self.embed_fn, self.embeddirs_fn,
                           self.bb_center, self.bb_scale, self.netchunk)

    def query_rays(self, ray_batch, z_vals, network_fn):
        return F_.mlp_forward_rays(_unwrap(network_fn).handle(), ray_batch, z_vals, self.bb_center, self.bb_scale,
                                   self.precision)


def compute_weights(raw, z_vals, rays_d, noise=0.):
    """RS:511-522"""
    nz = noise if torch.is_tensor(noise) else None
    return F_.raw2outputs(raw, z_vals, rays_d, nz)[3]


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, pytest=False):
    """RS:530-562 -> (rgb_map, disp_map, acc_map, weights, depth_map)"""
    noise = None
    if raw_noise_std > 0.:
        if pytest:                                                        # RS:548-552
            np.random.seed(0)
            noise = torch.tensor(np.random.rand(*list(raw[..., 3].shape)) * raw_noise_std, dtype=torch.float32,
                                 device=raw.device)
        else:
            noise = torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std   # RS:546
    return F_.raw2outputs(raw, z_vals, rays_d, noise)


def raw2depth(raw, z_vals, rays_d):
    """RS:524-528"""
    _, _, _, weights, depth = F_.raw2outputs(raw, z_vals, rays_d)
    std = (((z_vals - depth.unsqueeze(-1)).pow(2) * weights).sum(-1)).sqrt()
    return depth, std


def perturb_z_vals(z_vals, pytest):
    """RS:564-579"""
    if pytest:
        np.random.seed(0)
        t_rand = torch.tensor(np.random.rand(*list(z_vals.shape)), dtype=torch.float32, device=z_vals.device)
    else:
        t_rand = torch.rand_like(z_vals)
    return F_.perturb_z_vals(z_vals, t_rand)


def _rand(shape, pytest, device):
    if pytest:
        np.random.seed(0)
        return torch.tensor(np.random.rand(*shape), dtype=torch.float32, device=device)
    return torch.rand(shape, device=device)


def render_rays(ray_batch, use_viewdirs, network_fn, network_query_fn, N_samples, precomputed_z_samples=None,
                embedded_cam=None, retraw=False, lindisp=False, perturb=0., N_importance=0, network_fine=None,
                raw_noise_std=0., verbose=False, pytest=False, is_joint=False, cached_u=None, near=None, far=None,
                ndc=None, t_rand=None, u_coarse=None, _out=None):
    """RS:581-751.  Returns the reference's dict (RS:733-744); every value is a tensor.

    Extra keyword arguments ``t_rand`` / ``u_coarse`` inject the uniforms the reference draws at RS:570 and
    H:350 (``cached_u`` already injects the third draw, RS:726); near/far/ndc are swallowed because the
    reference's kwargs dict carries them (RS:500-502).
    """
    if not use_viewdirs:
        raise NotImplementedError("render_rays: the SCADE configuration is use_viewdirs=True (RS:1141)")
    if N_importance <= 0:
        raise NotImplementedError("render_rays: N_importance == 0 is a dead branch in the reference (RS:733 NameError)")
    ray_batch = F_.f32(ray_batch)
    if not ray_batch.is_cuda:
        raise F_._lib.ScadeError("render_rays needs CUDA tensors; scade_b200 has no CPU path")
    N = ray_batch.shape[0]
    dev = ray_batch.device
    fine = network_fine if network_fine is not None else network_fn                 # RS:716
    det = not (perturb > 0.)
    if not det:
        t_rand = _rand((N, N_samples), pytest, dev) if t_rand is None else t_rand    # RS:570
        if u_coarse is None:                                                        # H:350 / H:452
            u_coarse = _rand((N_importance,), pytest, dev).expand(N, N_importance) if is_joint and not pytest \
                else _rand((N, N_importance), pytest, dev)
        u_fine = cached_u                                                           # RS:726
        if u_fine is None:
            u_fine = _rand((N_importance,), pytest, dev).expand(N, N_importance) if is_joint and not pytest \
                else _rand((N, N_importance), pytest, dev)
        u_coarse, u_fine = F_.f32(u_coarse, dev), F_.f32(u_fine, dev)
    else:
        t_rand, u_coarse = None, None
        u_fine = None if cached_u is None else F_.f32(cached_u, dev)

    fused = isinstance(network_query_fn, NetworkQuery)
    nets = [_unwrap(network_fn), _unwrap(fine)]
    needs_grad = torch.is_grad_enabled() and any(p.requires_grad for n in nets for p in n.parameters())

    if fused and not needs_grad and raw_noise_std == 0. and (det == (u_fine is None)):
        # eval / benchmark path: the whole function is one C call on one stream
        return F_.render_rays_forward(ray_batch, nets[0].handle(), nets[1].handle(), N_samples, N_importance,
                                      network_query_fn.bb_center, network_query_fn.bb_scale,
                                      precision=network_query_fn.precision, lindisp=lindisp, is_joint=False,
                                      t_rand=t_rand, u_coarse=u_coarse, u_fine=u_fine, retraw=retraw, out=_out)

    rays_o, rays_d, viewdirs = ray_batch[:, 0:3], ray_batch[:, 3:6], ray_batch[:, 8:11]     # (views: the kernels take a row stride)

    def query(z, net):
        if fused:
            return network_query_fn.query_rays(ray_batch, z, net)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]         # RS:657
        return network_query_fn(pts, viewdirs, embedded_cam, net)

    z_vals = F_.coarse_z_vals(ray_batch, N_samples, lindisp, t_rand)                # RS:640-655
    raw = query(z_vals, network_fn)                                                 # RS:659
    rgb0, disp0, acc0, weights0, depth0 = raw2outputs(raw, z_vals, rays_d, raw_noise_std, pytest=pytest)   # RS:660
    z_vals0 = z_vals
    _, _, z_vals, _ = F_.resample_from_z(z_vals0, weights0.detach(), N_importance, u=u_coarse, merge=True)  # RS:702-713
    raw = query(z_vals, fine)                                                       # RS:718
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, raw_noise_std, pytest=pytest)
    pred_hyp, u, _, z_std = F_.resample_from_z(z_vals, weights, N_importance, u=u_fine, std=True)           # RS:723-730
    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map, 'depth_map': depth_map, 'z_vals': z_vals,
           'weights': weights, 'pred_hyp': pred_hyp, 'u': u}
    if retraw:
        ret['raw'] = raw
    ret.update(rgb0=rgb0, disp0=disp0, acc0=acc0, depth0=depth0, z_vals0=z_vals0, weights0=weights0, z_std=z_std)
    return ret


class GraphedRenderRays:
    """render_rays (RS:581-751, eval configuration: no autograd, deterministic sampling or injected uniforms) for a FIXED ray
    count, replayed as one CUDA graph: the 5 kernel launches of the pass, their output allocations and the host-side argument
    marshalling happen once at construction.  For loops that render many equal-sized chunks (RS:347, batchify_rays).

        g = GraphedRenderRays(n_rays, **render_kwargs_test)      # same kwargs as render_rays / create_nerf's dict
        out = g(ray_batch)                                        # ray_batch [n_rays, 11]: CUDA tensor or (pinned) host tensor

    `out` holds the graph's own output tensors (the reference's dict keys); the next call overwrites them.  The weights are read
    through the packed fp16 stream: call `g.refresh()` (re-capture) after an optimizer step.
    """

    def __init__(self, n_rays, use_viewdirs=True, device=None, host_outputs=None, **kwargs):
        """host_outputs: names of result tensors to be delivered in pinned host memory.  The graph then also contains the
        host->device copy of `self.rays_host` (pinned, [n_rays, 11]) and the device->host copies into `self.out_host[name]`:
        fill `rays_host`, call `g()`, synchronise the stream, read `out_host` (and synchronise before refilling `rays_host`: the
        copy inside the previous replay reads it asynchronously)."""
        self.n_rays, self.use_viewdirs, self.kwargs = int(n_rays), use_viewdirs, dict(kwargs)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.rays = torch.zeros((self.n_rays, 11), dtype=torch.float32, device=self.device)
        self.rays[:, 3:6] = 1.0
        self.rays[:, 6], self.rays[:, 7] = 0.1, 1.0
        self.host_outputs = tuple(host_outputs) if host_outputs else None
        self.rays_host = self.rays.cpu().pin_memory() if self.host_outputs else None
        self.out_host = None
        self.graph, self.out = None, None
        self.refresh()

    def refresh(self):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                               # warm-up outside the capture: packs weights, sets kernel attributes
                ret = render_rays(self.rays, self.use_viewdirs, **self.kwargs)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if self.host_outputs and self.out_host is None:
            # the maps that travel back live in ONE device buffer and ONE pinned host buffer: a single device->host copy per step
            sizes = [int(ret[k].numel()) for k in self.host_outputs]
            self._dev_flat = torch.empty(sum(sizes), dtype=torch.float32, device=self.device)
            self._host_flat = torch.empty(sum(sizes), dtype=torch.float32).pin_memory()
            self._dev_views, self.out_host, off = {}, {}, 0
            for k, n in zip(self.host_outputs, sizes):
                self._dev_views[k] = self._dev_flat[off:off + n].view(ret[k].shape)
                self.out_host[k] = self._host_flat[off:off + n].view(ret[k].shape)
                off += n
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: CUDA calls of other threads (e.g. an NCCL watchdog polling its events) must not invalidate the capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"), torch.no_grad():
            if self.host_outputs:
                self.rays.copy_(self.rays_host, non_blocking=True)
                self.out = render_rays(self.rays, self.use_viewdirs, _out=self._dev_views, **self.kwargs)
                for k, v in self._dev_views.items():          # (a path that allocated its own outputs: stage them)
                    if self.out[k].data_ptr() != v.data_ptr():
                        v.copy_(self.out[k])
                self._host_flat.copy_(self._dev_flat, non_blocking=True)
            else:
                self.out = render_rays(self.rays, self.use_viewdirs, **self.kwargs)

    def __call__(self, ray_batch=None):
        if ray_batch is not None:
            if tuple(ray_batch.shape) != (self.n_rays, 11):
                raise ValueError(f"GraphedRenderRays was captured for [{self.n_rays}, 11] ray batches, got {tuple(ray_batch.shape)}")
            if self.host_outputs:
                if ray_batch.is_cuda:
                    raise ValueError("this GraphedRenderRays takes its rays from pinned host memory (rays_host)")
                self.rays_host.copy_(ray_batch)
            else:
                self.rays.copy_(ray_batch, non_blocking=True)
        self.graph.replay()
        return self.out


class PipelinedRenderRays:
    """`depth` GraphedRenderRays slots, each on its own CUDA stream, used round-robin: while slot A's kernels run, slot B's ray
    batch crosses PCIe and slot A's previous maps travel back -- the copies of a serving loop overlap the compute instead of
    bracketing it.  Each slot owns its pinned host buffers, device buffers and graph.

        pipe = PipelinedRenderRays(n_rays, depth=2, host_outputs=("rgb_map", "depth_map"), **render_kwargs_test)
        t0 = pipe.submit(rays_host_0)                # host tensor [n_rays, 11]; returns a ticket
        t1 = pipe.submit(rays_host_1)
        maps0 = pipe.result(t0)                      # waits for that submission only; pinned host tensors, valid until the
                                                     # slot is submitted again (depth submissions later)
    """

    def __init__(self, n_rays, depth=2, host_outputs=("rgb_map", "disp_map", "acc_map", "depth_map"), device=None, **kwargs):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.depth = int(depth)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.depth)]
        self.slots = []
        for st in self.streams:
            st.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(st):
                self.slots.append(GraphedRenderRays(n_rays, device=self.device, host_outputs=host_outputs, **kwargs))
        torch.cuda.synchronize(self.device)
        self.events = [None] * self.depth
        self.count = 0

    def submit(self, ray_batch_host):
        i = self.count % self.depth
        if self.events[i] is not None:
            self.events[i].synchronize()             # the slot's previous run has read its rays_host and written its out_host
        slot = self.slots[i]
        slot.rays_host.copy_(ray_batch_host)
        with torch.cuda.stream(self.streams[i]):
            slot.graph.replay()
            ev = torch.cuda.Event()
            ev.record(self.streams[i])
        self.events[i] = ev
        self.count += 1
        return self.count - 1

    def result(self, ticket):
        if ticket < self.count - self.depth or ticket >= self.count:
            raise ValueError("PipelinedRenderRays.result: that submission's slot has been reused (or the ticket is unknown)")
        i = ticket % self.depth
        self.events[i].synchronize()
        return self.slots[i].out_host

    def drain(self):
        for ev in self.events:
            if ev is not None:
                ev.synchronize()


def batchify_rays(rays_flat, chunk=1024 * 32, use_viewdirs=False, **kwargs):
    """RS:66-78"""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], use_viewdirs, **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


def render(H, W, intrinsic, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., with_5_9=False,
           use_viewdirs=False, c2w_staticcam=None, rays_depth=None, **kwargs):
    """RS:80-155 -> [rgb_map, disp_map, acc_map, extras]"""
    if c2w_staticcam is not None or rays_depth is not None:
        raise NotImplementedError("c2w_staticcam / rays_depth are unused by the SCADE loops")
    if not use_viewdirs:
        raise NotImplementedError("the SCADE configuration is use_viewdirs=True (RS:1141)")
    if c2w is not None:
        col0, ncols = 0, W
        if with_5_9:                                                                # RS:109-116
            ncols = int(H / 9. * 16. / 3.)
            if ncols % 2 != 0:
                ncols -= 1
            col0 = (W - ncols) // 2
        device = c2w.device if torch.is_tensor(c2w) and c2w.is_cuda else torch.device("cuda")
        sh = (H, ncols, 3)
        ray_batch = F_.camera_ray_batch(int(H), int(W), intrinsic, c2w, near, far, 0, H * ncols, col0, ncols, device)
    else:
        rays_o, rays_d = rays[0], rays[1]                                           # RS:117-121
        sh = tuple(rays_d.shape)
        if torch.is_tensor(near) or torch.is_tensor(far):
            rays_o, rays_d = F_.f32(rays_o).reshape(-1, 3), F_.f32(rays_d).reshape(-1, 3)
            ones = torch.ones_like(rays_d[:, :1])
            viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
            ray_batch = torch.cat([rays_o, rays_d, near * ones, far * ones, viewdirs], -1)
        else:
            pre = getattr(rays, "scade_ray_batch", None)         # scade_b200.sampler already assembled the [N,11] batch
            if pre is not None and pre[1] == float(near) and pre[2] == float(far) and pre[0].shape[0] == rays_d.reshape(-1, 3).shape[0]:
                ray_batch = pre[0]
            else:
                ray_batch = F_.make_ray_batch(rays_o, rays_d, near, far)            # RS:123-141
    all_ret = batchify_rays(ray_batch, chunk, use_viewdirs, **kwargs)               # RS:147
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))   # RS:148-150
    k_extract = ['rgb_map', 'disp_map', 'acc_map']
    return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]


render_hyp = render     # RS:157-233 is a byte-for-byte copy of render in the reference


def create_nerf(args, scene_render_params, device=None, precision=None):
    """RS:422-509 without the checkpoint search: builds coarse + fine NeRF, the query object, Adam and the
    two kwargs dicts.  ``args`` needs the reference's fields (multires, multires_views, i_embed, use_viewdirs,
    N_importance, netdepth, netwidth, netdepth_fine, netwidth_fine, input_ch_cam, bb_center, bb_scale,
    netchunk_per_gpu, n_gpus, lrate, perturb, N_samples, raw_noise_std, lindisp)."""
    device = device or torch.device("cuda")
    precision = precision or getattr(args, "precision", "fp32")
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]
    model = NeRF(D=args.netdepth, W=args.netwidth, input_ch=input_ch, output_ch=output_ch, skips=skips,
                 input_ch_views=input_ch_views, input_ch_cam=args.input_ch_cam, use_viewdirs=args.use_viewdirs,
                 precision=precision).to(device)
    grad_vars, grad_names = [], []
    for name, param in model.named_parameters():
        grad_vars.append(param)
        grad_names.append(name)
    model_fine = None
    if args.N_importance > 0:
        model_fine = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, input_ch=input_ch, output_ch=output_ch,
                          skips=skips, input_ch_views=input_ch_views, input_ch_cam=args.input_ch_cam,
                          use_viewdirs=args.use_viewdirs, precision=precision).to(device)
        for name, param in model_fine.named_parameters():
            grad_vars.append(param)
            grad_names.append(name)
    network_query_fn = NetworkQuery(embed_fn, embeddirs_fn, args.bb_center, args.bb_scale,
                                    netchunk=args.netchunk_per_gpu * getattr(args, "n_gpus", 1), precision=precision)
    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))   # RS:469
    render_kwargs_train = {
        'network_query_fn': network_query_fn, 'embedded_cam': torch.tensor((), device=device),
        'perturb': args.perturb, 'N_importance': args.N_importance, 'network_fine': model_fine,
        'N_samples': args.N_samples, 'network_fn': model, 'use_viewdirs': args.use_viewdirs,
        'raw_noise_std': args.raw_noise_std,
    }
    render_kwargs_train.update(scene_render_params)
    render_kwargs_train['ndc'] = False
    render_kwargs_train['lindisp'] = args.lindisp
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    return render_kwargs_train, render_kwargs_test, 0, grad_vars, optimizer, grad_names
