"""Ray-sharded multi-GPU execution: one process per GPU, torch.distributed for the plumbing.

The reference scales with nn.DataParallel (RS:438,455): one process, weights re-broadcast and activations
scattered/gathered through GPU 0 on every forward.  Rays are independent, so here every rank
keeps resident weights and renders / differentiates its own contiguous slice of the ray list:

  * render: no data-path collective; one all_gather of the finished per-ray outputs (<= 32 B/ray) at the end;
  * train : local losses are normalised by the GLOBAL ray count, so the sum over ranks of the local gradients
            equals the single-GPU gradient; ONE all-reduce per step over a flat fp32 buffer holding the gradients
            of both networks, d_scale, d_shift and the three loss partial sums (4.72 MB for 8x256 nets).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of n items for `rank` of `world` (first n % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class FlatAllReduce:
    """Packs tensors into one flat fp32 buffer, all-reduces it once (sum) and copies the results back.
    Tensors may be None (skipped).  Works with any backend (nccl on GPUs, gloo in the CPU tests)."""

    def __init__(self, tensors):
        self.tensors = [t for t in tensors if t is not None]
        self.numel = sum(t.numel() for t in self.tensors)
        ref = self.tensors[0]
        self.flat = torch.empty(self.numel, dtype=torch.float32, device=ref.device)

    def pack(self):
        off = 0
        for t in self.tensors:
            n = t.numel()
            self.flat[off:off + n].copy_(t.reshape(-1))
            off += n
        return self.flat

    def unpack(self):
        off = 0
        for t in self.tensors:
            n = t.numel()
            t.copy_(self.flat[off:off + n].view_as(t))
            off += n

    def all_reduce(self, group=None):
        self.pack()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.unpack()
        return self.flat


def sharded_train_step(ray_batch, target_s, target_h, scale, shift, render_kwargs, n_global=None,
                       space_carving_weight=0.007, threshold=0.0, mask=None, t_rand=None, u_coarse=None, u_fine=None,
                       group=None, flat=None):
    """One SCADE training step (RS:954-985) on this rank's ray shard, followed by the single gradient all-reduce.

    ray_batch [n_local,11], target_s [n_local,3], target_h [K,n_local,1] are this rank's slices of the step's
    N_rand rays (all from one image, same scale/shift on every rank, RS:945-952).  After the call every rank holds
    the global gradient in ``.grad`` of the network parameters / scale / shift, exactly as after ``loss.backward()``
    on one GPU.  Returns a dict of GLOBAL losses (tensors).  `flat`: the FlatParams holding the networks' parameters and
    scale / shift, if the caller flattened them (then the all-reduce is in place on its gradient buffer)."""
    from . import nerf_helpers as NH
    from . import render as R_
    rank, world = _world(group)
    n_local = ray_batch.shape[0]
    n_global = int(n_global if n_global is not None else n_local * world)
    th = target_h * scale + shift                                                       # RS:954
    ret = R_.render_rays(ray_batch, True, cached_u=u_fine, t_rand=t_rand, u_coarse=u_coarse, retraw=False, **render_kwargs)
    # local sums divided by the global count: sum over ranks == the single-GPU mean (H:11, H:125-126)
    img_loss = F_img2mse(ret["rgb_map"], target_s, n_global * 3)
    img_loss0 = F_img2mse(ret["rgb0"], target_s, n_global * 3)
    sc = NH.compute_space_carving_loss(ret["pred_hyp"], th, is_joint=False, mask=mask, threshold=threshold) \
        * (float(n_local) / float(n_global))
    loss = img_loss + space_carving_weight * sc + img_loss0                             # RS:976,983
    loss.backward()                                                                     # RS:985
    losses = torch.stack([img_loss.detach(), sc.detach(), img_loss0.detach()])
    if flat is not None and flat.intact():
        losses = flat_exchange(flat, losses, group)
        return {"img_loss": losses[0], "space_carving": losses[1], "img_loss0": losses[2],
                "loss": losses[0] + space_carving_weight * losses[1] + losses[2]}
    nets = [R_._unwrap(render_kwargs["network_fn"]), R_._unwrap(render_kwargs["network_fine"] or render_kwargs["network_fn"])]
    params = [p for net in dict.fromkeys(nets) for p in net.parameters() if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    extras = [t.grad for t in (scale, shift) if torch.is_tensor(t) and t.requires_grad and t.grad is not None]
    FlatAllReduce([p.grad for p in params] + extras + [losses]).all_reduce(group)
    return {"img_loss": losses[0], "space_carving": losses[1], "img_loss0": losses[2],
            "loss": losses[0] + space_carving_weight * losses[1] + losses[2]}


class GraphedTrainStep:
    """One SCADE training step (zero_grad, sharded_train_step, optimizer steps; RS:954-997) recorded once as a CUDA graph and
    replayed: forward, losses, backward, the gradient all-reduce and the fused Adam launches cost one graph launch per step
    instead of ~100 kernel / collective launches from Python (which is what bounds the step once the rays are sharded 8 ways).

        step = GraphedTrainStep(render_kwargs, scale, shift, flat, [opt, opt_ss], n_global=4096)
        losses = step(ray_batch, target_s, target_h)       # this rank's shard; shapes must not change between calls

    Requirements: parameters and scale / shift in one FlatParams (`flat`), optimizers = FusedAdam(..., flat=flat,
    capturable=True).  The first `warmup` calls run eagerly (they are real steps), the next call captures and replays.
    Learning-rate changes through param_groups (update_learning_rate, RS:990) are uploaded before the next replay."""

    def __init__(self, render_kwargs, scale, shift, flat, optimizers, n_global=None, space_carving_weight=0.007, threshold=0.0,
                 group=None, warmup=3):
        self.kw, self.scale, self.shift, self.flat, self.opts = render_kwargs, scale, shift, flat, list(optimizers)
        self.n_global, self.scw, self.thr, self.group, self.warmup = n_global, space_carving_weight, threshold, group, int(warmup)
        self.calls, self.graph, self.static, self.losses = 0, None, None, None
        self.launches_per_step = None                     # library kernel launches recorded in the graph (diagnostic)
        for o in self.opts:
            if not getattr(o, "capturable", False):
                raise ValueError("GraphedTrainStep needs FusedAdam(..., capturable=True) optimizers")

    def release(self):
        """Drop the graph (and its private memory pool).  Call before destroying the process group: a live graph holds captured
        NCCL work."""
        torch.cuda.synchronize()
        self.graph, self.static, self.losses = None, None, None

    def _body(self, rb, ts, th):
        for o in self.opts:
            o.zero_grad(set_to_none=False)
        losses = sharded_train_step(rb, ts, th, self.scale, self.shift, self.kw, n_global=self.n_global,
                                    space_carving_weight=self.scw, threshold=self.thr, group=self.group, flat=self.flat)
        for o in self.opts:
            o.step()
        return losses

    def __call__(self, ray_batch, target_s, target_h):
        from .optim import note_replay
        self.calls += 1
        if self.graph is None and self.calls <= self.warmup:
            return self._body(ray_batch, target_s, target_h)
        if self.graph is None:
            self.static = (ray_batch.clone(), target_s.clone(), target_h.clone())
            torch.cuda.synchronize()
            from . import _lib
            l0 = _lib.load().scade_kernel_launch_count()
            self.graph = torch.cuda.CUDAGraph()
            # thread_local: other threads of the process (the NCCL watchdog polling its events) must not invalidate the capture
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.losses = self._body(*self.static)
            self.launches_per_step = int(_lib.load().scade_kernel_launch_count() - l0)
            for o in self.opts:
                o._step -= 1                                 # the capture only recorded the step; the replay below takes it
        else:
            for dst, src in zip(self.static, (ray_batch, target_s, target_h)):
                if dst.shape != src.shape:
                    raise ValueError(f"GraphedTrainStep was captured for shapes {tuple(dst.shape)}, got {tuple(src.shape)}")
        for dst, src in zip(self.static, (ray_batch, target_s, target_h)):
            dst.copy_(src, non_blocking=True)
        for o in self.opts:                                   # a learning rate changed through param_groups applies to THIS step
            lr = float(o.param_groups[0]["lr"])
            if o._lr_uploaded != lr:
                o._lr_t.fill_(lr)
                o._lr_uploaded = lr
        self.graph.replay()
        for o in self.opts:
            note_replay(o, upload_lr=False)
        return self.losses


def flat_exchange(flat, partials, group=None):
    """Gradient exchange when parameters / gradients live in flat buffers (scade_b200.optim.FlatParams): the loss partial
    sums ride in the spare tail behind the gradients and the exchange is ONE in-place all-reduce of that buffer -- nothing
    is packed or copied.  Returns the reduced partial sums."""
    k = partials.numel()
    tail = flat.tail()
    tail[:k].copy_(partials)
    if _world(group)[1] > 1:
        dist.all_reduce(flat.flat_grad, op=dist.ReduceOp.SUM, group=group)
    return tail[:k].clone()


def F_img2mse(x, y, denominator):
    from . import functional as F_
    return F_.img2mse(x, y, denominator)


def _batchify_graphed(rays, chunk, kw, keys, cache):
    """batchify_rays (RS:66-78) with every full chunk replayed through a cached GraphedRenderRays (one CUDA graph per chunk size);
    only `keys` are kept (the graph's output tensors are overwritten by the next replay, so they are copied out)."""
    from . import render as R_
    pieces = {k: [] for k in keys}
    for i in range(0, rays.shape[0], chunk):
        part = rays[i:i + chunk]
        if part.shape[0] == chunk:
            g = cache.get(chunk)
            if g is None:
                g = cache[chunk] = R_.GraphedRenderRays(chunk, True, device=rays.device, **kw)
            ret = g(part)
            for k in keys:
                pieces[k].append(ret[k].clone())
        else:
            ret = R_.render_rays(part, True, **kw)
            for k in keys:
                pieces[k].append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in pieces.items()}


def render_image_sharded(H, W, intrinsic, c2w, near, far, render_kwargs, chunk=1024 * 16, keys=("rgb_map", "depth_map", "acc_map"),
                         group=None, gather=True, device=None, graph_cache=None):
    """Full-image render (RS:106-108,147) with the pixel list split across ranks.  Every rank builds the rays of its
    own pixel range on the device (no full-image get_rays + scatter), renders them in `chunk`-ray pieces and, if
    `gather`, all ranks end with the [H,W,...] maps.  `graph_cache` (a dict the caller keeps across frames, eval only: the
    weights must not change) replays full chunks as CUDA graphs."""
    from . import functional as F_
    from . import render as R_
    rank, world = _world(group)
    device = device or torch.device("cuda", torch.cuda.current_device())
    n = H * W
    lo, hi = shard_range(n, rank, world)
    rays = F_.camera_ray_batch(H, W, intrinsic, c2w, near, far, pix0=lo, n=hi - lo, device=device)
    kw = {k: v for k, v in render_kwargs.items() if k not in ("near", "far", "ndc", "use_viewdirs")}
    with torch.no_grad():
        if graph_cache is not None:
            ret = _batchify_graphed(rays, chunk, kw, keys, graph_cache)
        else:
            ret = R_.batchify_rays(rays, chunk, True, **kw)
    out = {}
    for k in keys:
        local = ret[k].reshape(hi - lo, -1)
        if world > 1 and gather:
            width = local.shape[1]
            per = (n + world - 1) // world
            pad = torch.zeros((per, width), dtype=local.dtype, device=local.device)
            pad[:hi - lo] = local
            full = torch.empty((world * per, width), dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(full, pad, group=group)
            pieces = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                pieces.append(full[r * per:r * per + (b - a)])
            local = torch.cat(pieces, 0)
            out[k] = local.reshape(H, W, -1).squeeze(-1)
        else:
            out[k] = local.squeeze(-1)
    return out
