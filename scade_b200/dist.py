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


def _nets_of(render_kwargs):
    from . import render as R_
    coarse = R_._unwrap(render_kwargs["network_fn"])
    fine = R_._unwrap(render_kwargs.get("network_fine") or render_kwargs["network_fn"])
    return coarse, fine


class _BucketOverlap:
    """Gradient exchange of a flat-storage train step in two buckets: the fine network's gradient range
    is all-reduced asynchronously the moment its backward kernels are enqueued (functional._mlp_backward calls the hook), so
    the transfer runs on NCCL's stream under the coarse network's backward; everything else (coarse gradients, scale /
    shift, the loss partial sums in the tail) follows in ONE in-place all-reduce when backward() returns -- or two, if the
    fine range does not sit at one end of the flat buffer.  Autograd runs the fine branch first (its nodes are younger)."""

    def __init__(self, flat, fine_net, group):
        self.flat, self.group, self.work, self.handle = flat, group, None, None
        self.range = flat.range_of(list(fine_net.parameters())) if fine_net is not None else None
        total = flat.flat_grad.numel()
        if self.range is not None and _world(group)[1] > 1:
            self.handle = fine_net.handle()
            self.handle.grad_ready_hook = self._launch
        lo, hi = self.range if self.range is not None else (0, 0)
        self.rest = [(a, b) for a, b in ((0, lo), (hi, total)) if b > a] if self.range is not None else [(0, total)]

    def _launch(self, _handle):
        lo, hi = self.range
        self.work = dist.all_reduce(self.flat.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """The remaining range(s) + the stream-level wait for the early bucket.  Returns the number of all-reduce calls."""
        if self.handle is not None:
            self.handle.grad_ready_hook = None
        calls = 0
        if _world(self.group)[1] > 1:
            pieces = self.rest if self.work is not None else [(0, self.flat.flat_grad.numel())]
            for lo, hi in pieces:
                dist.all_reduce(self.flat.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
                calls += 1
            if self.work is not None:
                self.work.wait()                         # current stream waits for the early bucket (no host block)
                calls += 1
        return calls


def sharded_train_step(ray_batch, target_s, target_h, scale, shift, render_kwargs, n_global=None,
                       space_carving_weight=0.007, threshold=0.0, mask=None, t_rand=None, u_coarse=None, u_fine=None,
                       group=None, flat=None, is_joint=False, extra_params=None, overlap=True):
    """One SCADE training step (RS:954-985) on this rank's ray shard, followed by the gradient all-reduce.

    ray_batch [n_local,11], target_s [n_local,3], target_h [K,n_local,1] are this rank's slices of the step's
    N_rand rays (all from one image, same scale/shift on every rank, RS:945-952).  After the call every rank holds
    the global gradient in ``.grad`` of the network parameters / scale / shift, exactly as after ``loss.backward()``
    on one GPU.  Returns a dict of GLOBAL losses (tensors).

    flat         the FlatParams holding the networks' parameters and scale / shift, if the caller flattened them: the
                 exchange is then in place on its gradient buffer, in two buckets (fine net early, see _BucketOverlap).
    is_joint     RS:963 passes args.is_joint: the space-carving loss averages over ALL rays before the min over K
                 (H:115-119), which costs one extra all-reduce of [K, N_importance] partial sums.  As in the
                 reference the flag also reaches render_rays, where it only selects joint uniforms when perturb > 0.
    extra_params leaf tensors whose .grad must be reduced as well when the step's scale / shift are NOT leaves -- the
                 reference idiom ``DEPTH_SCALES[img_i]`` (RS:945-954) hands a view to the loss and autograd fills
                 ``DEPTH_SCALES.grad``; pass ``[DEPTH_SCALES, DEPTH_SHIFTS]`` (not needed when they live in `flat`)."""
    from . import functional as F_
    from . import nerf_helpers as NH
    from . import render as R_
    rank, world = _world(group)
    n_local = ray_batch.shape[0]
    n_global = int(n_global if n_global is not None else n_local * world)
    coarse, fine = _nets_of(render_kwargs)
    use_flat = flat is not None and flat.intact()
    if not use_flat:
        leaves = list(extra_params or [])
        for t in (scale, shift):
            if torch.is_tensor(t) and t.requires_grad:
                if t.is_leaf:
                    leaves.append(t)
                elif not extra_params:
                    raise ValueError("sharded_train_step: scale / shift require grad but are not leaf tensors (e.g. "
                                     "DEPTH_SCALES[img_i]); pass the leaves as extra_params=[DEPTH_SCALES, DEPTH_SHIFTS] or keep "
                                     "them in a FlatParams -- their gradients would otherwise stay rank-local")
    # RS:954 (target_h * scale + shift) is fused into the space-carving kernel when scale / shift are single CUDA elements
    affine = (not is_joint and torch.is_tensor(scale) and torch.is_tensor(shift) and scale.numel() == 1 and shift.numel() == 1
              and scale.is_cuda and shift.is_cuda and target_h.dim() == 3 and target_h.shape[-1] == 1)
    th = None if affine else target_h * scale + shift                                   # RS:954
    if (t_rand is None and u_coarse is None and u_fine is None and float(render_kwargs.get("perturb", 0.)) > 0. and not is_joint):
        # the step's three uniform draws (RS:570, H:350 twice) as ONE generator launch carved into views (same distribution
        # as the reference's three torch.rand calls, a different position in the Philox stream)
        Ns, Ni = int(render_kwargs["N_samples"]), int(render_kwargs["N_importance"])
        r = torch.rand(n_local * (Ns + 2 * Ni), device=ray_batch.device)
        t_rand = r[:n_local * Ns].view(n_local, Ns)
        u_coarse = r[n_local * Ns:n_local * (Ns + Ni)].view(n_local, Ni)
        u_fine = r[n_local * (Ns + Ni):].view(n_local, Ni)
    bucket = _BucketOverlap(flat, fine if (overlap and fine is not coarse) else None, group) if use_flat else None
    try:
        return _sharded_train_step_body(ray_batch, target_s, target_h, scale, shift, render_kwargs, n_global, n_local, world,
                                        space_carving_weight, threshold, mask, t_rand, u_coarse, u_fine, group, flat, is_joint,
                                        use_flat, affine, th, bucket, coarse, fine, list(extra_params or []) if use_flat else leaves)
    finally:
        if bucket is not None and bucket.handle is not None:
            bucket.handle.grad_ready_hook = None          # never leave the early-bucket hook armed after a failed step


def _reduce_grads(tensors, group):
    """In-place sum over the ranks of the .grad of a few small leaf tensors (scale / shift kept outside the flat storage)."""
    if _world(group)[1] > 1:
        for t in tensors:
            if t.grad is not None:
                dist.all_reduce(t.grad, op=dist.ReduceOp.SUM, group=group)


def _outside_flat(flat, scale, shift, extra_params):
    """Leaf tensors that receive gradients in the step but do not live in `flat` (their .grad needs its own all-reduce).
    Raises for non-leaf scale / shift whose leaves were not named."""
    inside = {id(p) for p in flat.params}
    out = [t for t in (extra_params or []) if id(t) not in inside]
    for t in (scale, shift):
        if not (torch.is_tensor(t) and t.requires_grad):
            continue
        if t.is_leaf and id(t) not in inside and all(t is not o for o in out):
            out.append(t)
        # (a non-leaf scale / shift such as DEPTH_SCALES[img_i] sends its gradient to its table: inside `flat`, or named in
        #  extra_params)
    return out


def _sharded_train_step_body(ray_batch, target_s, target_h, scale, shift, render_kwargs, n_global, n_local, world,
                             space_carving_weight, threshold, mask, t_rand, u_coarse, u_fine, group, flat, is_joint, use_flat,
                             affine, th, bucket, coarse, fine, leaves):
    from . import functional as F_
    from . import nerf_helpers as NH
    from . import render as R_
    kw = {k: v for k, v in render_kwargs.items() if k not in ("retraw", "is_joint", "cached_u", "t_rand", "u_coarse", "use_viewdirs")}
    ret = R_.render_rays(ray_batch, True, cached_u=u_fine, t_rand=t_rand, u_coarse=u_coarse, retraw=False, is_joint=is_joint, **kw)
    if use_flat and affine:
        # Loss heads without autograd nodes (3 launches): each kernel writes its local-sum / GLOBAL-count loss straight into a
        # slot of the flat gradient buffer's tail (so the partials ride in the exchange without a stack / copy) and returns the
        # gradient of the total loss w.r.t. its input; d scale / d shift are accumulated into the parameters' .grad slots by
        # the space-carving kernel itself when they are leaves with flat gradients, else handed to autograd.
        tail = flat.tail()
        d_rgb = F_.img2mse_head(ret["rgb_map"], target_s, n_global * 3, 1.0, tail[0:1])                   # RS:968
        d_rgb0 = F_.img2mse_head(ret["rgb0"], target_s, n_global * 3, 1.0, tail[2:3])                     # RS:981
        direct_ss = all(t.is_leaf and t.grad is not None and t.grad.is_cuda and t.grad.is_contiguous() for t in (scale, shift))
        if direct_ss:
            d_sc, d_sh = scale.grad.reshape(-1), shift.grad.reshape(-1)
        else:
            d_ss = torch.empty(2, dtype=torch.float32, device=ray_batch.device)
            d_sc, d_sh = d_ss[0:1], d_ss[1:2]
        d_pred = F_.space_carving_affine_head(ret["pred_hyp"], target_h, scale, shift, mask, threshold, n_global,
                                              space_carving_weight, tail[1:2], d_sc, d_sh, accumulate=direct_ss)   # RS:954,974
        outs, grads = [ret["rgb_map"], ret["rgb0"], ret["pred_hyp"]], [d_rgb, d_rgb0, d_pred]
        if not direct_ss:
            for t, g in ((scale, d_sc), (shift, d_sh)):
                if t.requires_grad:
                    outs.append(t)
                    grads.append(g.reshape(t.shape))
        torch.autograd.backward(outs, grads)                                            # RS:985
        bucket.finish()
        _reduce_grads(_outside_flat(flat, scale, shift, leaves), group)
        losses = tail[:3].clone()
        return {"img_loss": losses[0], "space_carving": losses[1], "img_loss0": losses[2],
                "loss": torch.add(losses[0] + losses[2], losses[1], alpha=float(space_carving_weight))}
    # local sums divided by the global count: sum over ranks == the single-GPU mean (H:11, H:125-126)
    img_loss = F_img2mse(ret["rgb_map"], target_s, n_global * 3)
    img_loss0 = F_img2mse(ret["rgb0"], target_s, n_global * 3)
    if is_joint:
        # the GLOBAL loss on every rank (one [K,P] all-reduce inside); its share of the rank-summed partials is 1/world
        sc_full = F_.space_carving_loss_joint_sharded(ret["pred_hyp"], th, n_global, mask=mask, threshold=threshold, group=group)
        sc, sc_part = sc_full, sc_full.detach() / float(world)
    elif affine:
        sc = F_.space_carving_loss_affine(ret["pred_hyp"], target_h, scale, shift, mask=mask, threshold=threshold,
                                          denominator=n_global)                          # local sum / GLOBAL ray count
        sc_part = sc.detach()
    else:
        sc = NH.compute_space_carving_loss(ret["pred_hyp"], th, is_joint=False, mask=mask, threshold=threshold) \
            * (float(n_local) / float(n_global))
        sc_part = sc.detach()
    loss = img_loss + space_carving_weight * sc + img_loss0                             # RS:976,983
    loss.backward()                                                                     # RS:985
    losses = torch.stack([img_loss.detach(), sc_part, img_loss0.detach()])
    if use_flat:
        k = losses.numel()
        tail = flat.tail()
        tail[:k].copy_(losses)
        bucket.finish()
        _reduce_grads(_outside_flat(flat, scale, shift, leaves), group)
        losses = tail[:k].clone()
    else:
        params = [p for net in dict.fromkeys([coarse, fine]) for p in net.parameters() if p.requires_grad]
        for p in params + leaves:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        FlatAllReduce([p.grad for p in params] + [t.grad for t in leaves] + [losses]).all_reduce(group)
    return {"img_loss": losses[0], "space_carving": losses[1], "img_loss0": losses[2],
            "loss": losses[0] + space_carving_weight * losses[1] + losses[2]}


class GraphedTrainStep:
    """One SCADE training step (zero_grad, sharded_train_step, optimizer steps; RS:954-997) recorded once as a CUDA graph and
    replayed: forward, losses, backward, the gradient all-reduces and the fused Adam launches cost one graph launch per step
    instead of ~100 kernel / collective launches from Python (which is what bounds the step once the rays are sharded 8 ways).

        step = GraphedTrainStep(render_kwargs, DEPTH_SCALES, DEPTH_SHIFTS, flat, [opt, opt_ss], n_global=4096)
        losses = step(ray_batch, target_s, target_h, img_i)   # this rank's shard; shapes must not change between calls

    scale / shift are either single-element tensors (one global scale / shift) or the reference's per-image tables
    ``DEPTH_SCALES`` / ``DEPTH_SHIFTS`` of shape [n_img, 1] (RS:878-879): then every call names the step's image `img_i`
    (RS:945-948) and the graph gathers row img_i through a device-side index, so replays follow the image; the gradient
    lands in row img_i of the table's .grad like ``DEPTH_SCALES[img_i]`` does under autograd.

    Requirements: parameters and scale / shift tables in one FlatParams (`flat`), optimizers = FusedAdam(..., flat=flat,
    capturable=True).  The first `warmup` calls run eagerly (they are real steps), the next call captures and replays.
    Learning-rate changes through param_groups (update_learning_rate, RS:990) are uploaded before the next replay."""

    def __init__(self, render_kwargs, scale, shift, flat, optimizers, n_global=None, space_carving_weight=0.007, threshold=0.0,
                 group=None, warmup=3, is_joint=False, overlap=True):
        self.kw, self.scale, self.shift, self.flat, self.opts = render_kwargs, scale, shift, flat, list(optimizers)
        self.n_global, self.scw, self.thr, self.group, self.warmup = n_global, space_carving_weight, threshold, group, int(warmup)
        self.is_joint, self.overlap = bool(is_joint), bool(overlap)
        self.calls, self.graph, self.static, self.losses = 0, None, None, None
        self.launches_per_step = None                     # library kernel launches recorded in the graph (diagnostic)
        for o in self.opts:
            if not getattr(o, "capturable", False):
                raise ValueError("GraphedTrainStep needs FusedAdam(..., capturable=True) optimizers")
        self.per_image = torch.is_tensor(scale) and scale.numel() > 1
        if self.per_image and (scale.dim() != 2 or shift.shape != scale.shape):
            raise ValueError("per-image scale / shift tables must both have shape [n_img, 1] (RS:878-879)")
        self.img_index = torch.zeros(1, dtype=torch.int64, device=scale.device) if self.per_image else None

    def release(self):
        """Drop the graph (and its private memory pool).  Call before destroying the process group: a live graph holds captured
        NCCL work."""
        torch.cuda.synchronize()
        self.graph, self.static, self.losses = None, None, None

    def _body(self, rb, ts, th):
        if self.flat is not None and self.flat.intact():
            self.flat.zero_grad()                             # one memset for every gradient + the loss-partial tail
        else:
            for o in self.opts:
                o.zero_grad(set_to_none=False)
        if self.per_image:                                  # curr_scale = DEPTH_SCALES[img_i] (RS:947-948), index on the device
            scale = self.scale.index_select(0, self.img_index).reshape(1)
            shift = self.shift.index_select(0, self.img_index).reshape(1)
        else:
            scale, shift = self.scale, self.shift
        losses = sharded_train_step(rb, ts, th, scale, shift, self.kw, n_global=self.n_global,
                                    space_carving_weight=self.scw, threshold=self.thr, group=self.group, flat=self.flat,
                                    is_joint=self.is_joint, overlap=self.overlap)
        for o in self.opts:
            o.step()
        return losses

    def __call__(self, ray_batch, target_s, target_h, img_i=None):
        from .optim import note_replay
        if self.per_image:
            if img_i is None:
                raise ValueError("GraphedTrainStep with per-image scale / shift tables needs img_i on every call")
            self.img_index.fill_(int(img_i))
        self.calls += 1
        if self.graph is None and self.calls <= self.warmup:
            return self._body(ray_batch, target_s, target_h)
        if self.graph is None:
            self.static = (ray_batch.clone(), target_s.clone(), target_h.clone())
            # the fp16 weight streams must be re-packed INSIDE the graph (every replay follows an optimizer step): drop the
            # cached streams so that the capture below records the pack launches whatever ran between the last step and now
            for net in _nets_of(self.kw):
                net.handle().invalidate_packed()
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
