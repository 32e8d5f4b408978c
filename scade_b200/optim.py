"""Flat parameter storage and a fused Adam for the SCADE training step.

The reference keeps 48 parameter tensors (two NeRFs) in one ``torch.optim.Adam`` (run_scade_scannet.py:469) and a second
Adam for the per-image depth scale / shift (RS:888); every step runs ``zero_grad`` -> ``backward`` -> ``step`` over all of
them (RS:966-997).  Here the same ``nn.Parameter`` objects are re-homed as views of ONE flat fp32 buffer, with ``.grad``
views of a second one, so that per step

  * ``zero_grad`` is one memset,
  * the CUDA backward kernels accumulate straight into ``.grad`` (no per-tensor temporaries, no AccumulateGrad adds),
  * the multi-GPU gradient exchange is one in-place all-reduce of the flat gradient (scade_b200/dist.py),
  * ``FusedAdam.step`` is one kernel launch (``scade_adam_step``).

``FusedAdam`` is a ``torch.optim.Optimizer``: ``param_groups[i]['lr']`` is honoured, so ``update_learning_rate``
(train_utils/hyperparameter_update.py:3-5) keeps working, and ``state_dict()`` has torch.optim.Adam's layout
(``step``, ``exp_avg``, ``exp_avg_sq`` per parameter), so the reference's checkpoints (RS:1006-1011) stay loadable.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

TAIL = 8        # spare floats behind the flat gradient: loss partial sums ride along in the single all-reduce


class FlatParams:
    """Re-homes `params` (fp32, same device) as views of one buffer; their gradients as views of another."""

    def __init__(self, params):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        offs, off = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatParams needs fp32 parameters on one device")
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4                   # keep every view 16-byte aligned
        self.offsets, self.numel = offs, off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(off + TAIL, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                n = p.numel()
                self.flat[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.flat[o:o + n].view(p.shape)
                p.grad = self.flat_grad[o:o + n].view(p.shape)
        from . import functional as F_
        F_.enable_direct_grads(self.params)         # the CUDA backward kernels accumulate straight into these .grad views

    def range_of(self, params):
        """[lo, hi) float offsets of `params` inside the flat buffers, or None unless they are a consecutive run of self.params."""
        ids = [id(p) for p in self.params]
        want = [id(p) for p in params]
        if not want or want[0] not in ids:
            return None
        i0 = ids.index(want[0])
        if ids[i0:i0 + len(want)] != want:
            return None
        hi = self.offsets[i0 + len(want)] if i0 + len(want) < len(ids) else self.numel
        return self.offsets[i0], hi

    def grads(self):
        return self.flat_grad[:self.numel]

    def tail(self):
        return self.flat_grad[self.numel:]

    def zero_grad(self):
        self.flat_grad.zero_()

    def intact(self):
        """True while every parameter and gradient still aliases the flat buffers (p.grad = None or p.data = ... breaks it)."""
        b, g = self.flat.data_ptr(), self.flat_grad.data_ptr()
        return all(p.data_ptr() == b + 4 * o and p.grad is not None and p.grad.data_ptr() == g + 4 * o
                   for p, o in zip(self.params, self.offsets))


def flatten_parameters(*modules_or_params):
    """FlatParams over the parameters of the given modules / iterables of parameters / parameters, in order, de-duplicated."""
    params, seen = [], set()
    for m in modules_or_params:
        it = m.parameters() if isinstance(m, torch.nn.Module) else ([m] if torch.is_tensor(m) else m)
        for p in it:
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
    return FlatParams(params)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps) semantics (RS:469) with one CUDA launch per contiguous parameter range.

    ``FusedAdam(flat)``: all parameters of a FlatParams, one launch.  ``FusedAdam(params, flat=flat)``: `params` must be a
    consecutive run of ``flat.params`` (e.g. the two networks, or scale / shift with their own learning rate, RS:888), one
    launch on that slice.  ``FusedAdam(params)``: any CUDA fp32 parameters, one launch each."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, flat=None, capturable=False):
        """capturable: keep the step count and the learning rate in device memory (scade_adam_step_graph) so that step() can be
        recorded in a CUDA graph (flat storage only); `param_groups[0]['lr']` is uploaded whenever it changed, outside captures."""
        if isinstance(params, FlatParams):
            flat, params = params, params.params
        plist = list(params)
        super().__init__(plist, dict(lr=lr, betas=betas, eps=eps))
        self.capturable = bool(capturable)
        self._step_t = self._lr_t = self._lr_uploaded = None
        self.flat, self._range = flat, None
        if flat is not None:
            ids = [id(p) for p in flat.params]
            i0 = ids.index(id(plist[0]))
            if [id(p) for p in plist] != ids[i0:i0 + len(plist)]:
                raise ValueError("FusedAdam(params, flat=...): params must be a consecutive run of flat.params")
            end = flat.offsets[i0 + len(plist)] if i0 + len(plist) < len(ids) else flat.numel
            self._range = (flat.offsets[i0], end, i0, i0 + len(plist))
        self._m = self._v = None
        self._step = 0

    def _flat_state(self):
        if self._m is None:
            f = self.flat
            o0, o1, i0, i1 = self._range
            self._m = torch.zeros(o1 - o0, dtype=torch.float32, device=f.flat.device)
            self._v = torch.zeros_like(self._m)
            for p, o in zip(f.params[i0:i1], f.offsets[i0:i1]):
                st = self.state[p]
                n, o = p.numel(), o - o0
                if "exp_avg" in st:                           # state loaded from a checkpoint: adopt it
                    self._m[o:o + n].copy_(st["exp_avg"].reshape(-1))
                    self._v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    self._step = int(st.get("step", self._step))
                st["exp_avg"], st["exp_avg_sq"] = self._m[o:o + n].view(p.shape), self._v[o:o + n].view(p.shape)
        return self._m, self._v

    def _sync_state_steps(self):
        if self.flat is not None and self._range is not None:
            for p in self.flat.params[self._range[2]:self._range[3]]:
                if p in self.state:
                    self.state[p]["step"] = self._step

    def state_dict(self):
        """torch.optim.Adam's layout.  The step count written is the CURRENT one (graph replays advance it on the device and in
        the host mirror, not in `state`), so a checkpoint taken after replays resumes with the right bias correction."""
        if self._m is not None:
            self._sync_state_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Adopts exp_avg / exp_avg_sq / step into the flat moment buffers at any time (not only before the first step), and
        resets the device-side step count of the capturable variant."""
        super().load_state_dict(state_dict)
        if self.flat is None or self._range is None:
            return
        o0, o1, i0, i1 = self._range
        params = self.flat.params[i0:i1]
        if not any("exp_avg" in self.state.get(p, {}) for p in params):
            return
        if self._m is None:
            self._m = torch.zeros(o1 - o0, dtype=torch.float32, device=self.flat.flat.device)
            self._v = torch.zeros_like(self._m)
        step = self._step
        with torch.no_grad():
            for p, o in zip(params, self.flat.offsets[i0:i1]):
                st = self.state[p]
                n, o = p.numel(), o - o0
                if "exp_avg" in st:
                    self._m[o:o + n].copy_(st["exp_avg"].reshape(-1))
                    self._v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    step = int(st.get("step", step))
                st["exp_avg"], st["exp_avg_sq"] = self._m[o:o + n].view(p.shape), self._v[o:o + n].view(p.shape)
        self._step = step
        if self._step_t is not None:
            self._step_t.fill_(self._step)
        self._lr_uploaded = None                       # param_groups came from the checkpoint: upload the rate again
        self._sync_state_steps()

    def zero_grad(self, set_to_none=False):
        if self.flat is not None and self.flat.intact():
            o0, o1 = self._range[:2]
            if o0 == 0 and o1 == self.flat.numel:
                self.flat.zero_grad()
            else:
                self.flat.flat_grad[o0:o1].zero_()
        else:
            super().zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        from . import functional as F_
        loss = closure() if closure is not None else None
        L = _lib.load()
        g0 = self.param_groups[0]
        if self.flat is not None and self.flat.intact() and len(self.param_groups) == 1:
            m, v = self._flat_state()
            f = self.flat
            o0, o1, i0, i1 = self._range
            if self.capturable:
                capturing = torch.cuda.is_current_stream_capturing()
                if self._step_t is None:
                    if capturing:
                        raise _lib.ScadeError("FusedAdam(capturable=True): take one eager step before capturing")
                    self._step_t = torch.full((1,), self._step, dtype=torch.int64, device=f.flat.device)
                    self._lr_t = torch.zeros(1, dtype=torch.float64, device=f.flat.device)
                if self._lr_uploaded is None or self._lr_uploaded != float(g0["lr"]):
                    if capturing:
                        raise _lib.ScadeError("FusedAdam(capturable=True): the learning rate changed inside a capture")
                    self._lr_t.fill_(float(g0["lr"]))
                    self._lr_uploaded = float(g0["lr"])
                self._step += 1                               # host mirror (replays are counted by note_replay)
                check(L.scade_adam_step_graph(ptr(f.flat[o0:o1]), ptr(f.flat_grad[o0:o1]), ptr(m), ptr(v), o1 - o0, ptr(self._lr_t),
                                              float(g0["betas"][0]), float(g0["betas"][1]), float(g0["eps"]), ptr(self._step_t),
                                              stream_ptr()), "scade_adam_step_graph")
            else:
                self._step += 1
                check(L.scade_adam_step(ptr(f.flat[o0:o1]), ptr(f.flat_grad[o0:o1]), ptr(m), ptr(v), o1 - o0, float(g0["lr"]),
                                        float(g0["betas"][0]), float(g0["betas"][1]), float(g0["eps"]), self._step, stream_ptr()),
                      "scade_adam_step")
            for p in f.params[i0:i1]:
                self.state[p]["step"] = self._step
            F_.mark_weights_changed()
            return loss
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"], st["exp_avg"], st["exp_avg_sq"] = 0, torch.zeros_like(p), torch.zeros_like(p)
                st["step"] = int(st["step"]) + 1
                if p.data_ptr() % 16 or p.grad.data_ptr() % 16 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise _lib.ScadeError("FusedAdam needs contiguous 16-byte aligned fp32 CUDA parameters and gradients")
                check(L.scade_adam_step(ptr(p.data), ptr(p.grad), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(),
                                        float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                                        st["step"], stream_ptr()), "scade_adam_step")
        F_.mark_weights_changed()
        return loss


def note_replay(optimizer, upload_lr=True):
    """Bookkeeping after a CUDA-graph replay that contained `optimizer.step()` (FusedAdam(capturable=True)): the host mirror of
    the step count advances, a learning rate changed through param_groups is uploaded for the NEXT replay, and the fp16 weight
    streams are marked stale for eager code."""
    from . import functional as F_
    optimizer._step += 1
    optimizer._sync_state_steps()
    if upload_lr and optimizer._lr_t is not None:
        lr = float(optimizer.param_groups[0]["lr"])
        if optimizer._lr_uploaded != lr:
            optimizer._lr_t.fill_(lr)
            optimizer._lr_uploaded = lr
    F_.mark_weights_changed()


def update_learning_rate(optimizer, learning_rate):
    """train_utils/hyperparameter_update.py:3-5 (RS:990): every param group takes the new rate; FusedAdam reads it at step()."""
    for group in optimizer.param_groups:
        group["lr"] = learning_rate


def get_learning_rate(init_learning_rate, iteration_num, decay_step, decay_rate, staircase=True):
    """train_utils/hyperparameter_update.py:9-15 (RS:988): init * decay_rate ** (iteration / decay_step), floored when `staircase`."""
    p = iteration_num / decay_step
    if staircase:
        p = int(p // 1)
    return init_learning_rate * (decay_rate ** p)
