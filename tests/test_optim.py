"""Flat parameter storage + FusedAdam.  Host logic on CPU; the kernel against torch.optim.Adam on GPU."""
import numpy as np
import pytest
import torch

from scade_b200.optim import FusedAdam, flatten_parameters


def test_flatten_keeps_values_and_aliases_grads():
    torch.manual_seed(0)
    a, b = torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)
    s = torch.ones(1, requires_grad=True)
    w0, b0 = a.weight.detach().clone(), b.bias.detach().clone()
    flat = flatten_parameters(a, b, [s], a)                       # duplicates are dropped
    assert len(flat.params) == 5 and flat.intact()
    assert torch.equal(a.weight, w0) and torch.equal(b.bias, b0)
    assert all(o % 4 == 0 for o in flat.offsets)                  # 16-byte aligned views
    (b(a(torch.randn(4, 5))).sum() * s).backward()                # autograd accumulates in place into the flat gradient
    assert flat.intact() and float(flat.grads().abs().sum()) > 0
    assert torch.equal(flat.flat_grad[flat.offsets[0]:flat.offsets[0] + 15].view(3, 5), a.weight.grad)
    nets = FusedAdam(list(a.parameters()) + list(b.parameters()), lr=5e-4, flat=flat)
    ss = FusedAdam([s], lr=1e-6, flat=flat)
    assert nets._range[:2] == (0, flat.offsets[4]) and ss._range[:2] == (flat.offsets[4], flat.numel)
    nets.zero_grad()
    assert float(a.weight.grad.abs().sum()) == 0 and float(s.grad.abs().sum()) > 0
    ss.zero_grad()
    assert float(flat.flat_grad.abs().sum()) == 0
    with pytest.raises(ValueError):
        FusedAdam([a.bias, b.bias], flat=flat)                    # not a consecutive run
    a.weight.grad = None
    assert not flat.intact()
    # the reference's LR schedule writes param_groups[...]['lr'] (train_utils/hyperparameter_update.py:3-5)
    for g in nets.param_groups:
        g["lr"] = 1e-5
    assert nets.param_groups[0]["lr"] == 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("use_flat", [True, False])
def test_fused_adam_matches_torch_adam(use_flat):
    from scade_b200 import _lib
    _lib.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    shapes = [(256, 57), (256,), (256, 313), (1, 256), (3,), (1,)]
    ref = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    opt_ref = torch.optim.Adam(ref, lr=5e-4, betas=(0.9, 0.999))
    if use_flat:
        flat = flatten_parameters(ours)
        opt = FusedAdam(flat, lr=5e-4, betas=(0.9, 0.999))
    else:
        opt = FusedAdam(ours, lr=5e-4, betas=(0.9, 0.999))
    g = torch.Generator(device=dev).manual_seed(2)
    for step in range(6):
        if step == 3:                                             # staircase decay (train_utils/hyperparameter_update.py:9-13)
            for o in (opt_ref, opt):
                for grp in o.param_groups:
                    grp["lr"] = 5e-5
        opt_ref.zero_grad(set_to_none=False)
        opt.zero_grad(set_to_none=False)
        for a, b in zip(ref, ours):
            grad = torch.randn(a.shape, device=dev, generator=g) * (10.0 ** (step - 3))
            a.grad = grad.clone() if a.grad is None else a.grad.copy_(grad)
            if b.grad is None:
                b.grad = grad.clone()
            else:
                b.grad.copy_(grad)
        opt_ref.step()
        opt.step()
        for a, b in zip(ref, ours):
            np.testing.assert_allclose(b.detach().cpu().numpy(), a.detach().cpu().numpy(), rtol=2e-6, atol=1e-7)
    sd = opt.state_dict()                                         # torch.optim.Adam's layout: step / exp_avg / exp_avg_sq per parameter
    st0, ref0 = sd["state"][0], opt_ref.state_dict()["state"][0]
    assert int(st0["step"]) == 6 and st0["exp_avg"].shape == tuple(shapes[0])
    np.testing.assert_allclose(st0["exp_avg_sq"].cpu().numpy(), ref0["exp_avg_sq"].cpu().numpy(), rtol=1e-5, atol=1e-12)


def test_learning_rate_schedule_matches_reference_formula():
    """train_utils/hyperparameter_update.py:9-15: staircase decay, and update_learning_rate writes every param group."""
    from scade_b200.optim import get_learning_rate, update_learning_rate
    assert get_learning_rate(5e-4, 0, 400000, 0.1) == 5e-4
    assert get_learning_rate(5e-4, 399999, 400000, 0.1) == 5e-4
    assert abs(get_learning_rate(5e-4, 400000, 400000, 0.1) - 5e-5) < 1e-12
    assert abs(get_learning_rate(5e-4, 200000, 400000, 0.1, staircase=False) - 5e-4 * 0.1 ** 0.5) < 1e-12
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    update_learning_rate(opt, 0.25)
    assert all(g["lr"] == 0.25 for g in opt.param_groups)


def test_graphed_train_step_requires_capturable_optimizers():
    """Host-side contract of scade_b200.dist.GraphedTrainStep (the replay itself is a GPU test): optimizers whose step count
    lives on the host cannot be recorded in a CUDA graph."""
    from scade_b200.dist import GraphedTrainStep
    lin = torch.nn.Linear(4, 4)
    flat = flatten_parameters(lin)
    eager = FusedAdam(list(lin.parameters()), lr=1e-3, flat=flat)
    with pytest.raises(ValueError):
        GraphedTrainStep({}, None, None, flat, [eager])
    cap = FusedAdam(list(lin.parameters()), lr=1e-3, flat=flat, capturable=True)
    step = GraphedTrainStep({}, None, None, flat, [cap], n_global=8, warmup=2)
    assert step.graph is None and step.calls == 0 and cap.capturable


def test_fused_adam_load_state_dict_adopts_moments_and_step():
    """load_state_dict() on a flat FusedAdam copies exp_avg / exp_avg_sq into the flat moment buffers and takes the
    step count, also after the optimizer has been used; state_dict() reports the live step (host logic, no kernel needed)."""
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    ref = torch.optim.Adam(lin.parameters(), lr=1e-3)
    for _ in range(5):
        ref.zero_grad()
        lin(torch.randn(2, 4)).pow(2).sum().backward()
        ref.step()
    sd = ref.state_dict()
    lin2 = torch.nn.Linear(4, 3)
    flat = flatten_parameters(lin2)
    opt = FusedAdam(list(lin2.parameters()), lr=1e-3, flat=flat)
    opt._flat_state()                                   # as after a first step: the flat moment buffers exist
    opt._step = 2
    opt.load_state_dict(sd)
    assert opt._step == 5
    w, b = list(lin2.parameters())
    assert torch.equal(opt._m[:w.numel()].view_as(w), sd["state"][0]["exp_avg"])
    assert torch.equal(opt._v[flat.offsets[1]:flat.offsets[1] + b.numel()].view_as(b), sd["state"][1]["exp_avg_sq"])
    assert opt.state[w]["exp_avg"].data_ptr() == opt._m.data_ptr()
    opt._step = 9                                       # what graph replays do (note_replay)
    out = opt.state_dict()
    assert {int(v["step"]) for v in out["state"].values()} == {9}
