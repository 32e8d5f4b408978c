"""Host-side logic of the multi-GPU path on CPU with the gloo backend, world_size 2 (no kernels involved):
ray partitioning and the single flat all-reduce of gradients + loss partial sums."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scade_b200.dist import FlatAllReduce, flat_exchange, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, 307200, 307201):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    # "gradients" of two layers + scale/shift grads + three loss partial sums, different on every rank
    grads = [torch.from_numpy(rng.standard_normal(s).astype(np.float32)) for s in [(256, 57), (256,), (1,), (1,), (3,)]]
    sent = [g.clone() for g in grads]
    FlatAllReduce(grads + [None]).all_reduce()
    torch.save({"sent": sent, "got": grads}, os.path.join(out_dir, f"r{rank}.pt"))
    # a ray-sharded mean equals the global mean when every rank divides by the GLOBAL count
    n = 1000
    x = torch.from_numpy(np.random.default_rng(7).random((n, 3)).astype(np.float32))
    lo, hi = shard_range(n, rank, world)
    part = (x[lo:hi] ** 2).sum().reshape(1) / (n * 3)
    FlatAllReduce([part]).all_reduce()
    assert torch.allclose(part, (x ** 2).mean().reshape(1), rtol=1e-5)
    # flat parameter storage: gradients + loss partials in ONE in-place all-reduce (scade_b200.optim.FlatParams)
    from scade_b200.optim import flatten_parameters
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 2))
    scale = torch.ones(1, requires_grad=True)
    flat = flatten_parameters(net, [scale])
    xs = torch.from_numpy(np.random.default_rng(9).random((8, 6)).astype(np.float32))
    lo, hi = shard_range(8, rank, world)
    loss = (net(xs[lo:hi]) * scale).pow(2).sum() / 8
    loss.backward()
    red = flat_exchange(flat, torch.stack([loss.detach(), loss.detach() * 0 + 1.0]))
    torch.save({"grads": [p.grad.clone() for p in flat.params], "red": red, "intact": flat.intact()},
               os.path.join(out_dir, f"flat{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    for i in range(len(res[0]["sent"])):
        want = sum(r["sent"][i] for r in res)
        for r in res:
            assert torch.allclose(r["got"][i], want, rtol=1e-6, atol=1e-6)
    # flat path: every rank ends with the single-process gradient of the global-mean loss, and the summed partials
    from scade_b200.optim import flatten_parameters
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 2))
    scale = torch.ones(1, requires_grad=True)
    flat = flatten_parameters(net, [scale])
    xs = torch.from_numpy(np.random.default_rng(9).random((8, 6)).astype(np.float32))
    loss = (net(xs) * scale).pow(2).sum() / 8
    loss.backward()
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"flat{r}.pt"))
        assert got["intact"]
        for a, b in zip(got["grads"], [p.grad for p in flat.params]):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
        assert torch.allclose(got["red"], torch.stack([loss.detach(), torch.tensor(float(world))]), rtol=1e-5)


def _bucket_worker(rank, world, port, out_dir):
    """The two-bucket gradient exchange of a flat-storage train step (scade_b200.dist._BucketOverlap): fine-net range early and
    asynchronously, the rest (coarse net, scale / shift, loss partials in the tail) afterwards -- for the fine net at the
    front, in the middle and at the end of the flat buffer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scade_b200.dist import _BucketOverlap
    from scade_b200.nerf_helpers import NeRF
    from scade_b200.optim import flatten_parameters
    results = {}
    for order in ("fine_first", "fine_middle", "fine_last"):
        torch.manual_seed(1)
        mk = lambda: NeRF(D=2, W=8, input_ch=9, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True)
        coarse, fine = mk(), mk()
        scale, shift = torch.ones(1, requires_grad=True), torch.zeros(1, requires_grad=True)
        groups = {"fine_first": (fine, coarse, [scale, shift]), "fine_middle": (coarse, fine, [scale, shift]),
                  "fine_last": (coarse, [scale, shift], fine)}[order]
        flat = flatten_parameters(*groups)
        g = torch.Generator().manual_seed(50 + rank)
        flat.flat_grad.copy_(torch.randn(flat.flat_grad.numel(), generator=g))
        sent = flat.flat_grad.clone()
        b = _BucketOverlap(flat, fine, None)
        assert b.range is not None and b.handle is not None and fine.handle().grad_ready_hook is not None
        fine.handle().grad_ready_hook(fine.handle())          # what functional._mlp_backward does after the fine backward
        calls = b.finish()
        assert fine.handle().grad_ready_hook is None
        results[order] = {"sent": sent, "got": flat.flat_grad.clone(), "calls": calls, "range": b.range, "intact": flat.intact()}
    # scale / shift leaves kept OUTSIDE the flat storage get their own (tiny) all-reduce; those inside do not
    from scade_b200.dist import _outside_flat, _reduce_grads
    net = NeRF(D=2, W=8, input_ch=9, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True)
    s_in, s_out, h_out = (torch.ones(1, requires_grad=True) for _ in range(3))
    flat = flatten_parameters(net, [s_in])
    s_out.grad, h_out.grad = torch.full((1,), float(rank + 1)), torch.full((1,), 10.0 * (rank + 1))
    assert _outside_flat(flat, s_in, h_out, None) == [h_out]
    outside = _outside_flat(flat, s_out, h_out, None)
    assert len(outside) == 2 and outside[0] is s_out and outside[1] is h_out
    _reduce_grads(outside, None)
    results["outside"] = (float(s_out.grad), float(h_out.grad))
    torch.save(results, os.path.join(out_dir, f"bucket{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_exchange_world2(tmp_path):
    world = 2
    mp.spawn(_bucket_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"bucket{r}.pt")) for r in range(world)]
    for order, want_calls in (("fine_first", 2), ("fine_middle", 3), ("fine_last", 3)):
        want = sum(r[order]["sent"] for r in res)
        for r in res:
            assert r[order]["intact"] and r[order]["calls"] == want_calls, (order, r[order]["calls"])
            assert torch.allclose(r[order]["got"], want, rtol=1e-6, atol=1e-6), order
    assert all(r["outside"] == (3.0, 30.0) for r in res)
