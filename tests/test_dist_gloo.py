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
