"""N > 1 host logic on CPU: world_size-2 gloo process groups (no GPU, no CUDA kernels)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taseg_b200 import parallel


def test_shard_indices_cover_every_scan():
    for n in (0, 1, 7, 8, 16, 23):
        for world in (1, 2, 4, 8):
            shards = [parallel.shard_indices(n, r, world) for r in range(world)]
            assert len({len(s) for s in shards}) == 1, "ranks must do equal work (weak scaling)"
            assert set(i for s in shards for i in s) == set(range(n))
            flat = [i for s in shards for i in s]
            assert len(flat) - n < world                     # at most world-1 padded repeats
    assert parallel.shard_indices(5, 1, 2) == [1, 3, 0]      # DistributedSampler order: [0,1,2,3,4,0][1::2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                               torch.nn.Linear(16, 3))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _make_net()
        reducer = parallel.GradientReducer(net.parameters(), bucket_mb=0.0005)    # several tiny buckets
        assert len(reducer.buckets) >= 3
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
        idx = parallel.shard_indices(8, rank, world)
        loss = torch.nn.functional.cross_entropy(net(x[idx]), y[idx])
        loss.backward()
        reducer.finish()
        grads = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        # a second step re-arms the buckets
        net.zero_grad()
        torch.nn.functional.cross_entropy(net(x[idx]), y[idx]).backward()
        reducer.finish()
        grads2 = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        t = parallel.max_over_ranks(10.0 + rank, "cpu")
        out[rank] = (grads, grads2, t)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_allreduce_matches_single_process_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90)
        assert p.exitcode == 0
    # reference: one process, the whole batch (mean over 8 = mean of the two half-batch means)
    net = _make_net()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
    torch.nn.functional.cross_entropy(net(x), y).backward()
    want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    for r in range(world):
        grads, grads2, t = out[r]
        assert torch.allclose(grads, want, atol=1e-6), "averaged gradients differ from the single-process gradient"
        assert torch.allclose(grads2, want, atol=1e-6)
        assert t == 11.0                                    # max over ranks
    assert torch.equal(out[0][0], out[1][0]), "ranks disagree after the all-reduce"


def test_lazy_scalar_reads_like_a_float():
    """utils/lazy.py: what the model's display dictionaries and train_step(sync=False) return instead of loss.item()."""
    import numpy as np
    import torch
    from taseg_b200.utils.lazy import LazyScalar
    x = LazyScalar(torch.tensor(1.5, requires_grad=True) * 1.0)
    assert '{:.2f}'.format(x) == '1.50' and repr(x) == '1.5' and float(x) == 1.5 and x.item() == 1.5
    assert x + 1 == 2.5 and 2 * x == 3.0 and 3 / x == 2.0 and x - 0.5 == 1.0 and x < 2 and x >= 1.5 and max(x, 1.0) is x
    assert np.isfinite(x) and np.isfinite([x, x]).all() and round(x) == 2 and int(x) == 1 and bool(x)
    assert x._t is None, "the device tensor is released after the first read"
