"""CPU, world_size 2, gloo: the N>1 host logic (sequence sharding, max/sum over ranks, gradient
all-reduce).  The data path itself has no collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from eas_snn_b200 import parallel, synth
    from oracle import binning as ob
    r, w, _ = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    # every rank builds the same global batch and takes its contiguous block of windows
    x, y, t, p, off = synth.make_batch(7, 5, 24, 32, 2e2, 2e3)
    xs, ys, ts, ps, offs = parallel.shard_events(x, y, t, p, off, rank, world)
    lo, hi = parallel.shard_range(5, rank, world)
    assert len(offs) == hi - lo + 1 and offs[0] == 0 and offs[-1] == len(xs)
    local = ob.micro_sum_batch(xs, ys, ts, ps, offs, 24, 32, 4)
    full = ob.micro_sum_batch(x, y, t, p, off, 24, 32, 4)
    assert np.array_equal(local, full[lo:hi])          # sharding == slicing the unsharded result
    total = parallel.sum_over_ranks(float(len(xs)))
    assert total == float(len(x))
    assert parallel.max_over_ranks(float(rank + 1)) == float(world)
    # gradient all-reduce (the only training collective, trainer.py:176)
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    for i, q_ in enumerate(lin.parameters()):
        q_.grad = torch.full_like(q_, float(rank + 1 + i))
    parallel.allreduce_gradients(lin.parameters(), bucket_bytes=16)
    for i, q_ in enumerate(lin.parameters()):
        want = sum(r_ + 1 + i for r_ in range(world)) / world
        assert torch.allclose(q_.grad, torch.full_like(q_, want))
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def test_two_rank_sharding_and_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=120)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    from eas_snn_b200 import parallel
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
