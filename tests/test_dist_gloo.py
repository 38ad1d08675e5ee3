"""world_size-2 gloo test of the data-parallel exchange step (host logic of r2d2_b200.dist)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Flat:
    def __init__(self, t):
        self.flat = t


class _FakeLearner:
    """CPU stand-in exposing exactly what the grad hook touches."""

    def __init__(self, grads, rows):
        self.grads = _Flat(grads)
        self.rows = torch.tensor([rows], dtype=torch.int32)
        self.grad_scale = torch.ones(1)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from r2d2_b200.dist import make_grad_hook, shard_of_actor
    torch.manual_seed(rank)
    rows = [2560, 1733][rank]                      # ragged shards: different numbers of learning rows
    g = torch.randn(1000) * (rank + 1)
    ln = _FakeLearner(g.clone(), rows)
    make_grad_hook()(ln)
    out[rank] = (g, rows, ln.grads.flat.clone(), float(ln.grad_scale.item()))
    assert [shard_of_actor(a, world) for a in range(4)] == [0, 1, 0, 1]
    dist.destroy_process_group()


def test_grad_hook_gives_global_mean_gradient():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    (g0, r0, s0, sc0), (g1, r1, s1, sc1) = out[0], out[1]
    torch.testing.assert_close(s0, g0 + g1)                 # SUM of d(loss_sum)
    torch.testing.assert_close(s1, g0 + g1)
    assert sc0 == pytest.approx(1.0 / (r0 + r1)) and sc1 == sc0  # one global mean over all rows
