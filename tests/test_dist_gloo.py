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
    # flat layout with a padding slot: the row count rides inside the single gradient all-reduce
    from r2d2_b200.dist import global_is_factor
    g2 = torch.cat([g, torch.zeros(4)])             # a last tensor of one element (index 1000, zero gradient) padded to four
    ln2 = _FakeLearner(g2.clone(), rows)
    ln2.grads.offsets = [0, 1000, 1004]
    make_grad_hook()(ln2)
    out[rank + 10] = (ln2.grads.flat.clone(), float(ln2.grad_scale.item()))
    # importance weights of one global sampler: per-rank factor from one MIN reduction
    m = torch.tensor([[0.002, 0.0005][rank]], dtype=torch.float64)
    out[rank + 20] = float(global_is_factor(m, 0.6))
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
    for r in (0, 1):
        flat, sc = out[r + 10]
        torch.testing.assert_close(flat[:1000], g0 + g1)
        assert float(flat[1000:].abs().max()) == 0.0 and sc == pytest.approx(1.0 / (r0 + r1))   # padding zero again
    assert out[20] == pytest.approx((0.002 / 0.0005) ** -0.6) and out[21] == pytest.approx(1.0)
