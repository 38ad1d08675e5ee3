"""Single-box data parallelism for the learner (one process per GPU, NCCL over NVLink/NVSwitch).

The reference has one learner and no collectives; this is the one exchange step the
multi-GPU learner adds.  Replay blocks are sharded by actor -> rank (each rank owns an HBM block store and
a sum tree over its own slots), every rank samples its local batch, and per update there is exactly one
gradient all-reduce plus a scalar row-count reduction:

    grads   <- SUM over ranks of d(loss_sum_rank)          (4.33 M fp32 = 17.3 MB)
    rows    <- SUM over ranks of rows_rank
    g       =  grads / rows                                  -> identical clip + Adam on every rank

which is exactly the gradient of the reference's ``(is_w * (q - target)**2).mean()`` (worker.py:354) taken over
the GLOBAL batch, also when ranks hold different numbers of learning rows (ragged sequences).
Priority updates stay shard-local; IS weights are normalised by the local batch minimum (documented deviation:
per-shard stratification).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str = "nccl"):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_of_actor(actor_id: int, world: int) -> int:
    """Which rank's replay receives the blocks of a given actor."""
    return actor_id % world


def make_grad_hook(group=None):
    """grad_hook for DeviceLearner: global-mean gradient across ranks (see module docstring)."""
    state = {}

    def hook(learner):
        rows_g = state.get("rows")
        if rows_g is None or rows_g.device != learner.grads.flat.device:
            rows_g = state["rows"] = torch.zeros(1, dtype=torch.float32, device=learner.grads.flat.device)
        work = dist.all_reduce(learner.grads.flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        rows_g.copy_(learner.rows)
        dist.all_reduce(rows_g, op=dist.ReduceOp.SUM, group=group)
        work.wait()
        torch.reciprocal(rows_g, out=learner.grad_scale)
    return hook


def make_overlapped_grad_hook(learner, group=None):
    """Same reduction as make_grad_hook, but the FC/LSTM/head range of the flat gradient (98 % of the bytes) is all-reduced
    from a side stream as soon as r2d2_net_backward has finished it (r2d2_net_set_dense_grads_event), i.e. WHILE the conv
    layers' data/weight gradients are still being computed; only the 0.3 MB conv range and the row count are reduced after
    the backward pass.  NCCL backend only (the collectives are ordered by CUDA streams)."""
    from . import _lib
    from .learner_core import PARAM_NAMES
    dev = learner.grads.flat.device
    side = torch.cuda.Stream(device=dev)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))                 # materialise the cudaEvent_t
    _lib.check(_lib.lib().r2d2_net_set_dense_grads_event(learner._h, ev.cuda_event))
    dense_off = learner.grads.offsets[PARAM_NAMES.index("feature.7.weight")]
    rows_g = torch.zeros(1, dtype=torch.float32, device=dev)
    keep = {"event": ev, "stream": side}                      # owned by the hook: the library only borrows the event

    def hook(lrn):
        assert lrn is learner and keep
        flat = lrn.grads.flat
        with torch.cuda.stream(side):
            side.wait_event(ev)                               # recorded inside the backward call that just returned
            w_dense = dist.all_reduce(flat[dense_off:], op=dist.ReduceOp.SUM, group=group, async_op=True)
        w_conv = dist.all_reduce(flat[:dense_off], op=dist.ReduceOp.SUM, group=group, async_op=True)
        rows_g.copy_(lrn.rows)
        dist.all_reduce(rows_g, op=dist.ReduceOp.SUM, group=group)
        w_dense.wait()
        w_conv.wait()
        torch.reciprocal(rows_g, out=lrn.grad_scale)
    return hook


def broadcast_parameters(learner, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s online/target parameters and optimizer state."""
    for t in (learner.online.flat, learner.target.flat, learner.exp_avg, learner.exp_avg_sq):
        dist.broadcast(t, src=src, group=group)
    learner.pack(0)
    learner.pack(1)
