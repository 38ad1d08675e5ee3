"""Single-box data parallelism for the learner (one process per GPU, NCCL over NVLink/NVSwitch).

The reference has one learner and no collectives; this is the one exchange step the
multi-GPU learner adds.  Replay blocks are sharded by actor -> rank (each rank owns an HBM block store and
a sum tree over its own slots), every rank samples its local batch, and per update there is one gradient
all-reduce (issued in two pieces by the overlapped hook) that also carries the row count:

    grads   <- SUM over ranks of d(loss_sum_rank)          (4.33 M fp32 = 17.3 MB)
    rows    <- SUM over ranks of rows_rank
    g       =  grads / rows                                  -> identical clip + Adam on every rank

which is exactly the gradient of the reference's ``(is_w * (q - target)**2).mean()`` (worker.py:354) taken over
the GLOBAL batch, also when ranks hold different numbers of learning rows (ragged sequences).
Priority updates stay shard-local.  Importance weights: GlobalISWeights computes one scalar per rank that turns its
weights into those of a single prioritized sampler over all shards (one MIN all-reduce of a float64, on a side
stream); the loss is linear in the weights, so the gradient hooks apply it to this rank's gradient right before the
reduction.  What remains of the sharding is stratification (every rank contributes exactly B sequences).
"""
from __future__ import annotations

import os
from typing import Optional

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


def _rows_slot(learner):
    """Index of a padding slot of the flat gradient layout inside the dense range (value.2.bias holds one element and is
    padded to four): the row count travels there, inside the gradient all-reduce, instead of in a collective of its own.
    None for objects without a layout (the CPU stand-in of the gloo test)."""
    off = getattr(learner.grads, "offsets", None)
    if off is None or off[-1] - off[-2] < 2:
        return None
    return off[-1] - 1


def make_grad_hook(group=None, is_sync=None):
    """grad_hook for DeviceLearner: global-mean gradient across ranks (see module docstring) with ONE all-reduce: the
    local row count rides in a padding slot of the flat gradient buffer (zeroed again before the optimizer sees it).
    is_sync: optional GlobalISWeights whose per-rank factor scales this rank's gradient before the reduction."""
    state = {}

    def hook(learner):
        flat = learner.grads.flat
        slot = _rows_slot(learner)
        if is_sync is not None:
            is_sync.wait()
            flat.mul_(is_sync.factor)
        if slot is None:                                      # no layout information: separate scalar reduction
            rows_g = state.get("rows")
            if rows_g is None or rows_g.device != flat.device:
                rows_g = state["rows"] = torch.zeros(1, dtype=torch.float32, device=flat.device)
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
            rows_g.copy_(learner.rows)
            dist.all_reduce(rows_g, op=dist.ReduceOp.SUM, group=group)
            work.wait()
            torch.reciprocal(rows_g, out=learner.grad_scale)
            return
        flat[slot:slot + 1].copy_(learner.rows)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        torch.reciprocal(flat[slot:slot + 1], out=learner.grad_scale)
        flat[slot:slot + 1].zero_()
    hook.split_graph = True        # the learner may replay its gradient and optimizer phases as CUDA graphs around this (eager) hook
    return hook


def make_overlapped_grad_hook(learner, group=None, is_sync=None):
    """Same reduction as make_grad_hook, but the FC/LSTM/head range of the flat gradient (98 % of the bytes) is all-reduced
    from a side stream as soon as r2d2_net_backward has finished it (r2d2_net_set_dense_grads_event), i.e. WHILE the conv
    layers' data/weight gradients are still being computed; only the 0.3 MB conv range is reduced after the backward pass.
    Two collectives per update (the row count rides in a padding slot of the dense range).  NCCL backend only (the
    collectives are ordered by CUDA streams)."""
    from . import _lib
    from .learner_core import PARAM_NAMES
    dev = learner.grads.flat.device
    side = torch.cuda.Stream(device=dev)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))                 # materialise the cudaEvent_t
    _lib.check(_lib.lib().r2d2_net_set_dense_grads_event(learner._h, ev.cuda_event))
    dense_off = learner.grads.offsets[PARAM_NAMES.index("feature.7.weight")]
    slot = _rows_slot(learner)
    assert slot is not None and slot >= dense_off
    keep = {"event": ev, "stream": side}                      # owned by the hook: the library only borrows the event

    def hook(lrn):
        assert lrn is learner and keep
        flat = lrn.grads.flat
        with torch.cuda.stream(side):
            side.wait_event(ev)                               # recorded inside the backward call / graph that was just launched
            if is_sync is not None:
                side.wait_event(is_sync.ev)
                flat[dense_off:].mul_(is_sync.factor)         # this rank's importance-weight correction (a scalar: the loss is linear in it)
            flat[slot:slot + 1].copy_(lrn.rows)               # the row count (K2, before the backward pass) is final by then as well
            w_dense = dist.all_reduce(flat[dense_off:], op=dist.ReduceOp.SUM, group=group, async_op=True)
        if is_sync is not None:
            is_sync.wait()
            flat[:dense_off].mul_(is_sync.factor)
        w_conv = dist.all_reduce(flat[:dense_off], op=dist.ReduceOp.SUM, group=group, async_op=True)
        w_dense.wait()
        w_conv.wait()
        torch.reciprocal(flat[slot:slot + 1], out=lrn.grad_scale)
        flat[slot:slot + 1].zero_()                           # padding must be zero again: the global norm runs over the flat buffer
    # Capturing the NCCL calls themselves into the update's CUDA graph bought nothing at 2 GPUs and left the process group
    # hanging at shutdown.  Instead the learner replays TWO graphs (gradients; optimizer) around this eager hook; the event the
    # side stream waits on is recorded by an external event-record node inside the gradient graph (cudaEventRecordExternal).
    hook.capturable = False
    hook.split_graph = True
    return hook


class PeerExchange:
    """The data-parallel exchange step on our own kernels over NVLink peer memory (csrc/dp.cu) instead of NCCL calls.

    torch's symmetric memory does the plumbing only (allocation, handle exchange between the processes, the NVSwitch
    multicast mapping); the gradient all-reduce (`allreduce`, multimem.ld_reduce / multimem.st with flag barriers in peer
    memory) and the global importance-weight factor (`is_factor`) are r2d2_dp_* launches on the caller's streams.  The
    learner's flat gradient buffer is moved into symmetric memory (DeviceLearner.use_grad_buffer)."""

    def __init__(self, learner, group=None, use_multicast: Optional[bool] = None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.learner = learner
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        dev = learner.device
        lib = _lib.lib()
        with torch.cuda.device(dev):
            self.grads = symm.empty(learner.grads.flat.numel(), dtype=torch.float32, device=dev)
            self.ctl = symm.empty(lib.r2d2_dp_ctl_bytes() // 4, dtype=torch.int32, device=dev)
            self.grads.zero_()
            self.ctl.zero_()
            self._hg = symm.rendezvous(self.grads, self.group)
            self._hc = symm.rendezvous(self.ctl, self.group)
            learner.use_grad_buffer(self.grads)
            torch.cuda.synchronize(dev)
            dist.barrier(self.group)                                # every control block is zero before any rank signals
            mc = int(self._hg.multicast_ptr or 0)
            if use_multicast is None:
                use_multicast = os.environ.get("R2D2_DP_NO_MULTICAST") != "1"
            self.multicast = bool(use_multicast and mc)
            gp = (C.c_ulonglong * self.world)(*[int(x) for x in self._hg.buffer_ptrs])
            cp = (C.c_ulonglong * self.world)(*[int(x) for x in self._hc.buffer_ptrs])
            h = C.c_void_p()
            _lib.check(lib.r2d2_dp_create(self.rank, self.world, gp, mc, cp, C.byref(h)))
        self._h = h
        self.device = dev

    def allreduce(self, off: int, length: int, channel: int, ctas: int, threads: int, with_rows: bool = False) -> None:
        """In-place SUM over ranks of grads[off, off+length) on the current stream; with_rows also turns the row count in
        the padding slot into learner.grad_scale = 1 / (global rows)."""
        from . import _lib
        lrn = self.learner
        slot = _rows_slot(lrn)
        _lib.check(_lib.lib().r2d2_dp_allreduce(self._h, off, length, channel, _lib.ptr(lrn.rows) if with_rows else None,
                                                slot if with_rows else 0, _lib.ptr(lrn.grad_scale) if with_rows else None,
                                                ctas, threads, 1 if self.multicast else 0,
                                                torch.cuda.current_stream(self.device).cuda_stream))

    def is_post(self, replay, idx) -> None:
        """Right after sampling: this rank's min(p_i)/root of the sampled leaves goes to every peer (nobody waits)."""
        from . import _lib
        tree = replay.tree
        leaf_base = (1 << (tree.num_layers - 1)) - 1
        _lib.check(_lib.lib().r2d2_dp_is_post(self._h, tree.nodes_device().data_ptr(), leaf_base, _lib.ptr(idx), idx.numel(),
                                              torch.cuda.current_stream(self.device).cuda_stream))

    def is_apply(self, is_weights: torch.Tensor, beta: float) -> None:
        """Before K2 of the same update: is_weights *= global_is_factor of this rank, in place on the current stream."""
        from . import _lib
        _lib.check(_lib.lib().r2d2_dp_is_apply(self._h, float(beta), _lib.ptr(is_weights), is_weights.numel(), None,
                                               torch.cuda.current_stream(self.device).cuda_stream))

    def error(self) -> int:
        import ctypes as C
        from . import _lib
        out = C.c_uint(0)
        _lib.check(_lib.lib().r2d2_dp_error(self._h, C.byref(out)))
        return out.value

    def close(self) -> None:
        from . import _lib
        if self._h is not None:
            _lib.lib().r2d2_dp_destroy(self._h)
            self._h = None


class PeerISWeights:
    """Learner.is_weight_sync for PeerExchange.  `correct` (after sampling) posts this rank's scalar to the peers; the learner's
    pre_td_hook (`apply`, a graph-capturable launch right before K2) turns the ranks' scalars into this rank's factor and scales
    batch['is_weights'] -- a millisecond after the post, so no rank waits for another one."""

    ahead_safe = True          # post(i+1) may run before apply(i): Learner.update_from_replay's sample-ahead pipeline

    def __init__(self, exchange: PeerExchange, beta: float):
        self.exchange, self.beta = exchange, beta
        self._weights = None
        self.apply.__func__.capturable = True
        dev = exchange.device
        self.side = torch.cuda.Stream(device=dev, priority=-1)   # the post's system-scope release took 45 us behind the 300 MB gather
        self._sampled, self._posted = torch.cuda.Event(), torch.cuda.Event()
        self._idx = None

    def correct(self, replay, batch, idx) -> None:
        main = torch.cuda.current_stream(self.exchange.device)
        self._sampled.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self._sampled)
            self.exchange.is_post(replay, idx)
            self._posted.record(self.side)
        self._idx = idx                                           # keeps the index tensor alive until `join`
        self._weights = batch["is_weights"]

    def join(self) -> None:
        """Before the tree / index buffers the post reads are modified again (priority update): long finished by then."""
        torch.cuda.current_stream(self.exchange.device).wait_event(self._posted)
        self._idx = None

    def apply(self, learner) -> None:
        # only for the batch `correct` posted for (every post needs exactly one apply and vice versa): batches that did not
        # come from this rank's replay shard (reference-format tuples from the host) carry their sampler's own weights
        b = learner._live
        if self._weights is None or b is None or b.get("is_weights") is not self._weights:
            return
        self.exchange.is_apply(self._weights, self.beta)


def make_peer_grad_hook(learner, exchange: PeerExchange, dense_ctas: Optional[int] = None, conv_ctas: Optional[int] = None):
    """grad_hook on PeerExchange: the FC/LSTM/head range (98 % of the bytes, with the row count) is reduced from a side stream
    as soon as the backward pass has finished it -- a few small CTAs running next to the conv layers' backward kernels --
    and the 0.3 MB conv range on the learner's stream after the backward pass.  Two launches, no NCCL, no torch ops."""
    from . import _lib
    from .learner_core import PARAM_NAMES
    dev = learner.device
    side = torch.cuda.Stream(device=dev, priority=-1)
    ev, done = torch.cuda.Event(), torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))                 # materialise the cudaEvent_t
    _lib.check(_lib.lib().r2d2_net_set_dense_grads_event(learner._h, ev.cuda_event))
    n = learner.grads.flat.numel()
    dense_off = learner.grads.offsets[PARAM_NAMES.index("feature.7.weight")]
    slot = _rows_slot(learner)
    assert slot is not None and slot >= dense_off and dense_off % 4 == 0 and n % 4 == 0
    # dense range: spread thin (one warp per SM) so that no SM of the concurrent backward kernels is slowed much; conv range:
    # nothing else is running, latency matters
    dense_ctas = dense_ctas or int(os.environ.get("R2D2_DP_DENSE_CTAS", "148"))
    dense_threads = int(os.environ.get("R2D2_DP_DENSE_THREADS", "32"))
    conv_ctas = conv_ctas or int(os.environ.get("R2D2_DP_CONV_CTAS", "8"))
    keep = {"event": ev, "stream": side}                      # owned by the hook: the library only borrows the event

    def hook(lrn):
        assert lrn is learner and keep
        main = torch.cuda.current_stream(dev)
        with torch.cuda.stream(side):
            side.wait_event(ev)                               # recorded inside the backward call / graph that was just launched
            exchange.allreduce(dense_off, n - dense_off, 0, dense_ctas, dense_threads, with_rows=True)
            done.record(side)
        exchange.allreduce(0, dense_off, 1, conv_ctas, 256)
        main.wait_event(done)
    hook.capturable = False
    hook.split_graph = True
    return hook


def global_is_factor(local_min_over_root: torch.Tensor, beta: float, group=None) -> torch.Tensor:
    """Correction that turns shard-local importance weights into the weights of ONE prioritized sampler over all shards.

    Rank s samples its part of the global batch with probability p_i / root_s from its own tree and normalises by its
    own batch minimum: w_i = (p_i / min_s)^-beta (priority_tree.py:39-41).  Sampled jointly, the weight of i would be
    ((p_i / root_s) / m)^-beta with m = min over the GLOBAL batch of p_j / root_s(j).  The two differ by the per-rank
    scalar ((min_s / root_s) / m)^-beta <= 1, which needs one MIN all-reduce of one float64."""
    m = local_min_over_root.detach().clone()
    dist.all_reduce(m, op=dist.ReduceOp.MIN, group=group)
    return torch.pow(local_min_over_root / m, -float(beta))


class GlobalISWeights:
    """Computes this rank's global_is_factor for a sampled batch on a side stream (tiny kernels + one scalar MIN all-reduce,
    overlapping the forward unroll).  The loss is linear in the importance weights and the factor is one scalar per rank, so
    instead of rescaling batch['is_weights'] before K2 the gradient hooks multiply this rank's gradient of loss_sum by it
    right before the reduction -- same global-mean gradient, nothing on the learner's critical path.  (TD errors and
    priorities do not depend on the weights; the reported loss stays in local-weight units.)"""

    def __init__(self, device, beta: float, group=None):
        self.beta, self.group = beta, group
        self.device = device
        self.side = torch.cuda.Stream(device=device)
        self.ev = torch.cuda.Event()
        self.factor = torch.ones(1, dtype=torch.float32, device=device)
        self.ev.record(torch.cuda.current_stream(device))

    def correct(self, replay, batch, idx) -> None:
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.side):
            self.side.wait_stream(main)                       # after the sample kernels of this update (and the previous hook's reads of `factor`)
            nodes = replay.tree.nodes_device()
            leaf_base = (1 << (replay.tree.num_layers - 1)) - 1
            m = (nodes[leaf_base + idx].min() / nodes[0]).reshape(1)
            self.factor.copy_(global_is_factor(m, self.beta, self.group))
            self.ev.record(self.side)

    def wait(self, learner=None) -> None:
        torch.cuda.current_stream(self.device).wait_event(self.ev)


def broadcast_parameters(learner, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s online/target parameters and optimizer state."""
    for t in (learner.online.flat, learner.target.flat, learner.exp_avg, learner.exp_avg_sq):
        dist.broadcast(t, src=src, group=group)
    learner.pack(0)
    learner.pack(1)
