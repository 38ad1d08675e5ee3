"""Device-resident learner core: flat parameter buffers + the K1/K1b/K2/K5 pipeline.

This is the engine under ``worker.Learner`` (worker.py:278-381 of the reference):
one update = unroll(online) + unroll(target) -> fused TD -> BPTT -> [gradient
all-reduce hook] -> clip+Adam -> re-pack, all launched on the current CUDA
stream through the C ABI, no host synchronisation inside.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional

import torch

from . import _lib

PARAM_NAMES = [
    "feature.0.weight", "feature.0.bias", "feature.2.weight", "feature.2.bias", "feature.4.weight", "feature.4.bias",
    "feature.7.weight", "feature.7.bias", "recurrent.weight_ih_l0", "recurrent.weight_hh_l0", "recurrent.bias_ih_l0",
    "recurrent.bias_hh_l0", "advantage.0.weight", "advantage.0.bias", "advantage.2.weight", "advantage.2.bias",
    "value.0.weight", "value.0.bias", "value.2.weight", "value.2.bias",
]
HIDDEN = 512
_BATCH_KEYS = ("obs", "last_action", "last_reward", "hidden", "action", "n_step_reward", "gamma", "burn_in", "learning", "forward", "is_weights")


def param_shapes(action_dim: int, in_channels: int = 1):
    """model.py:39-63 (conv1 in-channels generalised to C)."""
    A, Cc, H = action_dim, in_channels, HIDDEN
    return {
        "feature.0.weight": (32, Cc, 8, 8), "feature.0.bias": (32,), "feature.2.weight": (64, 32, 4, 4),
        "feature.2.bias": (64,), "feature.4.weight": (64, 64, 3, 3), "feature.4.bias": (64,),
        "feature.7.weight": (512, 3136), "feature.7.bias": (512,),
        "recurrent.weight_ih_l0": (4 * H, 512 + A + 1), "recurrent.weight_hh_l0": (4 * H, H),
        "recurrent.bias_ih_l0": (4 * H,), "recurrent.bias_hh_l0": (4 * H,),
        "advantage.0.weight": (H, H), "advantage.0.bias": (H,), "advantage.2.weight": (A, H), "advantage.2.bias": (A,),
        "value.0.weight": (H, H), "value.0.bias": (H,), "value.2.weight": (1, H), "value.2.bias": (1,),
    }


def param_offsets(action_dim: int, in_channels: int = 1):
    off = (C.c_int64 * 21)()
    _lib.check(_lib.lib().r2d2_net_param_layout(action_dim, in_channels, off))
    return list(off)


class FlatParams:
    """One flat fp32 device buffer + per-tensor views named like the reference state_dict."""

    def __init__(self, action_dim: int, in_channels: int, device, flat: Optional[torch.Tensor] = None):
        self.offsets = param_offsets(action_dim, in_channels)
        self.shapes = param_shapes(action_dim, in_channels)
        if flat is None:
            flat = torch.zeros(self.offsets[-1], dtype=torch.float32, device=device)
        assert flat.dtype == torch.float32 and flat.numel() == self.offsets[-1] and flat.is_contiguous()
        self.flat = flat
        self.views: Dict[str, torch.Tensor] = {}
        for i, name in enumerate(PARAM_NAMES):
            shape = self.shapes[name]
            n = 1
            for s in shape:
                n *= s
            self.views[name] = self.flat[self.offsets[i]:self.offsets[i] + n].view(shape)

    def load(self, state_dict) -> None:
        for name, v in self.views.items():
            v.copy_(state_dict[name].to(device=v.device, dtype=torch.float32))

    def state_dict(self, device=None):
        return {k: (v.detach().clone() if device is None else v.detach().to(device)) for k, v in self.views.items()}


class WeightPublisher:
    """Asynchronous publication of the online parameters to a host-side (shared-memory) model -- the reference's
    `shared_model.load_state_dict(...)` every 4 updates (worker.py:306-307,372-373), off the learner's critical path:
    a device snapshot of the flat buffer, one 17 MB device->pinned-host copy of it on a side stream, then a daemon thread waits for it and copies
    the host views into the shared model.  A publication that is still in flight is skipped, never queued (actors only
    ever want the newest weights)."""

    def __init__(self, params: FlatParams, shared_model):
        import queue
        import threading
        self.params = params
        self.shared = shared_model
        self.stream = torch.cuda.Stream(device=params.flat.device)
        self.host = torch.empty(params.flat.numel(), dtype=torch.float32).pin_memory()
        self.snap = torch.empty_like(params.flat)                   # device snapshot: later optimizer steps may run during the D2H
        self.host_views = {}
        for i, name in enumerate(PARAM_NAMES):
            shape = params.shapes[name]
            n = 1
            for d in shape:
                n *= d
            self.host_views[name] = self.host[params.offsets[i]:params.offsets[i] + n].view(shape)
        self._done = torch.cuda.Event()
        self._busy = threading.Event()
        self._q = queue.Queue()
        self.published = 0
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def publish(self) -> bool:
        """Enqueue one publication of the CURRENT parameters; False if the previous one is still in flight."""
        if self._busy.is_set():
            return False
        self._busy.set()
        self.snap.copy_(self.params.flat)                                       # 17 MB device copy on the launching stream
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.params.flat.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.host.copy_(self.snap, non_blocking=True)
            self._done.record(self.stream)
        self._q.put(1)
        return True

    def _run(self):
        while True:
            self._q.get()
            self._done.synchronize()
            with torch.no_grad():
                self.shared.load_state_dict(self.host_views)
            self.published += 1
            self._busy.clear()

    def wait(self, timeout: float = 10.0) -> None:
        import time
        t0 = time.time()
        while self._busy.is_set() and time.time() - t0 < timeout:
            time.sleep(0.001)


class BatchStager:
    """Double-buffered host -> device staging of reference-format batches on a copy stream.

    The reference learner prefetches batches from its queue in a background thread (worker.py:309-316) and then
    moves each one to the device synchronously (worker.py:331-334).  Here the move of batch k+1 (pinned host memory,
    one cudaMemcpyAsync per field into preallocated device slots) runs on a side stream while batch k is being
    trained on; events hand slots back and forth between the two streams."""

    def __init__(self, core: "DeviceLearner", slots: int = 2):
        d, B, T, A, R = core.device, core.B, core.T, core.A, core.rows_cap
        self.core = core
        # high priority: its copies are dispatched as soon as their slot is free instead of queueing behind the graph /
        # kernels of the update that is running (measured: with a default-priority copy stream the 154 MB copy sometimes
        # serialises behind the update's CUDA graph, 4.5-5.2 ms per step instead of 3.1)
        self.stream = torch.cuda.Stream(device=d, priority=-1)
        self.slots = []
        for _ in range(slots):
            buf = dict(obs=torch.zeros(B, T, core.C, 84, 84, dtype=torch.uint8, device=d),
                       last_action=torch.zeros(B, T, A, dtype=torch.uint8, device=d), last_reward=torch.zeros(B, T, device=d),
                       hidden=torch.zeros(B, 2, HIDDEN, device=d), action=torch.zeros(R, dtype=torch.uint8, device=d),
                       n_step_reward=torch.zeros(R, device=d), gamma=torch.zeros(R, device=d),
                       burn_in=torch.zeros(B, dtype=torch.uint8, device=d), learning=torch.zeros(B, dtype=torch.uint8, device=d),
                       forward=torch.zeros(B, dtype=torch.uint8, device=d), is_weights=torch.zeros(R, device=d))
            self.slots.append(dict(buf=buf, ready=torch.cuda.Event(), free=None, tmax=T))
        # the zero fills above were enqueued on the CURRENT stream, possibly behind a running update: the copy stream must not
        # overtake them (the first staged batch would be zeroed after it had arrived)
        filled = torch.cuda.Event()
        filled.record(torch.cuda.current_stream(d))
        self.stream.wait_event(filled)
        self._next = 0

    @staticmethod
    def _as_u8(t):
        return t.view(torch.uint8) if t.dtype == torch.bool else (t if t.dtype == torch.uint8 else (t != 0).to(torch.uint8))

    @property
    def device(self):
        return self.core.device

    @_lib.on_device
    def stage(self, fields: dict) -> int:
        """Enqueue the copies of one batch (dict with the 14-tuple's learner-visible fields); returns a slot handle."""
        i = self._next
        self._next = (i + 1) % len(self.slots)
        slot = self.slots[i]
        buf, core = slot["buf"], self.core
        with torch.cuda.stream(self.stream):
            if slot["free"] is not None:
                self.stream.wait_event(slot["free"])            # the update that last used this slot has finished
            Tb = fields["obs"].shape[1]
            assert Tb <= core.T
            buf["obs"][:, :Tb].copy_(fields["obs"], non_blocking=True)
            buf["last_action"][:, :Tb].copy_(self._as_u8(fields["last_action"]), non_blocking=True)
            buf["last_reward"][:, :Tb].copy_(fields["last_reward"], non_blocking=True)
            if Tb < slot["tmax"] or Tb < core.T:                  # pad_sequence-style zero padding at the END of time
                buf["obs"][:, Tb:].zero_(); buf["last_action"][:, Tb:].zero_(); buf["last_reward"][:, Tb:].zero_()
            slot["tmax"] = Tb
            h = fields["hidden"]
            if h.shape[0] == 2 and h.shape[1] == core.B and h.shape[0] != core.B:
                h = h.transpose(0, 1)                              # (2,B,H) view of the stacked (B,2,H) array, worker.py:223
            buf["hidden"].copy_(h, non_blocking=True)
            rows = fields["action"].numel()
            buf["action"][:rows].copy_(fields["action"].reshape(-1), non_blocking=True)
            buf["n_step_reward"][:rows].copy_(fields["n_step_reward"], non_blocking=True)
            buf["gamma"][:rows].copy_(fields["gamma"], non_blocking=True)
            buf["is_weights"][:rows].copy_(fields["is_weights"], non_blocking=True)
            for k in ("burn_in", "learning", "forward"):
                buf[k].copy_(fields[k], non_blocking=True)
            slot["ready"].record(self.stream)
        return i

    def acquire(self, handle: int) -> dict:
        torch.cuda.current_stream(self.core.device).wait_event(self.slots[handle]["ready"])
        return self.slots[handle]["buf"]

    def release(self, handle: int) -> None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.core.device))
        self.slots[handle]["free"] = ev


class DeviceLearner:
    def __init__(self, action_dim: int, batch_size: int, seq_frames: int, in_channels: int = 1, max_learning: int = 40,
                 max_forward: int = 5, lr: float = 1e-4, eps: float = 1e-3, grad_norm: float = 40.0,
                 betas=(0.9, 0.999), device=None):
        _lib.require_device()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.A, self.B, self.T, self.C = action_dim, batch_size, seq_frames, in_channels
        self.Lmax, self.F = max_learning, max_forward
        self.lr, self.eps, self.grad_norm, self.betas = lr, eps, grad_norm, betas
        self.online = FlatParams(action_dim, in_channels, self.device)
        self.target = FlatParams(action_dim, in_channels, self.device)
        self.grads = FlatParams(action_dim, in_channels, self.device)
        self.exp_avg = torch.zeros_like(self.online.flat)
        self.exp_avg_sq = torch.zeros_like(self.online.flat)
        self._num_updates = 0
        self._s2d_idx = 0
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().r2d2_net_create(batch_size, seq_frames, in_channels, action_dim, max_learning,
                                                  max_forward, C.byref(h)))
        self._h = h
        self.rows_cap = _lib.lib().r2d2_net_rows_capacity(h)
        self.KU = _lib.lib().r2d2_net_ku(h)
        d = self.device
        R, A = self.rows_cap, action_dim
        self.q = torch.zeros(R, A, device=d)
        self.qn_online = torch.zeros(R, A, device=d)
        self.qn_target = torch.zeros(R, A, device=d)
        self.dq = torch.zeros(R, A, device=d)
        self.td = torch.zeros(R, device=d)
        self.prio = torch.zeros(batch_size, device=d)
        self.loss_sum = torch.zeros(1, device=d)
        self.rows = torch.zeros(1, dtype=torch.int32, device=d)
        self.grad_scale = torch.ones(1, device=d)
        self.norm = torch.zeros(1, device=d)
        self._norm_ws = torch.zeros(592, dtype=torch.float64, device=d)
        self._step_dev = torch.zeros(1, dtype=torch.int64, device=d)     # device mirror of num_updates (Adam bias correction)
        # CUDA-graph replay of update(): captured per set of batch buffers the second time that set is seen (persistent
        # staging slots / the replay's gather buffers); R2D2_CUDA_GRAPH=0 keeps every update eager
        import os
        self.use_graph = os.environ.get("R2D2_CUDA_GRAPH", "1") != "0"
        self._graphs, self._graph_seen = {}, {}
        # hook called between backward and the optimizer: (learner) -> None, e.g. NCCL all-reduce
        self.grad_hook: Optional[Callable[["DeviceLearner"], None]] = None
        # hook called between the forward unroll and the TD kernel (e.g. join a side stream that rescales is_weights)
        self.pre_td_hook: Optional[Callable[["DeviceLearner"], None]] = None
        # recorded after K2 (an external event-record node when the update is replayed from a graph): "priorities final" --
        # Learner's sampling stream waits on it to apply them and sample the next batch while the backward pass runs
        self.td_event = torch.cuda.Event()
        self.td_event.record(torch.cuda.current_stream(self.device))       # materialise the cudaEvent_t

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().r2d2_net_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def num_updates(self) -> int:
        return self._num_updates

    @num_updates.setter
    def num_updates(self, v: int) -> None:
        self._num_updates = int(v)
        self._step_dev.fill_(int(v))

    # ------------------------------------------------------------------ parameters
    @_lib.on_device
    def load_state_dict(self, sd, target_sd=None) -> None:
        self.online.load(sd)
        self.target.load(target_sd if target_sd is not None else sd)
        self.pack(0)
        self.pack(1)

    @_lib.on_device
    def sync_target(self) -> None:                      # worker.py:376-377
        self.target.flat.copy_(self.online.flat)
        self.pack(1)

    @_lib.on_device
    def pack(self, which: int) -> None:
        flat = self.online.flat if which == 0 else self.target.flat
        _lib.check(_lib.lib().r2d2_net_pack(self._h, which, _lib.ptr(flat), _lib.stream_ptr()))

    # ------------------------------------------------------------------ passes
    def _pad_time(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape[1] == self.T:
            return x.contiguous()
        assert x.shape[1] < self.T, "batch longer than the workspace"
        out = torch.zeros((x.shape[0], self.T) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
        out[:, :x.shape[1]] = x
        return out

    def prepare(self, batch: dict) -> dict:
        """Move/shape a batch dict (see worker.py:219-238) for the kernels: device, contiguous,
        time axis padded to T, hidden as (B,2,H), last_action as bytes."""
        d = self.device
        t = lambda v: v.to(d, non_blocking=True) if isinstance(v, torch.Tensor) else torch.as_tensor(v).to(d)
        out = dict(batch)
        out["obs"] = self._pad_time(t(batch["obs"]))
        la = t(batch["last_action"])
        if la.dtype == torch.bool:
            la = la.view(torch.uint8)
        elif la.dtype != torch.uint8:
            la = (la != 0).to(torch.uint8)
        out["last_action"] = self._pad_time(la)
        out["last_reward"] = self._pad_time(t(batch["last_reward"]).float())
        out["hidden"] = t(batch["hidden"]).float().contiguous()
        assert out["hidden"].shape == (self.B, 2, HIDDEN), out["hidden"].shape
        for k in ("burn_in", "learning", "forward"):
            out[k] = t(batch[k]).to(torch.uint8).contiguous()
        if "action" in batch:
            out["action"] = t(batch["action"]).to(torch.uint8).reshape(-1).contiguous()
            out["n_step_reward"] = t(batch["n_step_reward"]).float().contiguous()
            out["gamma"] = t(batch["gamma"]).float().contiguous()
            out["is_weights"] = t(batch["is_weights"]).float().contiguous()
        return out

    @_lib.on_device
    @_lib.on_device
    def select_s2d(self, idx: int) -> None:
        """Which of the two frame staging buffers the following forward / backward calls read (the other one may be filled
        with the next batch meanwhile: DeviceReplay.sample(fuse_into=self, slot=...))."""
        if idx != self._s2d_idx:
            _lib.check(_lib.lib().r2d2_net_select_s2d(self._h, int(idx)))
            self._s2d_idx = int(idx)

    def use_grad_buffer(self, flat: torch.Tensor) -> None:
        """Rebind the flat gradient buffer (same layout) to caller-provided device memory -- e.g. symmetric memory that the
        other ranks of a data-parallel run have mapped (dist.PeerExchange).  Call before the first update."""
        assert not self._graphs, "rebind the gradient buffer before any update has been captured"
        assert flat.device == self.device
        flat.zero_()
        self.grads = FlatParams(self.A, self.C, self.device, flat=flat)

    def forward(self, which: int, b: dict, q_learn: Optional[torch.Tensor], q_shift: Optional[torch.Tensor]) -> None:
        flat = self.online.flat if which == 0 else self.target.flat
        p = _lib.ptr
        assert b["obs"].shape == (self.B, self.T, self.C, 84, 84) and b["obs"].dtype == torch.uint8
        _lib.check(_lib.lib().r2d2_net_forward(self._h, which, p(flat), p(b["obs"]), p(b["last_action"]),
                                               p(b["last_reward"]), p(b["hidden"]), p(b["burn_in"]), p(b["learning"]),
                                               p(b["forward"]), p(q_learn), p(q_shift), _lib.stream_ptr()))

    @_lib.on_device
    def backward(self, dq: torch.Tensor) -> None:
        p = _lib.ptr
        _lib.check(_lib.lib().r2d2_net_backward(self._h, p(self.online.flat), p(dq), p(self.grads.flat),
                                                _lib.stream_ptr()))

    @_lib.on_device
    def compute_forward(self, b: dict) -> None:
        """worker.py:345-357: the three Q tensors, then TD / loss / priorities / dLoss/dQ (K1 + K2)."""
        self._live = b                                   # keep obs/hidden alive until backward ran
        p = _lib.ptr
        _lib.check(_lib.lib().r2d2_net_forward_pair(self._h, p(self.online.flat), p(self.target.flat), p(b["obs"]),
                                                    p(b["last_action"]), p(b["last_reward"]), p(b["hidden"]),
                                                    p(b["burn_in"]), p(b["learning"]), p(b["forward"]), p(self.q),
                                                    p(self.qn_online), p(self.qn_target), _lib.stream_ptr()))
        if self.pre_td_hook is not None:
            self.pre_td_hook(self)
        _lib.check(_lib.lib().r2d2_td_loss(p(self.q), p(self.qn_online), p(self.qn_target), p(b["action"]),
                                           p(b["n_step_reward"]), p(b["gamma"]), p(b["is_weights"]), p(b["learning"]),
                                           self.B, self.A, p(self.td), p(self.prio), p(self.loss_sum), p(self.rows),
                                           p(self.dq), _lib.stream_ptr()))
        if self.td_event is not None:
            _lib.check(_lib.lib().r2d2_event_record(self.td_event.cuda_event, _lib.stream_ptr()))

    @_lib.on_device
    def compute_gradients(self, b: dict) -> None:
        """worker.py:345-363: Q passes, TD/loss/priorities, backward.  Results stay on device in
        self.td / self.prio / self.loss_sum / self.rows / self.grads (grads of loss_sum)."""
        self.compute_forward(b)
        self.backward(self.dq)

    @_lib.on_device
    def apply_gradients(self, _count: bool = True) -> None:
        """worker.py:364-365 (+ re-pack of the online weights).  The mean over rows of worker.py:354 is applied here: with
        a grad_hook (data parallel) the hook leaves 1/global_rows in self.grad_scale, otherwise the kernel divides by
        this batch's device-side row count."""
        if _count:
            self._num_updates += 1
        p = _lib.ptr                                     # (the device mirror of the update count is incremented by the norm kernel)
        b1, b2 = self.betas
        hooked = self.grad_hook is not None
        _lib.check(_lib.lib().r2d2_clip_adam_step(p(self.online.flat), p(self.grads.flat), p(self.exp_avg), p(self.exp_avg_sq),
                                                 self.online.flat.numel(), p(self.grad_scale) if hooked else None,
                                                 None if hooked else p(self.rows), p(self._norm_ws), float(self.grad_norm),
                                                 float(self.lr), float(b1), float(b2), float(self.eps), p(self._step_dev),
                                                 p(self.norm), _lib.stream_ptr()))
        self.pack(0)

    def _update_eager(self, b: dict, _count: bool = True) -> None:
        self.compute_gradients(b)
        if self.grad_hook is not None:
            self.grad_hook(self)
        self.apply_gradients(_count)

    def _graph_for(self, key, fn):
        """CUDA graph of `fn` for this key, captured the second time the key is seen; None while still eager."""
        g = self._graphs.get(key)
        if g is not None:
            return g
        if len(self._graph_seen) > 256:
            self._graph_seen.clear()
        seen = self._graph_seen.get(key, 0) + 1
        self._graph_seen[key] = seen
        if seen < 2:
            return None                                       # first sight: eager (also sets per-function attributes, not capturable)
        if len(self._graphs) >= 16:
            self._graphs.clear()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g):
                fn()
        except Exception as e:                                # stay eager from now on
            import warnings
            warnings.warn(f"CUDA-graph capture of the learner update failed ({e}); continuing with eager launches")
            self.use_graph = False
            torch.cuda.synchronize(self.device)
            return None
        self._graphs[key] = g
        return g

    @_lib.on_device
    def update(self, b: dict) -> None:
        """One learner update on prepared device buffers.  Nothing in the launch sequence depends on host values that change
        between updates (update count and row count live on the device), so for a recurring set of buffers the ~45 launches
        are captured once in a CUDA graph and replayed.  With a data-parallel grad_hook that declares `split_graph` the
        update is two graphs -- gradients, optimizer -- with the (eager) collective between them."""
        hook = self.grad_hook
        split = hook is not None and getattr(hook, "split_graph", False)
        pre = self.pre_td_hook                                   # plain stream-ordered launches may declare themselves capturable
        if not self.use_graph or (pre is not None and not getattr(pre, "capturable", False)) or (hook is not None and not split):
            return self._update_eager(b)
        key = (_lib.lib().r2d2_config_epoch(), self._s2d_idx) + tuple(None if v is None else (v.data_ptr() if isinstance(v, torch.Tensor) else v)
                                                      for v in (b.get(k) for k in _BATCH_KEYS))
        self._live = b
        if not split:
            g = self._graph_for(key, lambda: self._update_eager(b, _count=False))
            if g is None:
                return self._update_eager(b)
            self._num_updates += 1
            g.replay()
            return
        ga = self._graph_for(key + ("grad",), lambda: self.compute_gradients(b))
        if ga is None:
            self.compute_gradients(b)
        else:
            ga.replay()
        hook(self)
        gb = self._graph_for(key[:1] + ("apply",), lambda: self.apply_gradients(_count=False))
        if gb is None:
            self.apply_gradients()
        else:
            self._num_updates += 1
            gb.replay()

    # ------------------------------------------------------------------ debug
    def debug_split(self, which: int, name: str, numel: int) -> torch.Tensor:
        """fp32 value (hi + lo) of a split bf16 tensor of the workspace."""
        hi = self.debug_tensor(which, name + ".hi", numel, torch.bfloat16)
        lo = self.debug_tensor(which, name + ".lo", numel, torch.bfloat16)
        return hi.float() + lo.float()

    def debug_tensor(self, which: int, name: str, numel: int, dtype=torch.float32) -> torch.Tensor:
        from .priority_tree import _from_device_ptr
        addr = _lib.lib().r2d2_net_debug_ptr(self._h, which, name.encode())
        assert addr, name
        return _from_device_ptr(addr, numel, dtype, self.device)
