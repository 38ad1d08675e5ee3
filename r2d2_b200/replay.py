"""HBM-resident prioritized sequence replay: K4 block store + K3 sum tree.

Device-side engine under ``worker.ReplayBuffer`` (worker.py:38-261 of the reference).  Blocks
arrive from actors as host ``Block`` objects (worker.py:23-35); ``add`` packs one into a pinned
staging slot and enqueues ONE async H2D copy on a side stream, writes the block's slot priorities
into the GPU sum tree and keeps the reference's bookkeeping (size, env_steps, block_ptr).
``sample`` = tree sample + gather, entirely on the device.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .priority_tree import PriorityTree

_FIELDS = ["obs", "last_action", "last_reward", "action", "n_step_reward", "gamma", "hidden", "burn", "learn", "fwd",
           "num_seq", "total"]


class DeviceReplay:
    def __init__(self, buffer_capacity: int, block_length: int, burn_in_steps: int, learning_steps: int, forward_steps: int,
                 action_dim: int, obs_shape=(1, 84, 84), hidden_dim: int = 512, alpha: float = 0.9, beta: float = 0.6,
                 batch_size: int = 64, device=None, seed: int = 0, staging_slots: int = 6,
                 tree_capacity: Optional[int] = None):
        _lib.require_device()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.block_len, self.burn_in, self.learning, self.forward = block_length, burn_in_steps, learning_steps, forward_steps
        self.A, self.C, self.H = action_dim, int(obs_shape[0]), hidden_dim
        assert tuple(obs_shape[1:]) == (84, 84)
        self.num_blocks = buffer_capacity // block_length                 # worker.py:47
        self.seq_per_block = block_length // learning_steps               # worker.py:48
        self.num_sequences = buffer_capacity // learning_steps            # worker.py:45
        self.batch_size = batch_size
        self.T = burn_in_steps + learning_steps + forward_steps
        self.rows_cap = (batch_size * learning_steps + 7) // 8 * 8
        self.tree = PriorityTree(max(self.num_sequences, tree_capacity or 0), alpha, beta, device=self.device, seed=seed)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().r2d2_replay_create(self.num_blocks, block_length, burn_in_steps, learning_steps, forward_steps,
                                                     self.C, action_dim, hidden_dim, C.byref(h)))
        self._h = h
        off = (C.c_int64 * 12)()
        _lib.check(_lib.lib().r2d2_replay_layout(h, off))
        self.layout = dict(zip(_FIELDS, list(off)))
        self.blob_bytes = self.layout["total"]
        # pinned staging ring + side stream for actor -> HBM ingest
        self._staging = [torch.zeros(self.blob_bytes, dtype=torch.uint8).pin_memory() for _ in range(staging_slots)]
        self._staging_events = [None] * staging_slots
        self._slot = 0
        import threading
        self._stage_lock = threading.Lock()
        self._slot_free = [threading.Event() for _ in range(staging_slots)]
        for e in self._slot_free:
            e.set()
        self._pending = []                      # committed blocks whose priorities have not entered the tree yet (commit(defer=True))
        self._gather_event = None               # recorded after the most recent gather: an ingest copy may not overtake it
        self.ingested_bytes = 0
        self.ingest_stream = torch.cuda.Stream(device=self.device, priority=-1)   # copies must not queue behind a running update
        # reference bookkeeping (worker.py:50-68)
        self.block_ptr = 0
        self.size = 0
        self.env_steps = 0
        self.num_episodes = 0
        self.episode_reward = 0.0
        self._block_steps = np.zeros(self.num_blocks, dtype=np.int64)
        self._alloc_batch()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().r2d2_replay_destroy(h)
            except Exception:
                pass
            self._h = None

    def __len__(self):
        return self.size

    def _alloc_batch(self):
        self.batches = [self._new_batch()]          # a second set is created on first use (sample(slot=1): sampling ahead)
        self.batch = self.batches[0]

    def _new_batch(self):
        d, B, T, A = self.device, self.batch_size, self.T, self.A
        R = self.rows_cap
        return dict(
            obs=torch.zeros(B, T, self.C, 84, 84, dtype=torch.uint8, device=d),
            last_action=torch.zeros(B, T, A, dtype=torch.uint8, device=d),
            last_reward=torch.zeros(B, T, device=d), hidden=torch.zeros(B, 2, self.H, device=d),
            action=torch.zeros(R, dtype=torch.uint8, device=d), n_step_reward=torch.zeros(R, device=d),
            gamma=torch.zeros(R, device=d), burn_in=torch.zeros(B, dtype=torch.uint8, device=d),
            learning=torch.zeros(B, dtype=torch.uint8, device=d), forward=torch.zeros(B, dtype=torch.uint8, device=d),
            is_weights=torch.zeros(R, device=d), rows=torch.zeros(1, dtype=torch.int32, device=d))

    # ------------------------------------------------------------------ ReplayBuffer.add (worker.py:141-161)
    def pack(self, block, out: np.ndarray) -> int:
        """Pack a Block into the slot layout; returns the number of leading bytes that must travel."""
        L = self.layout
        n_frames = block.obs.shape[0]
        fb = self.C * 84 * 84
        out[L["obs"]:L["obs"] + n_frames * fb] = block.obs.reshape(-1)
        la = np.ascontiguousarray(block.last_action).view(np.uint8).reshape(-1)
        out[L["last_action"]:L["last_action"] + la.size] = la

        def put(name, arr, dtype):
            a = np.ascontiguousarray(arr, dtype=dtype).view(np.uint8).reshape(-1)
            out[L[name]:L[name] + a.size] = a
        put("last_reward", block.last_reward, np.float32)
        put("action", block.action, np.uint8)
        put("n_step_reward", block.n_step_reward, np.float32)
        put("gamma", block.gamma, np.float32)
        put("hidden", block.hidden, np.float32)
        for name, arr in (("burn", block.burn_in_steps), ("learn", block.learning_steps), ("fwd", block.forward_steps)):
            out[L[name]:L[name] + self.seq_per_block] = 0
            put(name, arr, np.uint8)
        put("num_seq", np.array([block.num_sequences], dtype=np.int32), np.int32)
        return self.blob_bytes

    def stage(self, block):
        """Host half of an insertion: pack `block` into a free pinned staging slot.  Thread-safe and CUDA-free apart from
        waiting for the slot's previous copy, so a prefetch thread can run it while the learner trains (12.5 MB of memcpy
        per 4-channel block).  Returns a handle for commit()."""
        with self._stage_lock:
            slot = self._slot
            self._slot = (slot + 1) % len(self._staging)
        self._slot_free[slot].wait()              # staged but not yet committed by the consumer
        self._slot_free[slot].clear()
        ev = self._staging_events[slot]
        if ev is not None:
            ev.synchronize()                      # the previous copy out of this pinned slot has completed
        nbytes = self.pack(block, self._staging[slot].numpy())
        return slot, nbytes, int(np.sum(block.learning_steps, dtype=np.int64))

    @_lib.on_device
    def commit(self, handle, priority: np.ndarray, episode_reward: Optional[float] = None, defer: bool = False) -> None:
        """Device half: ONE async H2D copy of the staged blob on the ingest stream (ordered after the last gather, which may
        still read the ring slot being overwritten) + the reference's bookkeeping (worker.py:141-161).  The slot priorities
        enter the sum tree, and the launching stream joins the copy, either now or -- defer=True -- at the next sample(),
        so that the copy overlaps the update that is running."""
        slot, nbytes, steps = handle
        with torch.cuda.stream(self.ingest_stream):
            if self._gather_event is not None:
                self.ingest_stream.wait_event(self._gather_event)
            _lib.check(_lib.lib().r2d2_replay_ingest(self._h, self.block_ptr, self._staging[slot].data_ptr(), nbytes,
                                                     self.ingest_stream.cuda_stream))
            done = torch.cuda.Event()
            done.record(self.ingest_stream)
        self._staging_events[slot] = done
        self._slot_free[slot].set()
        idxes = np.arange(self.block_ptr * self.seq_per_block, (self.block_ptr + 1) * self.seq_per_block, dtype=np.int64)
        self._pending.append((done, idxes, np.asarray(priority, dtype=np.float32)))
        self.ingested_bytes += nbytes
        self.size += steps - int(self._block_steps[self.block_ptr])
        self._block_steps[self.block_ptr] = steps
        self.env_steps += steps
        self.block_ptr = (self.block_ptr + 1) % self.num_blocks
        if episode_reward:
            self.episode_reward += episode_reward
            self.num_episodes += 1
        if not defer:
            self._activate_pending()

    def _activate_pending(self) -> None:
        main = torch.cuda.current_stream(self.device)
        for done, idxes, prio in self._pending:
            main.wait_event(done)                 # later samples see the block
            self.tree.update(idxes, prio)
        self._pending.clear()

    def add(self, block, priority: np.ndarray, episode_reward: Optional[float] = None) -> None:
        """ReplayBuffer.add (worker.py:141-161): stage + commit, priorities visible immediately."""
        self.commit(self.stage(block), priority, episode_reward)

    # ------------------------------------------------------------------ sample_batch (worker.py:163-240), on device
    @_lib.on_device
    def sample(self, unit_uniforms: Optional[torch.Tensor] = None, fuse_into=None, slot: int = 0):
        """Returns (batch dict of device tensors, idxes int64 device, old_ptr).  With fuse_into = a DeviceLearner whose
        shape matches, frames are written straight into its space-to-depth staging buffer and batch["obs"] is None.
        slot 0/1: which of two sets of output buffers (and of the learner's two staging buffers) receives the batch, so that
        batch i+1 can be gathered while update i still reads batch i."""
        self._activate_pending()
        idx, isw = self.tree.sample_device(self.batch_size, unit_uniforms)
        while len(self.batches) <= slot:
            self.batches.append(self._new_batch())
        out = self.gather_fused(idx, isw, fuse_into, slot) if fuse_into is not None else self.gather(idx, isw, slot)
        if self._gather_event is None:
            self._gather_event = torch.cuda.Event()
        self._gather_event.record(torch.cuda.current_stream(self.device))
        return out, idx, self.block_ptr

    def set_copy_smem(self, nbytes: int) -> None:
        """Shared-memory footprint of the gather's copy CTAs (placement control, r2d2_replay_set_copy_smem)."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().r2d2_replay_set_copy_smem(self._h, int(nbytes)))

    def gather_fused(self, idx: torch.Tensor, isw: torch.Tensor, core, slot: int = 0) -> dict:
        assert core.B == self.batch_size and core.T == self.T and core.C == self.C
        b, p = self.batches[slot], _lib.ptr
        s2d = _lib.lib().r2d2_net_s2d_buffer_at(core._h, slot)
        assert s2d, "no staging buffer"
        _lib.check(_lib.lib().r2d2_replay_gather_s2d(self._h, p(idx), p(isw), self.batch_size, self.T, s2d, p(b["last_action"]),
                                                     p(b["last_reward"]), p(b["hidden"]), p(b["action"]), p(b["n_step_reward"]),
                                                     p(b["gamma"]), p(b["burn_in"]), p(b["learning"]), p(b["forward"]),
                                                     p(b["is_weights"]), p(b["rows"]), _lib.stream_ptr()))
        out = dict(b)
        out["obs"] = None
        return out

    def gather(self, idx: torch.Tensor, isw: torch.Tensor, slot: int = 0) -> dict:
        b, p = self.batches[slot], _lib.ptr
        _lib.check(_lib.lib().r2d2_replay_gather(self._h, p(idx), p(isw), self.batch_size, self.T, p(b["obs"]), p(b["last_action"]),
                                                 p(b["last_reward"]), p(b["hidden"]), p(b["action"]), p(b["n_step_reward"]),
                                                 p(b["gamma"]), p(b["burn_in"]), p(b["learning"]), p(b["forward"]),
                                                 p(b["is_weights"]), p(b["rows"]), _lib.stream_ptr()))
        return b

    # ------------------------------------------------------------------ update_priorities (worker.py:242-261)
    @_lib.on_device
    def update_priorities(self, idxes: torch.Tensor, priorities: torch.Tensor, old_ptr: int) -> None:
        self.tree.update_device(idxes, priorities, old_ptr=old_ptr, cur_ptr=self.block_ptr, seq_per_block=self.seq_per_block)
