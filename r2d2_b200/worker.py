"""Replay buffer, learner and actor -- the reference's ``worker`` module surface
(worker.py:1-575 upstream) on top of the B200 kernels.

Same class names, constructor signatures, queue protocol and ``run()`` entry points as the
reference, so its ``train.py`` runs unmodified against this module:

* ``Learner``       owns the CUDA learner core (K1/K1b/K2/K5) AND, in the drop-in topology, the
                    HBM-resident replay (K3/K4).  ``run()`` consumes either reference-format
                    14-tuples (worker.py:219-238: a CPU sampler feeding host batches) or block
                    messages forwarded by our ``ReplayBuffer.run`` stager.
* ``ReplayBuffer``  in-process: HBM block store + GPU sum tree behind add / sample_batch /
                    update_priorities.  As a forked process (``run()``; CUDA is initialised in the
                    parent before the fork, train.py:29->36) it is a CPU stager that drains the actor
                    queues, forwards blocks to the learner and prints the reference's log lines.
* ``Actor`` / ``LocalBuffer`` / ``Block``  host-side episode accumulation and the block wire format
                    (worker.py:23-35,395-575), CPU code like upstream.
"""
import math
import os
import random
import threading
import time
from copy import deepcopy
from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch

from . import _lib, config
from .environment import create_env
from .model import AgentState, Network
from .priority_tree import PriorityTree  # noqa: F401  (re-exported like the reference module does)

############################## Replay Buffer ##############################


@dataclass
class Block:                       # worker.py:23-35
    obs: np.array
    last_action: np.array
    last_reward: np.array
    action: np.array
    n_step_reward: np.array
    gamma: np.array
    hidden: np.array
    num_sequences: int
    burn_in_steps: np.array
    learning_steps: np.array
    forward_steps: np.array


BLOCK_MSG = "r2d2_b200.block"      # tag of (tag, block, priority, episode_reward) messages on batch_queue
STAGED_MSG = "r2d2_b200.staged"    # (tag, staging handle, priority, episode_reward): a block already packed into pinned memory
STATS_MSG = "r2d2_b200.stats"      # tag of (tag, training_steps, sum_loss, size, env_steps) messages on priority_queue


class ReplayBuffer:
    def __init__(self, sample_queue_list, batch_queue, priority_queue, buffer_capacity=config.buffer_capacity,
                 sequence_len=config.block_length, alpha=config.prio_exponent, beta=config.importance_sampling_exponent,
                 batch_size=config.batch_size, action_dim=None, obs_shape=None, device=None):
        self.buffer_capacity = buffer_capacity
        self.sequence_len = config.learning_steps
        self.num_sequences = buffer_capacity // self.sequence_len
        self.block_len = config.block_length
        self.num_blocks = self.buffer_capacity // self.block_len
        self.seq_pre_block = self.block_len // self.sequence_len
        self.alpha, self.beta, self.batch_size = alpha, beta, batch_size
        self.action_dim, self.obs_shape, self.device = action_dim, obs_shape, device

        self.env_steps = 0
        self.num_episodes = 0
        self.episode_reward = 0
        self.training_steps = 0
        self.last_training_steps = 0
        self.sum_loss = 0
        self.lock = threading.Lock()
        self.size = 0
        self.last_size = 0
        self.sample_queue_list, self.batch_queue, self.priority_queue = sample_queue_list, batch_queue, priority_queue
        self._dev = None                     # DeviceReplay, created lazily (never inside a forked stager)

    # ------------------------------------------------------------------ in-process, HBM-resident mode
    def _device_replay(self, block=None):
        if self._dev is None:
            from .replay import DeviceReplay
            A = self.action_dim if self.action_dim is not None else block.last_action.shape[1]
            shape = self.obs_shape if self.obs_shape is not None else (tuple(block.obs.shape[1:]) if block is not None
                                                                        else config.obs_shape)
            self._dev = DeviceReplay(self.buffer_capacity, self.block_len, config.burn_in_steps, config.learning_steps,
                                     config.forward_steps, A, shape, config.hidden_dim, self.alpha, self.beta, self.batch_size,
                                     device=self.device)
        return self._dev

    @property
    def block_ptr(self):
        return self._dev.block_ptr if self._dev is not None else 0

    @property
    def priority_tree(self):
        return self._device_replay().tree

    def __len__(self):
        return self.size

    def add(self, block: Block, priority: np.array, episode_reward: float):          # worker.py:141-161
        with self.lock:
            dev = self._device_replay(block)
            dev.add(block, priority, episode_reward)
            self.size, self.env_steps = dev.size, dev.env_steps
            if episode_reward:
                self.episode_reward += episode_reward
                self.num_episodes += 1

    def sample_batch(self):
        '''sample one batch of training data (worker.py:163-240): the reference 14-tuple, tensors on the GPU.
        Indices come from NumPy's global legacy generator exactly like upstream, so np.random.seed reproduces them.'''
        with self.lock:
            dev = self._device_replay()
            r = torch.from_numpy(np.random.random_sample(self.batch_size)).to(dev.device)
            b, idx, old_ptr = dev.sample(r)
            rows = int(b["rows"].item())
            tmax = int((b["burn_in"].int() + b["learning"].int() + b["forward"].int()).max().item())
            data = (
                b["obs"][:, :tmax].clone(), b["last_action"][:, :tmax].bool(), b["last_reward"][:, :tmax].clone(),
                b["hidden"].clone().transpose(0, 1),
                b["action"][:rows].clone().unsqueeze(1), b["n_step_reward"][:rows].clone(), b["gamma"][:rows].clone(),
                b["burn_in"].clone(), b["learning"].clone(), b["forward"].clone(),
                idx.cpu().numpy(), b["is_weights"][:rows].clone(), old_ptr, np.int32(dev.env_steps),
            )
        return data

    def update_priorities(self, idxes, td_errors, old_ptr: int, loss: float):      # worker.py:242-261
        """Update priorities of sampled transitions (stale-index mask fused into the tree kernel)."""
        with self.lock:
            dev = self._device_replay()
            idx = idxes if isinstance(idxes, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(idxes, dtype=np.int64))
            td = td_errors if isinstance(td_errors, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(td_errors, dtype=np.float32))
            dev.update_priorities(idx.to(dev.device), td.to(dev.device).float(), old_ptr)
        self.training_steps += 1
        self.sum_loss += loss

    # ------------------------------------------------------------------ forked process: CPU stager (train.py:36)
    def _log_lines(self, interval):
        """The reference's periodic report (worker.py:89-106), one string per line."""
        out = [f'buffer size: {self.size}', f'buffer update speed: {(self.size-self.last_size)/interval}/s',
               f'number of environment steps: {self.env_steps}']
        self.last_size = self.size
        if self.num_episodes != 0:
            out.append(f'average episode return: {self.episode_reward/self.num_episodes:.4f}')
            self.episode_reward, self.num_episodes = 0, 0
        done = self.training_steps - self.last_training_steps
        out += [f'number of training steps: {self.training_steps}', f'training speed: {done/interval}/s']
        if done:
            out.append(f'loss: {self.sum_loss/done:.4f}')
            self.last_training_steps, self.sum_loss = self.training_steps, 0
        return out

    def run(self):
        for target in (self.add_data, self.update_data):
            threading.Thread(target=target, daemon=True).start()
        while True:
            print("\n".join(self._log_lines(config.log_interval)) + "\n")
            if self.training_steps >= config.training_steps:
                return
            time.sleep(config.log_interval)

    def add_data(self):
        """Drain the actor queues and forward every block to the process that owns the GPU (worker.py:124-129)."""
        while True:
            moved = False
            for sample_queue in self.sample_queue_list:
                if not sample_queue.empty():
                    block, priority, episode_reward = sample_queue.get_nowait()
                    self.batch_queue.put((BLOCK_MSG, block, priority, episode_reward))
                    if episode_reward:
                        self.episode_reward += episode_reward
                        self.num_episodes += 1
                    moved = True
            if not moved:
                time.sleep(0.001)

    def update_data(self):
        """Consume learner statistics for the log loop (the priorities themselves stay on the GPU)."""
        while True:
            if not self.priority_queue.empty():
                msg = self.priority_queue.get_nowait()
                if isinstance(msg, tuple) and len(msg) == 5 and msg[0] == STATS_MSG:
                    _, steps, loss_sum, self.size, self.env_steps = msg
                    self.sum_loss += loss_sum
                    self.training_steps = steps
            else:
                time.sleep(0.1)


############################## Learner ##############################

def calculate_mixed_td_errors(td_error, learning_steps):                           # worker.py:268-276
    """Per-sequence priority 0.9*max + 0.1*mean over ragged segments (host version, used by the actors;
    the learner's copy is fused into the TD kernel).  The running offset is a Python int: the upstream uint8
    accumulator wraps at 256 under NumPy >= 2."""
    start_idx = 0
    mixed_td_errors = np.empty(learning_steps.shape, dtype=td_error.dtype)
    for i, steps in enumerate(learning_steps):
        steps = int(steps)
        seg = td_error[start_idx:start_idx + steps]
        mixed_td_errors[i] = 0.9 * seg.max() + 0.1 * seg.mean()
        start_idx += steps
    return mixed_td_errors


class Learner:
    def __init__(self, batch_queue, priority_queue, model, grad_norm: int = config.grad_norm,
                 lr: float = config.lr, eps: float = config.eps, game_name: str = config.game_name,
                 target_net_update_interval: int = config.target_net_update_interval, save_interval: int = config.save_interval,
                 device=None):
        from .learner_core import DeviceLearner
        self.device = torch.device(device if device is not None else 'cuda')      # worker.py:283 -- no CPU fallback here
        self._start_time = time.time()                                             # reset by run(); checkpoints record minutes since then
        self.action_dim = model.action_dim
        self.obs_shape = tuple(model.obs_shape)
        self.batch_size = config.batch_size
        self.seq_frames = config.burn_in_steps + config.learning_steps + config.forward_steps
        self.core = DeviceLearner(self.action_dim, self.batch_size, self.seq_frames, in_channels=self.obs_shape[0],
                                  max_learning=config.learning_steps, max_forward=config.forward_steps, lr=lr, eps=eps,
                                  grad_norm=grad_norm, device=self.device)
        self.core.load_state_dict(model.state_dict())                              # online and target start equal (worker.py:284-288)
        self.grad_norm = grad_norm
        self.batch_queue = batch_queue
        self.priority_queue = priority_queue
        self.num_updates = 0
        self.done = False
        self.target_net_update_interval = target_net_update_interval
        self.save_interval = save_interval
        self.batched_data = []
        self.shared_model = model
        self.game_name = game_name
        self.replay = None                     # DeviceReplay, created on the first forwarded block
        self.is_weight_sync = None             # dist.GlobalISWeights in data-parallel runs
        # update_from_replay can run the priority update of update i and the sampling + gather of batch i+1 on a second
        # stream WHILE update i's backward pass runs (the priorities are final after K2): same sequence of tree operations
        # as the sequential loop, bit-identical results (tests/test_gpu_sample_ahead.py), off the critical path.
        # R2D2_SAMPLE_AHEAD=0 / sample_ahead = False keep the plain sequential loop.
        self.sample_ahead = os.environ.get("R2D2_SAMPLE_AHEAD", "1") != "0"
        self._ahead = None
        self._ahead_at = -1
        self._results = None                   # two pinned result slots (priorities, loss) for enqueue_update/collect
        self._sum_loss = 0.0
        self.env_steps = 0
        self._stager = None
        from .learner_core import WeightPublisher                        # eager: pinning 17 MB and starting the thread take
        self._publisher = WeightPublisher(self.core.online, self.shared_model)                                # milliseconds

    # -- parameters ---------------------------------------------------------------------------------------------
    def state_dict(self):
        return self.core.online.state_dict()

    def store_weights(self, wait: bool = False):                                   # worker.py:306-307
        """Publish the online weights to the shared (CPU) model the actors read; asynchronous (side-stream D2H into pinned
        memory + a daemon thread doing the host copy), so the learner loop does not stall on it."""
        if self._publisher is None:
            from .learner_core import WeightPublisher
            self._publisher = WeightPublisher(self.core.online, self.shared_model)
        self._publisher.publish()
        if wait:
            self._publisher.wait()

    # -- one update from a reference-format host/device 14-tuple (worker.py:330-369) -----------------------------
    def prefetch(self, data):
        """Start moving a 14-tuple to the device on the copy stream (the reference prefetches from its queue in a thread,
        worker.py:309-316).  Returns a staged handle accepted by update_from_batch."""
        from .learner_core import BatchStager
        if self._stager is None:
            self._stager = BatchStager(self.core)
        (batch_obs, batch_last_action, batch_last_reward, batch_hidden, batch_action, batch_n_step_reward, batch_n_step_gamma,
         burn_in_steps, learning_steps, forward_steps, idxes, is_weights, old_ptr, env_steps) = data
        handle = self._stager.stage(dict(obs=batch_obs, last_action=batch_last_action, last_reward=batch_last_reward,
                                         hidden=batch_hidden, action=batch_action, n_step_reward=batch_n_step_reward,
                                         gamma=batch_n_step_gamma, burn_in=burn_in_steps, learning=learning_steps,
                                         forward=forward_steps, is_weights=is_weights))
        return ("staged", handle, idxes, old_ptr, env_steps)

    def enqueue_update(self, data):
        """Launch one update on a (staged or raw) 14-tuple WITHOUT waiting for it: the priorities / loss are copied into a
        pinned result slot behind the kernels and an event marks them ready.  Returns a ticket for collect().  Two tickets
        may be outstanding, so a caller can launch update i+1 before it reads update i's results -- the ~0.4 ms the host
        needs to enqueue an update's 60 launches then overlaps the previous update instead of idling the GPU."""
        if not (isinstance(data, tuple) and len(data) == 5 and isinstance(data[0], str) and data[0] == "staged"):
            data = self.prefetch(data)
        _, handle, idxes, old_ptr, env_steps = data
        b = self._stager.acquire(handle)
        self.core.update(b)
        self._stager.release(handle)
        if self._results is None:
            self._results = [dict(prio=torch.empty(self.batch_size, dtype=torch.float32).pin_memory(),
                                  loss=torch.empty(2, dtype=torch.float32).pin_memory(), ready=torch.cuda.Event())
                             for _ in range(2)]
            self._result_i = 0
        slot = self._results[self._result_i]
        self._result_i ^= 1
        slot["prio"].copy_(self.core.prio, non_blocking=True)                      # worker.py:357: priorities back to the host
        slot["loss"][0:1].copy_(self.core.loss_sum, non_blocking=True)
        slot["loss"][1:2].copy_(self.core.rows.float(), non_blocking=True)
        slot["ready"].record(torch.cuda.current_stream(self.device))
        self.env_steps = env_steps
        self._after_update()
        return (slot, idxes, old_ptr)

    def collect(self, ticket):
        """(idxes, priorities f32[B], old_ptr, loss) of an enqueued update -- the message worker.py:369 puts on the queue."""
        slot, idxes, old_ptr = ticket
        slot["ready"].synchronize()
        loss = float(slot["loss"][0] / slot["loss"][1])
        return idxes, slot["prio"].numpy().copy(), old_ptr, loss

    def update_from_batch(self, data):
        return self.collect(self.enqueue_update(data))

    # -- one update from the HBM-resident replay (sample -> update -> priority update, no host round trip) --------
    def update_from_replay(self):
        if self.sample_ahead and (self.is_weight_sync is None or getattr(self.is_weight_sync, "ahead_safe", False)):
            return self._update_from_replay_ahead()
        batch, idx, old_ptr = self.replay.sample(fuse_into=self.core)     # frames go straight into conv1's staging layout
        self.core.select_s2d(0)
        if self.is_weight_sync is not None:                               # data parallel: weights of one sampler over all shards
            self.is_weight_sync.correct(self.replay, batch, idx)          # side stream; consumed by the gradient hook
        self.core.update(batch)
        if self.is_weight_sync is not None and hasattr(self.is_weight_sync, "join"):
            self.is_weight_sync.join()                                     # side-stream reads of the tree end before it changes
        self.replay.update_priorities(idx, self.core.prio, old_ptr)
        self.env_steps = self.replay.env_steps
        self._after_update()

    # -- the same update with the NEXT batch sampled and gathered while this one trains ---------------------------
    def _sample_ahead(self, slot: int, after=None) -> dict:
        """Sampling stream: [priority update of the running update `after`] -> sample -> gather into buffer set `slot`.
        With `after`, everything sits behind r2d2_net_shadow_gate: it starts when that update's BPTT recurrence is executing --
        its priorities are final by then (K2 precedes the backward pass) -- and runs on the SMs the recurrence leaves idle.
        The tree therefore sees exactly the sequence priority-update(i) -> sample(i+1) of the sequential loop."""
        st = self._sample_stream
        sync = self.is_weight_sync
        with torch.cuda.stream(st):
            st.wait_event(self._tree_safe)                                # tree writes issued on the learner's stream so far
            if after is not None:
                st.wait_event(self.core.td_event)                         # K2 of the running update has written its priorities
                _lib.check(_lib.lib().r2d2_net_shadow_gate(self.core._h, st.cuda_stream))   # placement: start under its BPTT kernel
                if sync is not None and hasattr(sync, "join"):
                    sync.join()                                           # its reads of the tree end before the tree changes
                self.replay.update_priorities(after["idx"], self.core.prio, after["old_ptr"])
            batch, idx, old_ptr = self.replay.sample(fuse_into=self.core, slot=slot)
            if sync is not None:
                sync.correct(self.replay, batch, idx)
            ready = torch.cuda.Event()
            ready.record(st)
        return dict(batch=batch, idx=idx, old_ptr=old_ptr, ready=ready, slot=slot)

    @_lib.on_device
    def _update_from_replay_ahead(self):
        main = torch.cuda.current_stream(self.device)
        if self._ahead is not None and self._ahead_at != self.core.num_updates:
            self._ahead = None                                            # other updates ran in between: the gates lost their pairing
        if self._ahead is None:
            if getattr(self, "_sample_stream", None) is None:
                self._sample_stream = torch.cuda.Stream(device=self.device)
                self._tree_safe = torch.cuda.Event()
                self.replay.set_copy_smem(32 * 1024)                      # copy CTAs cannot share an SM with a recurrence / GEMM CTA
            _lib.check(_lib.lib().r2d2_net_shadow_gate_reset(self.core._h, main.cuda_stream))
            self._tree_safe.record(main)
            self._ahead = self._sample_ahead(0)
        cur = self._ahead
        main.wait_event(cur["ready"])
        self.core.select_s2d(cur["slot"])
        self._tree_safe.record(main)                                      # e.g. blocks ingested since the last update
        self.core.update(cur["batch"])
        self._ahead = self._sample_ahead(1 - cur["slot"], after=cur)      # enqueued now, runs mid-update
        main.wait_event(self._ahead["ready"])                             # the learner's stream finds the tree quiescent after the update
        self._ahead_at = self.core.num_updates
        self.env_steps = self.replay.env_steps
        self._after_update()

    def _after_update(self):
        self.num_updates += 1
        if self.num_updates % 4 == 0:                                              # worker.py:372-373
            self.store_weights()
        if self.num_updates % self.target_net_update_interval == 0:                # worker.py:376-377
            self.core.sync_target()
        if self.num_updates % self.save_interval == 0:                             # worker.py:380-381
            os.makedirs('models', exist_ok=True)
            torch.save((self.core.online.state_dict(device='cpu'), self.num_updates, self.env_steps,
                        (time.time() - self._start_time) / 60),
                       os.path.join('models', '{}{}.pth'.format(self.game_name, self.num_updates)))

    def _ingest(self, msg):
        if msg[0] == STAGED_MSG:                                                   # packed into pinned memory by the prefetch thread
            _, handle, priority, episode_reward = msg
            self.replay.commit(handle, priority, episode_reward, defer=True)       # the H2D copy overlaps the running update
            return
        _, block, priority, episode_reward = msg
        if self.replay is None:
            from .replay import DeviceReplay
            self.replay = DeviceReplay(config.buffer_capacity, config.block_length, config.burn_in_steps, config.learning_steps,
                                       config.forward_steps, self.action_dim, self.obs_shape, config.hidden_dim,
                                       config.prio_exponent, config.importance_sampling_exponent, self.batch_size,
                                       device=self.device)
        self.replay.add(block, priority, episode_reward)

    def prepare_data(self):                                                        # worker.py:309-316
        while True:
            if not self.batch_queue.empty() and len(self.batched_data) < 64:
                data = self.batch_queue.get_nowait()
                if self.replay is not None and isinstance(data, tuple) and len(data) == 4 and isinstance(data[0], str) and data[0] == BLOCK_MSG:
                    # host half of the insertion here, off the learner thread: 3-12 MB of packing per block
                    data = (STAGED_MSG, self.replay.stage(data[1]), data[2], data[3])
                self.batched_data.append(data)
            else:
                time.sleep(0.001)

    def run(self):
        threading.Thread(target=self.prepare_data, daemon=True).start()
        self._start_time = time.time()
        last_stats = time.time()
        loss_acc = torch.zeros(1, device=self.device)
        pending = None
        while self.num_updates < config.training_steps:
            worked = False
            while self.batched_data:
                data = self.batched_data.pop(0)
                worked = True
                if isinstance(data, tuple) and len(data) == 4 and isinstance(data[0], str) and data[0] in (BLOCK_MSG, STAGED_MSG):
                    self._ingest(data)
                else:                                                              # reference-format 14-tuple
                    staged = data if (isinstance(data[0], str) and data[0] == "staged") else self.prefetch(data)
                    if self.batched_data and not isinstance(self.batched_data[0][0], str):
                        self.batched_data[0] = self.prefetch(self.batched_data[0])   # H2D of the next batch overlaps this update
                    ticket = self.enqueue_update(staged)                           # launched; the PREVIOUS update's results go out
                    if pending is not None:                                        # while this one runs (the reference's queue of
                        self.priority_queue.put(self.collect(pending))             # prefetched batches lags priorities further)
                    pending = ticket
                    break
            if self.replay is not None and len(self.replay) >= config.learning_starts:
                self.update_from_replay()
                loss_acc += self.core.loss_sum / self.core.rows
                worked = True
                if time.time() - last_stats > 1.0 or self.num_updates >= config.training_steps:
                    self.priority_queue.put((STATS_MSG, self.num_updates, float(loss_acc.item()), len(self.replay),
                                             self.replay.env_steps))
                    loss_acc.zero_()
                    last_stats = time.time()
            if not worked:
                if pending is not None:
                    self.priority_queue.put(self.collect(pending))
                    pending = None
                time.sleep(0.01)
        if pending is not None:
            self.priority_queue.put(self.collect(pending))

    @staticmethod
    def value_rescale(value, eps=1e-3):                                            # worker.py:383-385
        return value.sign() * ((value.abs() + 1).sqrt() - 1) + eps * value

    @staticmethod
    def inverse_value_rescale(value, eps=1e-3):                                    # worker.py:387-390
        temp = ((1 + 4 * eps * (value.abs() + 1 + eps)).sqrt() - 1) / (2 * eps)
        return value.sign() * (temp.square() - 1)


############################## Actor ##############################

class LocalBuffer:
    '''store transitions of one episode and cut them into blocks (worker.py:395-497)'''

    def __init__(self, action_dim: int, forward_steps: int = config.forward_steps,
                 burn_in_steps=config.burn_in_steps, learning_steps: int = config.learning_steps,
                 gamma: float = config.gamma, hidden_dim: int = config.hidden_dim, block_length: int = config.block_length):
        self.action_dim = action_dim
        self.gamma = gamma
        self.hidden_dim = hidden_dim
        self.forward_steps = forward_steps
        self.learning_steps = learning_steps
        self.burn_in_steps = burn_in_steps
        self.block_length = block_length
        self.curr_burn_in_steps = 0

    def __len__(self):
        return self.size

    def _onehot(self, a):
        v = np.zeros(self.action_dim, dtype=bool)
        v[a] = True
        return v

    def reset(self, init_obs: np.ndarray):
        self.obs_buffer = [init_obs]
        self.last_action_buffer = [self._onehot(0)]
        self.last_reward_buffer = [0]
        self.hidden_buffer = [np.zeros((2, self.hidden_dim), dtype=np.float32)]
        self.action_buffer = []
        self.reward_buffer = []
        self.qval_buffer = []
        self.curr_burn_in_steps = 0
        self.size = 0
        self.sum_reward = 0
        self.done = False

    def add(self, action: int, reward: float, next_obs: np.ndarray, q_value: np.ndarray, hidden_state: np.ndarray):
        self.action_buffer.append(action)
        self.reward_buffer.append(reward)
        self.hidden_buffer.append(hidden_state)
        self.obs_buffer.append(next_obs)
        self.last_action_buffer.append(self._onehot(action))
        self.last_reward_buffer.append(reward)
        self.qval_buffer.append(q_value)
        self.sum_reward += reward
        self.size += 1

    def finish(self, last_qval: np.ndarray = None) -> Tuple:
        n, F, L = self.size, self.forward_steps, self.learning_steps
        assert n <= self.block_length
        num_sequences = math.ceil(n / L)
        tail = min(n, F)

        # n-step discount per step: gamma^F, shrinking to gamma^k over the last steps when the block is cut
        # mid-episode (bootstrapped from last_qval), or 0 when the episode ended (stands in for `done`)
        n_step_gamma = np.full(n, self.gamma ** F, dtype=np.float64)
        if last_qval is not None:
            self.qval_buffer.append(last_qval)
            n_step_gamma[n - tail:] = [self.gamma ** k for k in range(tail, 0, -1)]
        else:
            self.done = True
            self.qval_buffer.append(np.zeros_like(self.qval_buffer[0]))
            n_step_gamma[n - tail:] = 0
        n_step_gamma = n_step_gamma.astype(np.float32)

        obs = np.stack(self.obs_buffer)
        last_action = np.stack(self.last_action_buffer)
        last_reward = np.array(self.last_reward_buffer, dtype=np.float32)
        hiddens = np.stack(self.hidden_buffer[0:n:L])
        actions = np.array(self.action_buffer, dtype=np.uint8)
        qval_buffer = np.concatenate(self.qval_buffer)

        # R_t = sum_{i<F} gamma^i r_{t+i}, rewards past the block end count as zero
        rewards = np.array(self.reward_buffer + [0] * (F - 1), dtype=np.float64)
        n_step_reward = np.convolve(rewards, [self.gamma ** (F - 1 - i) for i in range(F)], 'valid').astype(np.float32)

        starts = np.arange(num_sequences) * L
        burn_in_steps = np.minimum(starts + self.curr_burn_in_steps, self.burn_in_steps).astype(np.uint8)
        learning_steps = np.minimum(L, n - starts).astype(np.uint8)
        ends = np.cumsum(learning_steps.astype(np.int64))
        forward_steps = np.minimum(F, n + 1 - ends).astype(np.uint8)
        assert forward_steps[-1] == 1 and burn_in_steps[0] == self.curr_burn_in_steps

        # initial priorities from the actor's own Q-values: plain n-step max-Q TD, no value rescaling
        max_qval = np.max(qval_buffer[tail:n + 1], axis=1)
        max_qval = np.pad(max_qval, (0, tail - 1), 'edge')
        target_qval = qval_buffer[np.arange(n), actions]
        td_errors = np.abs(n_step_reward + n_step_gamma * max_qval - target_qval, dtype=np.float32)
        priorities = np.zeros(self.block_length // L, dtype=np.float32)
        priorities[:num_sequences] = calculate_mixed_td_errors(td_errors, learning_steps)

        # keep the last burn_in_steps+1 frames as burn-in context of the next block
        keep = self.burn_in_steps + 1
        self.obs_buffer = self.obs_buffer[-keep:]
        self.last_action_buffer = self.last_action_buffer[-keep:]
        self.last_reward_buffer = self.last_reward_buffer[-keep:]
        self.hidden_buffer = self.hidden_buffer[-keep:]
        self.action_buffer.clear()
        self.reward_buffer.clear()
        self.qval_buffer.clear()
        self.curr_burn_in_steps = len(self.obs_buffer) - 1
        self.size = 0

        block = Block(obs, last_action, last_reward, actions, n_step_reward, n_step_gamma, hiddens, num_sequences,
                      burn_in_steps, learning_steps, forward_steps)
        return [block, priorities, self.sum_reward if self.done else None]


class Actor:
    """epsilon-greedy environment worker (worker.py:500-574): CPU env stepping + batch-1 CPU inference, ships
    [Block, priorities, episode_reward|None] triples and refreshes its weights from the shared model."""

    WEIGHT_REFRESH_STEPS = 400                      # hard-coded upstream (worker.py:560)

    def __init__(self, epsilon: float, model, sample_queue, obs_shape: np.ndarray = config.obs_shape,
                 max_episode_steps: int = config.max_episode_steps, block_length: int = config.block_length):
        self.env = create_env(noop_start=True)
        self.action_dim = self.env.action_space.n
        self.model = Network(self.action_dim)
        self.model.eval()
        self.local_buffer = LocalBuffer(self.action_dim)
        self.epsilon, self.shared_model, self.sample_queue = epsilon, model, sample_queue
        self.max_episode_steps, self.block_length = max_episode_steps, block_length
        self.actor_steps = 0

    def _infer(self, state):
        with torch.no_grad():
            return self.model(state)

    def _ship(self, triple, episode_over):
        if not episode_over and self.epsilon > 0.01:
            triple[2] = None                        # only near-greedy actors report returns of cut episodes
        self.sample_queue.put(triple)

    def play_episode(self):
        state, steps, done = self.reset(), 0, False
        while not done and steps < self.max_episode_steps:
            q_value, hidden = self._infer(state)
            explore = random.random() < self.epsilon
            action = self.env.action_space.sample() if explore else torch.argmax(q_value, 1).item()
            next_obs, reward, done, _ = self.env.step(action)
            state.update(next_obs, action, reward, hidden)
            steps += 1
            self.actor_steps += 1
            self.local_buffer.add(action, reward, next_obs, q_value.numpy(), torch.cat(hidden).numpy())
            if done:
                self._ship(self.local_buffer.finish(), True)
            elif len(self.local_buffer) == self.block_length or steps == self.max_episode_steps:
                boot_q, _ = self._infer(state)
                self._ship(self.local_buffer.finish(boot_q.numpy()), False)
            if self.actor_steps % self.WEIGHT_REFRESH_STEPS == 0:
                self.update_weights()

    def run(self):
        while True:
            self.play_episode()

    def update_weights(self):
        '''load the latest weights from shared model'''
        self.model.load_state_dict(self.shared_model.state_dict())

    def reset(self):
        obs = self.env.reset()
        self.local_buffer.reset(obs)
        return AgentState(torch.from_numpy(obs).unsqueeze(0), self.action_dim)


class VectorActor:
    """N epsilon-greedy environment workers stepped together with ONE batched GPU inference per environment step
    (actor_inference.BatchedPolicy) instead of N batch-1 CPU forwards (worker.py:526-544) -- SURVEY 8(f) item 3.
    Ships the same [Block, priorities, episode_reward|None] triples as `Actor`; per-env epsilons as in train.py:28-31.

    The bootstrap value of a block that is cut mid-episode (worker.py:548-552: one extra inference on the new state) is
    the Q of the NEXT batched step, so it costs nothing: a cut block is finished at the start of the following step."""

    WEIGHT_REFRESH_STEPS = Actor.WEIGHT_REFRESH_STEPS

    def __init__(self, epsilons, model, sample_queue, obs_shape=None, max_episode_steps: int = config.max_episode_steps,
                 block_length: int = config.block_length, device=None, envs=None, policy=None):
        """policy: object with load_state_dict(sd) and step(obs, last_action, last_reward, hidden) -> (q, next_hidden);
        defaults to the CUDA BatchedPolicy (tests inject a stub to exercise the block logic without a GPU)."""
        self.envs = list(envs) if envs is not None else [create_env(noop_start=True) for _ in epsilons]
        self.N = len(self.envs)
        assert self.N == len(epsilons)
        self.action_dim = self.envs[0].action_space.n
        obs_shape = tuple(obs_shape if obs_shape is not None else config.obs_shape)
        if policy is None:
            from .actor_inference import BatchedPolicy
            policy = BatchedPolicy(self.action_dim, self.N, obs_shape=obs_shape, device=device)
        self.policy = policy
        self.shared_model, self.sample_queue = model, sample_queue
        self.policy.load_state_dict(model.state_dict())
        self.epsilons = list(epsilons)
        self.max_episode_steps, self.block_length = max_episode_steps, block_length
        self.buffers = [LocalBuffer(self.action_dim, block_length=block_length) for _ in range(self.N)]
        self.obs = np.zeros((self.N,) + obs_shape, dtype=np.uint8)
        self.last_action = np.zeros((self.N, self.action_dim), dtype=np.uint8)
        self.last_reward = np.zeros(self.N, dtype=np.float32)
        self.hidden = torch.zeros(self.N, 2, config.hidden_dim)
        self.episode_steps = [0] * self.N
        self.pending_cut = [False] * self.N
        self.actor_steps = 0
        for i in range(self.N):
            self._reset(i)

    def _reset(self, i):
        obs = self.envs[i].reset()
        self.buffers[i].reset(obs)
        self.obs[i] = obs
        self.last_action[i] = 0
        self.last_reward[i] = 0.0
        self.hidden[i] = 0.0
        self.episode_steps[i] = 0
        self.pending_cut[i] = False

    def _ship(self, i, triple, episode_over):
        if not episode_over and self.epsilons[i] > 0.01:
            triple[2] = None
        self.sample_queue.put(triple)

    def step(self):
        """one environment step of every actor"""
        q, hidden = self.policy.step(self.obs, self.last_action, self.last_reward, self.hidden)
        q = q.cpu().numpy()
        self.hidden = hidden.cpu()
        hid_np = self.hidden.numpy()
        for i, env in enumerate(self.envs):
            qi = q[i:i + 1]
            if self.pending_cut[i]:                                 # bootstrap of the block cut at the previous step
                self._ship(i, self.buffers[i].finish(qi), False)
                self.pending_cut[i] = False
                if self.episode_steps[i] >= self.max_episode_steps:
                    self._reset(i)
                    continue                                        # this actor acts again from the fresh episode next step
            explore = random.random() < self.epsilons[i]
            action = env.action_space.sample() if explore else int(np.argmax(qi, 1)[0])
            next_obs, reward, done, _ = env.step(action)
            self.buffers[i].add(action, reward, next_obs, qi, hid_np[i].copy())
            self.episode_steps[i] += 1
            self.obs[i] = next_obs
            self.last_action[i] = 0
            self.last_action[i, action] = 1
            self.last_reward[i] = reward
            if done:
                self._ship(i, self.buffers[i].finish(), True)
                self._reset(i)
            elif len(self.buffers[i]) == self.block_length or self.episode_steps[i] == self.max_episode_steps:
                self.pending_cut[i] = True
        self.actor_steps += 1
        if self.actor_steps % self.WEIGHT_REFRESH_STEPS == 0:
            self.update_weights()

    def run(self):
        while True:
            self.step()

    def update_weights(self):
        self.policy.load_state_dict(self.shared_model.state_dict())
