"""NumPy restatement of the reference's replay side of the hot path (TEST INFRASTRUCTURE).

  * ``ActorBlockOracle``  -- ``LocalBuffer``  worker.py:395-497  (the wire format
    the learner-side buffer ingests: how an actor cuts <=block_length steps into
    sequences, n-step rewards / gammas, per-sequence burn-in/learning/forward
    counts and initial priorities)
  * ``ReplayOracle``      -- ``ReplayBuffer`` worker.py:141-261  (``add``,
    ``sample_batch``, ``update_priorities``) on top of ``SumTreeOracle``.

Sequences are identified exactly as in the reference: slot = block*seq_per_block
+ sequence; ragged rows are sequence-major; frames are zero padded at the END
of the time axis (``pad_sequence``, worker.py:212-214).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .sumtree import SumTreeOracle
from .learner import mixed_priorities


@dataclass
class Block:                      # worker.py:23-35
    obs: np.ndarray               # u8  (curr_burn_in + size + 1, C, 84, 84)
    last_action: np.ndarray       # bool (same rows, A)
    last_reward: np.ndarray       # f32 (same rows,)
    action: np.ndarray            # u8  (size,)
    n_step_reward: np.ndarray     # f32 (size,)
    gamma: np.ndarray             # f32 (size,)
    hidden: np.ndarray            # f32 (num_sequences, 2, H)
    num_sequences: int
    burn_in_steps: np.ndarray     # u8 (num_sequences,)
    learning_steps: np.ndarray    # u8 (num_sequences,)
    forward_steps: np.ndarray     # u8 (num_sequences,)


class ActorBlockOracle:
    """Episode accumulator, worker.py:395-497."""

    def __init__(self, action_dim, forward_steps=5, burn_in_steps=40, learning_steps=40,
                 gamma=0.997, hidden_dim=512, block_length=400):
        self.A, self.F, self.BI, self.L = action_dim, forward_steps, burn_in_steps, learning_steps
        self.gamma, self.H, self.block_length = gamma, hidden_dim, block_length
        self.carry = 0

    def reset(self, init_obs):                                    # worker.py:413-424
        first_action = np.zeros(self.A, dtype=bool)
        first_action[0] = True
        self.obs = [init_obs]
        self.last_action = [first_action]
        self.last_reward = [0]
        self.hidden = [np.zeros((2, self.H), dtype=np.float32)]
        self.action, self.reward, self.qval = [], [], []
        self.carry, self.size, self.sum_reward, self.done = 0, 0, 0, False

    def add(self, action, reward, next_obs, q_value, hidden_state):   # worker.py:426-435
        onehot = np.zeros(self.A, dtype=bool)
        onehot[action] = True
        self.action.append(action)
        self.reward.append(reward)
        self.hidden.append(hidden_state)
        self.obs.append(next_obs)
        self.last_action.append(onehot)
        self.last_reward.append(reward)
        self.qval.append(q_value)
        self.sum_reward += reward
        self.size += 1

    def finish(self, last_qval: Optional[np.ndarray] = None):         # worker.py:437-497
        n, F_, L = self.size, self.F, self.L
        assert n <= self.block_length
        num_seq = math.ceil(n / L)
        tail = min(n, F_)
        gam = [self.gamma ** F_] * (n - tail)
        if last_qval is not None:          # block cut mid-episode: bootstrap with gamma^k
            self.qval.append(last_qval)
            gam += [self.gamma ** k for k in range(tail, 0, -1)]
        else:                              # episode ended: gamma 0 stands in for `done`
            self.done = True
            self.qval.append(np.zeros_like(self.qval[0]))
            gam += [0] * tail
        gam = np.array(gam, dtype=np.float32)

        obs = np.stack(self.obs)
        last_action = np.stack(self.last_action)
        last_reward = np.array(self.last_reward, dtype=np.float32)
        hiddens = np.stack(self.hidden[0:n:L])
        actions = np.array(self.action, dtype=np.uint8)
        qvals = np.concatenate(self.qval)

        padded_r = self.reward + [0] * (F_ - 1)
        kernel = [self.gamma ** (F_ - 1 - i) for i in range(F_)]
        n_step_reward = np.convolve(padded_r, kernel, 'valid').astype(np.float32)

        burn = np.array([min(i * L + self.carry, self.BI) for i in range(num_seq)], dtype=np.uint8)
        learn = np.array([min(L, n - i * L) for i in range(num_seq)], dtype=np.uint8)
        fwd = np.array([min(F_, n + 1 - int(np.sum(learn[:i + 1], dtype=np.int64)))
                        for i in range(num_seq)], dtype=np.uint8)
        assert fwd[-1] == 1 and burn[0] == self.carry

        # actor-side initial priorities: plain max-Q n-step TD, NO value rescale (worker.py:477-483)
        max_q = np.max(qvals[tail:n + 1], axis=1)
        max_q = np.pad(max_q, (0, tail - 1), 'edge')
        taken_q = qvals[np.arange(n), actions]
        td = np.abs(n_step_reward + gam * max_q - taken_q, dtype=np.float32)
        prio = np.zeros(self.block_length // L, dtype=np.float32)
        prio[:num_seq] = mixed_priorities(td, learn)

        keep = self.BI + 1                                          # worker.py:486-494
        self.obs, self.last_action = self.obs[-keep:], self.last_action[-keep:]
        self.last_reward, self.hidden = self.last_reward[-keep:], self.hidden[-keep:]
        self.action, self.reward, self.qval = [], [], []
        self.carry = len(self.obs) - 1
        self.size = 0

        blk = Block(obs, last_action, last_reward, actions, n_step_reward, gam, hiddens, num_seq,
                    burn, learn, fwd)
        return [blk, prio, self.sum_reward if self.done else None]


class ReplayOracle:
    """worker.py:38-75,141-261 without threads/queues."""

    def __init__(self, buffer_capacity, block_length=400, learning_steps=40, alpha=0.9, beta=0.6,
                 batch_size=64):
        self.seq_len = learning_steps
        self.block_len = block_length
        self.num_blocks = buffer_capacity // block_length
        self.seq_per_block = block_length // learning_steps
        self.tree = SumTreeOracle(buffer_capacity // learning_steps, alpha, beta)
        self.batch_size = batch_size
        self.block_ptr = 0
        self.size = 0
        self.env_steps = 0
        self.buffer: List[Optional[Block]] = [None] * self.num_blocks

    def add(self, block: Block, priority: np.ndarray, episode_reward=None):      # worker.py:141-161
        slots = np.arange(self.block_ptr * self.seq_per_block,
                          (self.block_ptr + 1) * self.seq_per_block, dtype=np.int64)
        self.tree.update(slots, priority)
        old = self.buffer[self.block_ptr]
        if old is not None:
            self.size -= int(np.sum(old.learning_steps, dtype=np.int64))
        self.size += int(np.sum(block.learning_steps, dtype=np.int64))
        self.buffer[self.block_ptr] = block
        self.env_steps += int(np.sum(block.learning_steps, dtype=np.int64))
        self.block_ptr = (self.block_ptr + 1) % self.num_blocks

    def gather(self, idxes: np.ndarray):
        """The slicing half of ``sample_batch`` (worker.py:172-214) for given slots.

        Returns numpy arrays: obs (B,Tmax,C,84,84) u8, last_action (B,Tmax,A) bool,
        last_reward (B,Tmax) f32, hidden (B,2,H), action (sumL,), n_step_reward,
        gamma, burn/learn/fwd (B,) u8.
        """
        obs_l, la_l, lr_l, hid_l, act_l, rew_l, gam_l, b_l, l_l, f_l = ([] for _ in range(10))
        for slot in idxes:
            blk = self.buffer[slot // self.seq_per_block]
            s = int(slot % self.seq_per_block)
            assert s < blk.num_sequences
            b, l, f = int(blk.burn_in_steps[s]), int(blk.learning_steps[s]), int(blk.forward_steps[s])
            before = int(np.sum(blk.learning_steps[:s], dtype=np.int64))
            start = int(blk.burn_in_steps[0]) + before
            obs_l.append(blk.obs[start - b:start + l + f])
            la_l.append(blk.last_action[start - b:start + l + f])
            lr_l.append(blk.last_reward[start - b:start + l + f])
            act_l.append(blk.action[before:before + l])
            rew_l.append(blk.n_step_reward[before:before + l])
            gam_l.append(blk.gamma[before:before + l])
            hid_l.append(blk.hidden[s])
            b_l.append(b); l_l.append(l); f_l.append(f)
        T = max(x.shape[0] for x in obs_l)

        def pad(rows):
            out = np.zeros((len(rows), T) + rows[0].shape[1:], dtype=rows[0].dtype)
            for i, r in enumerate(rows):
                out[i, :r.shape[0]] = r
            return out
        return dict(obs=pad(obs_l), last_action=pad(la_l), last_reward=pad(lr_l),
                    hidden=np.stack(hid_l), action=np.concatenate(act_l),
                    n_step_reward=np.concatenate(rew_l), gamma=np.concatenate(gam_l),
                    burn_in=np.array(b_l, dtype=np.uint8), learning=np.array(l_l, dtype=np.uint8),
                    forward=np.array(f_l, dtype=np.uint8))

    def sample_batch(self, unit_uniforms=None):                                  # worker.py:163-240
        idxes, is_w = self.tree.sample(self.batch_size, unit_uniforms)
        g = self.gather(idxes)
        g["idxes"] = idxes
        g["is_weights"] = np.repeat(is_w, g["learning"]).astype(np.float32)       # worker.py:216,234
        g["old_ptr"] = self.block_ptr
        g["env_steps"] = self.env_steps
        return g

    def stale_mask(self, idxes: np.ndarray, old_ptr: int) -> np.ndarray:          # worker.py:247-256
        lo_new, spb = self.block_ptr, self.seq_per_block
        if lo_new > old_ptr:
            return (idxes < old_ptr * spb) | (idxes >= lo_new * spb)
        if lo_new < old_ptr:
            return (idxes < old_ptr * spb) & (idxes >= lo_new * spb)
        return np.ones(idxes.shape, dtype=bool)

    def update_priorities(self, idxes, td_errors, old_ptr, loss=0.0):            # worker.py:242-261
        keep = self.stale_mask(np.asarray(idxes), old_ptr)
        self.tree.update(np.asarray(idxes)[keep], np.asarray(td_errors)[keep])
