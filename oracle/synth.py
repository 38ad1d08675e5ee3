"""Seeded synthetic actor traffic (TEST INFRASTRUCTURE).

Everything is a pure function of integer seeds through numpy's PCG64, so the
build container (where the reference can be imported to make golden vectors)
and the GPU box (where it cannot) regenerate byte-identical inputs without
shipping frames.  Shapes/distributions follow SURVEY.md section 8(d): frames
u8 ~ U{0..255}, actions U{0..A-1}, rewards Bernoulli(0.5), actor q ~ N(0,1),
stored recurrent state 0.1*N(0,1).

``drive_actor`` works with any object exposing the reference ``LocalBuffer``
protocol (reset/add/finish): the reference's own class, the oracle restatement
or the product's host-side class.
"""
from __future__ import annotations

import numpy as np


def episode_frames(seed: int, n: int, channels: int = 1) -> np.ndarray:
    rng = np.random.default_rng([int(seed), 0xF7A3E5])
    return rng.integers(0, 256, size=(n, channels, 84, 84), dtype=np.uint8)


def drive_actor(local_buffer, seed: int, n_steps: int, done: bool, action_dim: int,
                hidden_dim: int = 512, block_length: int = 400, channels: int = 1):
    """Play one scripted episode of ``n_steps`` env steps; return the emitted
    [block, priorities, episode_reward] triples (worker.py:546-558 protocol)."""
    frames = episode_frames(seed, n_steps + 1, channels)
    rng = np.random.default_rng([int(seed), 0x51DE])
    out = []
    local_buffer.reset(frames[0])
    in_block = 0
    for t in range(n_steps):
        action = int(rng.integers(0, action_dim))
        reward = float(rng.integers(0, 2))
        q = rng.standard_normal((1, action_dim)).astype(np.float32)
        hid = (0.1 * rng.standard_normal((2, hidden_dim))).astype(np.float32)
        local_buffer.add(action, reward, frames[t + 1], q, hid)
        in_block += 1
        last = t == n_steps - 1
        if last and done:
            out.append(local_buffer.finish())
        elif in_block == block_length or last:
            q_boot = rng.standard_normal((1, action_dim)).astype(np.float32)
            out.append(local_buffer.finish(q_boot))
            in_block = 0
    return out


# The episode script used by the learner golden fixtures: (seed, steps, done)
RAGGED_SCRIPT = [
    (11, 930, True),    # 400 (burn-in 0 first block) + 400 + 130-step done tail (l=40,40,40,10)
    (12, 57, True),     # short episode: l = 40,17 ; burn-in 0 / 40 ; gamma tail 0
    (13, 400, False),   # exactly one full block, cut mid-episode (bootstrap gammas)
    (14, 83, False),    # short, cut mid-episode (max_episode_steps style), l = 40,40,3
]


def synthetic_batch(B: int, action_dim: int, burn_in: int = 40, learning: int = 40, forward: int = 5,
                    channels: int = 1, hidden_dim: int = 512, seed: int = 0, ragged: bool = False):
    """A learner batch drawn directly (no replay), BASELINE config #2 shape.

    Returns the dict layout of ``ReplayOracle.sample_batch`` (numpy arrays).
    With ``ragged`` a few sequences get short burn-in / learning / forward counts.
    """
    rng = np.random.default_rng([int(seed), 0xBA7C4])
    b = np.full(B, burn_in, dtype=np.uint8)
    l = np.full(B, learning, dtype=np.uint8)
    f = np.full(B, forward, dtype=np.uint8)
    if ragged and B >= 4:
        b[0] = 0
        l[1], f[1] = max(1, learning // 3), 1
        f[2] = 1
        b[3], l[3], f[3] = burn_in // 2, max(1, learning - 3), min(forward, 2)
    T = int((b.astype(int) + l + f).max())
    obs = rng.integers(0, 256, size=(B, T, channels, 84, 84), dtype=np.uint8)
    act_prev = rng.integers(0, action_dim, size=(B, T))
    last_action = np.zeros((B, T, action_dim), dtype=bool)
    np.put_along_axis(last_action, act_prev[..., None], True, axis=2)
    last_reward = rng.integers(0, 2, size=(B, T)).astype(np.float32)
    for n in range(B):                       # zero padding at the END of the time axis
        Tn = int(b[n]) + int(l[n]) + int(f[n])
        obs[n, Tn:] = 0
        last_action[n, Tn:] = False
        last_reward[n, Tn:] = 0
    sumL = int(l.astype(int).sum())
    hidden = (0.1 * rng.standard_normal((B, 2, hidden_dim))).astype(np.float32)
    gamma = np.full(sumL, 0.997 ** forward, dtype=np.float32)
    prio = rng.uniform(0.5, 1.5, size=B)
    is_w = np.power(prio / prio.min(), -0.6)
    return dict(obs=obs, last_action=last_action, last_reward=last_reward, hidden=hidden,
                action=rng.integers(0, action_dim, size=sumL).astype(np.uint8),
                n_step_reward=rng.uniform(0, 3, size=sumL).astype(np.float32), gamma=gamma,
                burn_in=b, learning=l, forward=f,
                idxes=np.arange(B, dtype=np.int64),
                is_weights=np.repeat(is_w, l).astype(np.float32), old_ptr=0, env_steps=0)


def to_torch_batch(d):
    """dict-of-numpy (above / ReplayOracle.sample_batch) -> oracle.learner.Batch."""
    import torch
    from .learner import Batch
    t = torch.from_numpy
    return Batch(obs=t(d["obs"]), last_action=t(d["last_action"]), last_reward=t(d["last_reward"]),
                 hidden=t(np.ascontiguousarray(d["hidden"])).transpose(0, 1),
                 action=t(d["action"]).unsqueeze(1), n_step_reward=t(d["n_step_reward"]),
                 n_step_gamma=t(d["gamma"]), burn_in=t(d["burn_in"]), learning=t(d["learning"]),
                 forward=t(d["forward"]), is_weights=t(d["is_weights"]))
