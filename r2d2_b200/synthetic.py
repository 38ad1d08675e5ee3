"""Seeded synthetic learner inputs for benchmarks and examples (SURVEY.md 8d distributions).

Product-side generators: bench.py's own arm and the examples build their inputs here, so nothing on the product
path imports the test oracle.  Frames u8 ~ U{0..255}, last action one-hot of U{0..A-1}, rewards Bernoulli(0.5),
stored recurrent state 0.1*N(0,1), n-step discount 0.997^5; weights follow PyTorch's default
U(-1/sqrt(fan_in), 1/sqrt(fan_in)) bounds (model.py:39-63 constructs them with the default initialisers).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .learner_core import PARAM_NAMES, param_shapes


def init_state_dict(action_dim: int, in_channels: int = 1, seed: int = 0, gain: float = 1.0, hidden_dim: int = 512):
    """state_dict-shaped random parameters; `gain` scales every tensor (larger |Q| for stress tests)."""
    rng = np.random.default_rng([int(seed), 0x1A17])
    shapes = param_shapes(action_dim, in_channels)
    out = {}
    for name in PARAM_NAMES:
        shape = shapes[name]
        if name.startswith("recurrent"):
            bound = 1.0 / math.sqrt(hidden_dim)
        else:
            w = shapes[name.rsplit(".", 1)[0] + ".weight"]
            bound = 1.0 / math.sqrt(int(np.prod(w[1:])))
        out[name] = torch.from_numpy((rng.uniform(-bound, bound, size=shape) * gain).astype(np.float32))
    return out


def synthetic_batch(B: int, action_dim: int, burn_in: int = 40, learning: int = 40, forward: int = 5, channels: int = 1,
                    hidden_dim: int = 512, seed: int = 0):
    """One learner batch as a dict of NumPy arrays in the field layout of ReplayBuffer.sample_batch (worker.py:219-238)."""
    rng = np.random.default_rng([int(seed), 0x5B47])
    T = burn_in + learning + forward
    obs = rng.integers(0, 256, size=(B, T, channels, 84, 84), dtype=np.uint8)
    last_action = np.zeros((B, T, action_dim), dtype=bool)
    np.put_along_axis(last_action, rng.integers(0, action_dim, size=(B, T))[..., None], True, axis=2)
    rows = B * learning
    prio = rng.uniform(0.5, 1.5, size=B)
    return dict(obs=obs, last_action=last_action, last_reward=rng.integers(0, 2, size=(B, T)).astype(np.float32),
                hidden=(0.1 * rng.standard_normal((B, 2, hidden_dim))).astype(np.float32),
                action=rng.integers(0, action_dim, size=rows).astype(np.uint8),
                n_step_reward=rng.uniform(0, 3, size=rows).astype(np.float32),
                gamma=np.full(rows, 0.997 ** forward, dtype=np.float32),
                burn_in=np.full(B, burn_in, dtype=np.uint8), learning=np.full(B, learning, dtype=np.uint8),
                forward=np.full(B, forward, dtype=np.uint8), idxes=np.arange(B, dtype=np.int64),
                is_weights=np.repeat(np.power(prio / prio.min(), -0.6), learning).astype(np.float32), old_ptr=0, env_steps=0)


def reference_tuple(d: dict, pinned: bool = False):
    """The reference's 14-tuple (worker.py:219-238) of host tensors from a synthetic_batch dict."""
    t = (lambda a: torch.from_numpy(a).pin_memory()) if pinned else torch.from_numpy
    return (t(d["obs"]), t(d["last_action"]), t(d["last_reward"]), t(np.ascontiguousarray(d["hidden"])).transpose(0, 1),
            t(d["action"]).unsqueeze(1), t(d["n_step_reward"]), t(d["gamma"]), t(d["burn_in"]), t(d["learning"]), t(d["forward"]),
            d["idxes"], t(d["is_weights"]), 0, np.int32(0))


def synthetic_blocks(n: int, action_dim: int, channels: int, seed: int, burn_in: int = 40, learning: int = 40, forward: int = 5,
                     block_len: int = 400):
    """n full actor blocks (block_len steps) as (Block, initial priorities) pairs, built directly as arrays."""
    from .worker import Block
    rng = np.random.default_rng([int(seed), 0xB10C])
    frames = burn_in + block_len + 1
    spb = block_len // learning
    out = []
    for _ in range(n):
        la = np.zeros((frames, action_dim), dtype=bool)
        la[np.arange(frames), rng.integers(0, action_dim, frames)] = True
        fwd = np.full(spb, forward, dtype=np.uint8)
        fwd[-1] = 1
        blk = Block(obs=rng.integers(0, 256, size=(frames, channels, 84, 84), dtype=np.uint8), last_action=la,
                    last_reward=rng.integers(0, 2, frames).astype(np.float32),
                    action=rng.integers(0, action_dim, block_len).astype(np.uint8),
                    n_step_reward=rng.uniform(0, 3, block_len).astype(np.float32),
                    gamma=np.full(block_len, 0.997 ** forward, dtype=np.float32),
                    hidden=(0.1 * rng.standard_normal((spb, 2, 512))).astype(np.float32), num_sequences=spb,
                    burn_in_steps=np.full(spb, burn_in, dtype=np.uint8), learning_steps=np.full(spb, learning, dtype=np.uint8),
                    forward_steps=fwd)
        out.append((blk, rng.uniform(0.1, 1.0, spb).astype(np.float32)))
    return out
