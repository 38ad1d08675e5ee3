"""Shared builders for oracle/GPU tests (seed-driven; no reference needed)."""
import zlib

import numpy as np

from oracle import synth
from oracle.replay import ActorBlockOracle, ReplayOracle

A = 9


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def build_oracle_replay(script, num_blocks, batch_size, bl=400, ls=40, bi=40, fs=5, actor_cls=None,
                        replay=None):
    rb = replay if replay is not None else ReplayOracle(num_blocks * bl, bl, ls, 0.9, 0.6, batch_size)
    blocks = []
    for seed, steps, done in script:
        if actor_cls is None:
            lb = ActorBlockOracle(A, fs, bi, ls, 0.997, 512, bl)
        else:
            lb = actor_cls(A, forward_steps=fs, burn_in_steps=bi, learning_steps=ls, block_length=bl)
        for blk, prio, ep in synth.drive_actor(lb, seed, steps, done, A, block_length=bl):
            rb.add(blk, prio, ep)
            blocks.append((blk, prio, ep))
    return rb, blocks


def sample_with_seed(rb, seed):
    r = np.random.RandomState(seed).random_sample(rb.batch_size)
    return rb.sample_batch(r)
