"""Learner.update_from_replay with sample_ahead: batch i+1 is sampled and gathered (second set of buffers, sampling stream,
behind the shadow gate) while update i runs.  The pipeline must train on exactly the batches it sampled, in order: a
sequential learner that is fed the recorded (indices, weights) sequence through plain gathers must end with bit-identical
parameters, priorities and tree."""
import numpy as np
import pytest
import torch

from oracle.gen_golden import CFG0_SCRIPT
from oracle.learner import init_params
from helpers import build_oracle_replay, A

pytestmark = pytest.mark.gpu


def _make(monkeypatch, seed):
    from r2d2_b200 import config
    from r2d2_b200.model import Network
    from r2d2_b200.replay import DeviceReplay
    from r2d2_b200.worker import Learner, LocalBuffer
    nb, bs, bl, ls, bi, fs = 8, 8, 40, 8, 8, 4
    for k, v in dict(batch_size=bs, burn_in_steps=bi, learning_steps=ls, forward_steps=fs, block_length=bl).items():
        monkeypatch.setattr(config, k, v)
    model = Network(A)
    model.load_state_dict(init_params(A, seed=3))
    learner = Learner(None, None, model, target_net_update_interval=7, save_interval=10 ** 9)
    replay = DeviceReplay(nb * bl, bl, bi, ls, fs, A, (1, 84, 84), 512, 0.9, 0.6, bs, seed=seed)
    _, blocks = build_oracle_replay(CFG0_SCRIPT, nb, bs, bl, ls, bi, fs, actor_cls=LocalBuffer)
    for blk, prio, ep in blocks:
        replay.add(blk, prio, ep)
    learner.replay = replay
    return learner, replay


def test_sample_ahead_trains_on_the_batches_it_sampled(monkeypatch):
    K = 12                                                   # covers eager updates, graph capture and replays for both buffer sets
    ahead, rp_a = _make(monkeypatch, seed=5)
    ahead.sample_ahead = True
    record = []
    orig = rp_a.tree.sample_device

    def recording(n, unit_uniforms=None, want_f64=False):
        out = orig(n, unit_uniforms, want_f64)
        record.append((out[0].clone(), out[1].clone()))      # on the sampling stream, like the sample itself
        return out
    rp_a.tree.sample_device = recording
    prios_a = []
    for _ in range(K):
        ahead.update_from_replay()
        prios_a.append(ahead.core.prio.clone())
    torch.cuda.synchronize()
    assert len(record) == K + 1                              # one batch is always in flight

    seq, rp_s = _make(monkeypatch, seed=5)
    for k in range(K):
        idx, isw = record[k]
        batch = rp_s.gather_fused(idx, isw, seq.core)
        seq.core.select_s2d(0)
        seq.core.update(batch)
        rp_s.update_priorities(idx, seq.core.prio, rp_s.block_ptr)
        seq._after_update()
        torch.cuda.synchronize()
        assert torch.equal(seq.core.prio, prios_a[k]), k
    assert torch.equal(seq.core.online.flat, ahead.core.online.flat)
    assert torch.equal(seq.core.target.flat, ahead.core.target.flat)
    np.testing.assert_array_equal(rp_s.tree.ptree, rp_a.tree.ptree)
