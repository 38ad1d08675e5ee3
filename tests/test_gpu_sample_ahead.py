"""Learner.update_from_replay with sample_ahead: the priority update of update i and the sampling + gather of batch i+1 run
on a second stream (second set of buffers, behind the shadow gate) while update i's backward pass runs.  The tree sees the
same sequence of operations as in the plain sequential loop, so both loops must produce bit-identical priorities, parameters
and trees from the same seeds -- and the pipeline must train on exactly the batches it sampled, in order."""
import numpy as np
import pytest
import torch

from oracle.gen_golden import CFG0_SCRIPT
from oracle.learner import init_params
from helpers import build_oracle_replay, A

pytestmark = pytest.mark.gpu


def _make(monkeypatch, seed, ahead):
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
    learner.sample_ahead = ahead
    replay = DeviceReplay(nb * bl, bl, bi, ls, fs, A, (1, 84, 84), 512, 0.9, 0.6, bs, seed=seed)
    _, blocks = build_oracle_replay(CFG0_SCRIPT, nb, bs, bl, ls, bi, fs, actor_cls=LocalBuffer)
    for blk, prio, ep in blocks:
        replay.add(blk, prio, ep)
    learner.replay = replay
    record = []
    orig = replay.tree.sample_device

    def recording(n, unit_uniforms=None, want_f64=False):
        out = orig(n, unit_uniforms, want_f64)
        record.append((out[0].clone(), out[1].clone()))      # on the stream the sample itself runs on
        return out
    replay.tree.sample_device = recording
    return learner, replay, record


def test_sample_ahead_is_bit_identical_to_the_sequential_loop(monkeypatch):
    K = 14                                                   # eager updates, graph captures and replays for both buffer sets; a target sync
    ahead, rp_a, rec_a = _make(monkeypatch, 5, True)
    seq, rp_s, rec_s = _make(monkeypatch, 5, False)
    for k in range(K):
        ahead.update_from_replay()
        seq.update_from_replay()
        torch.cuda.synchronize()
        assert torch.equal(rec_a[k][0], rec_s[k][0]), f"update {k}: different sampled indices"
        assert torch.equal(rec_a[k][1], rec_s[k][1]), f"update {k}: different importance weights"
        assert torch.equal(ahead.core.prio, seq.core.prio), f"update {k}: different priorities"
    assert len(rec_a) == K + 1 and len(rec_s) == K           # one batch is always in flight
    assert torch.equal(seq.core.online.flat, ahead.core.online.flat)
    assert torch.equal(seq.core.target.flat, ahead.core.target.flat)
    # the in-flight sample-ahead has already applied update K-1's priorities; so has the sequential loop
    np.testing.assert_array_equal(rp_s.tree.ptree, rp_a.tree.ptree)


def test_sample_ahead_survives_interleaved_host_batches(monkeypatch):
    """Updates from host batches between replay updates drop the in-flight batch and re-prime the pipeline (the gate pairing
    is re-based); training continues and stays finite."""
    from oracle import synth
    ahead, rp, _ = _make(monkeypatch, 7, True)
    d = synth.synthetic_batch(8, A, burn_in=8, learning=8, forward=4, seed=11)
    t = torch.from_numpy
    tup = (t(d["obs"]), t(d["last_action"]), t(d["last_reward"]), t(np.ascontiguousarray(d["hidden"])).transpose(0, 1),
           t(d["action"]).unsqueeze(1), t(d["n_step_reward"]), t(d["gamma"]), t(d["burn_in"]), t(d["learning"]), t(d["forward"]),
           d["idxes"], t(d["is_weights"]), 0, np.int32(0))
    for k in range(9):
        ahead.update_from_replay()
        if k % 3 == 2:
            ahead.update_from_batch(tup)
    torch.cuda.synchronize()
    assert torch.isfinite(ahead.core.online.flat).all() and ahead.num_updates == 12
