"""K4 HBM replay (ingest + gather) and the ReplayBuffer surface vs the oracle / reference golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import synth
from oracle.gen_golden import CFG0_SCRIPT
from helpers import build_oracle_replay, crc, A

pytestmark = pytest.mark.gpu


def _build_device_replay(script, num_blocks, batch_size, bl=400, ls=40, bi=40, fs=5):
    from r2d2_b200.replay import DeviceReplay
    from r2d2_b200.worker import LocalBuffer
    dev = DeviceReplay(num_blocks * bl, bl, bi, ls, fs, A, (1, 84, 84), 512, 0.9, 0.6, batch_size)
    cpu, blocks = build_oracle_replay(script, num_blocks, batch_size, bl, ls, bi, fs, actor_cls=LocalBuffer)
    for blk, prio, ep in blocks:
        dev.add(blk, prio, ep)
    return dev, cpu


@pytest.mark.parametrize("script,nb,bs,bl,ls,bi,fs", [(synth.RAGGED_SCRIPT, 8, 8, 400, 40, 40, 5), (CFG0_SCRIPT, 8, 4, 40, 8, 8, 4)])
def test_device_replay_matches_oracle_bit_for_bit(script, nb, bs, bl, ls, bi, fs):
    dev, cpu = _build_device_replay(script, nb, bs, bl, ls, bi, fs)
    assert (dev.size, dev.env_steps, dev.block_ptr) == (cpu.size, cpu.env_steps, cpu.block_ptr)
    # leaves may differ from NumPy's float32 pow in the last ulp (tested in test_gpu_tree); pin them for exact sampling
    L = cpu.tree.num_layers
    dev.tree.set_leaves_device(torch.arange(1 << (L - 1), device="cuda"), torch.from_numpy(cpu.tree.leaves.copy()).cuda())
    for seed in (7, 8, 9):
        r = np.random.RandomState(seed).random_sample(bs)
        want = cpu.sample_batch(r)
        got, idx, old_ptr = dev.sample(torch.from_numpy(r).cuda())
        torch.cuda.synchronize()
        rows = int(got["rows"].item())
        np.testing.assert_array_equal(idx.cpu().numpy(), want["idxes"])
        assert old_ptr == want["old_ptr"] and rows == want["action"].shape[0]
        T = want["obs"].shape[1]
        np.testing.assert_array_equal(got["obs"][:, :T].cpu().numpy(), want["obs"])
        assert not got["obs"][:, T:].any()                                      # zero padding at the end
        np.testing.assert_array_equal(got["last_action"][:, :T].cpu().numpy().astype(bool), want["last_action"])
        np.testing.assert_array_equal(got["last_reward"][:, :T].cpu().numpy(), want["last_reward"])
        np.testing.assert_array_equal(got["hidden"].cpu().numpy(), want["hidden"])
        for k in ("action", "n_step_reward", "gamma"):
            np.testing.assert_array_equal(got[k][:rows].cpu().numpy(), want[k])
        for k in ("burn_in", "learning", "forward"):
            np.testing.assert_array_equal(got[k].cpu().numpy(), want[k])
        np.testing.assert_allclose(got["is_weights"][:rows].cpu().numpy(), want["is_weights"], rtol=1e-6)
        # priority update with the stale mask: same kept set as the oracle
        td = np.linspace(0.1, 2.0, bs).astype(np.float32)
        for old in (cpu.block_ptr, 2, (cpu.block_ptr + 1) % nb):
            keep = cpu.stale_mask(want["idxes"], old)
            before = dev.tree.ptree[cpu.tree.leaf_base:].copy()
            dev.update_priorities(idx, torch.from_numpy(td).cuda(), old)
            after = dev.tree.ptree[cpu.tree.leaf_base:]
            changed = np.zeros(after.shape, bool)
            changed[want["idxes"][keep]] = True
            np.testing.assert_array_equal(after[~changed], before[~changed])
            exp = (td[keep].astype(np.float32) ** np.float32(0.9)).astype(np.float64)
            last = {}
            for i, v in zip(want["idxes"][keep], exp):
                last[int(i)] = v
            for i, v in last.items():
                assert abs(after[i] - v) <= 3e-7 * max(1.0, v)
            dev.tree.set_leaves_device(torch.arange(1 << (L - 1), device="cuda"), torch.from_numpy(before).cuda())


def test_replay_buffer_surface_reproduces_reference_sample(golden_dir):
    """worker.ReplayBuffer.add / sample_batch (np.random.seed driven) against the reference's recorded batch."""
    from r2d2_b200 import config
    from r2d2_b200.worker import LocalBuffer, ReplayBuffer
    g = np.load(os.path.join(golden_dir, "replay_ragged.npz"))
    rb = ReplayBuffer([], None, None, buffer_capacity=8 * 400, batch_size=8)
    for seed, steps, done in synth.RAGGED_SCRIPT:
        lb = LocalBuffer(A)
        for blk, prio, ep in synth.drive_actor(lb, seed, steps, done, A):
            rb.add(blk, prio, ep)
    assert rb.block_ptr == int(g["block_ptr"]) and len(rb) == rb.size
    tree = rb.priority_tree
    want_tree = g["tree_after_add"]
    got_tree = tree.ptree
    np.testing.assert_allclose(got_tree, want_tree, rtol=1e-6)
    L = tree.num_layers
    tree.set_leaves_device(torch.arange(1 << (L - 1), device="cuda"), torch.from_numpy(want_tree[(1 << (L - 1)) - 1:].copy()).cuda())
    np.random.seed(7)
    data = rb.sample_batch()
    obs, la, lr_, hid, act, nsr, gam, b, l, f, idxes, isw, old_ptr, env_steps = data
    np.testing.assert_array_equal(idxes, g["s0_idxes"])                      # bit-exact sampled indices
    assert crc(obs.cpu().numpy()) == int(g["s0_obs_crc"])
    assert crc(la.cpu().numpy()) == int(g["s0_last_action_crc"])
    assert crc(np.ascontiguousarray(hid.cpu().numpy())) == int(g["s0_hidden_crc"])
    np.testing.assert_array_equal(lr_.cpu().numpy(), g["s0_last_reward"])
    np.testing.assert_array_equal(act.cpu().numpy(), g["s0_action"])
    np.testing.assert_array_equal(nsr.cpu().numpy(), g["s0_n_step_reward"])
    np.testing.assert_array_equal(gam.cpu().numpy(), g["s0_gamma"])
    for t, k in ((b, "burn_in"), (l, "learning"), (f, "forward")):
        np.testing.assert_array_equal(t.cpu().numpy(), g[f"s0_{k}"])
    np.testing.assert_allclose(isw.cpu().numpy(), g["s0_is_weights"], rtol=1e-6)
    assert old_ptr == int(g["s0_old_ptr"]) and int(env_steps) == int(g["s0_env_steps"])
    rb.update_priorities(g["upd_idx"], g["upd_td"], int(g["upd_gt_old_ptr"]), 0.5)
    np.testing.assert_allclose(rb.priority_tree.ptree, g["upd_gt_tree"], rtol=1e-6)
    assert rb.training_steps == 1 and rb.sum_loss == 0.5


def test_fused_gather_into_s2d_staging_matches_unfused():
    """r2d2_replay_gather_s2d (frames written straight into conv1's space-to-depth staging buffer, obs == NULL in the
    forward call) must give bit-identical learner outputs to gather -> raw frames -> s2d pass."""
    from oracle.learner import init_params
    from r2d2_b200.learner_core import DeviceLearner
    dev, cpu = _build_device_replay(synth.RAGGED_SCRIPT, 8, 8)
    core = DeviceLearner(A, 8, 85)
    core.load_state_dict(init_params(A, seed=3))
    r = torch.from_numpy(np.random.RandomState(5).random_sample(8)).cuda()
    b1, idx1, _ = dev.sample(r)
    core.compute_forward(b1)
    q1, td1 = core.q.clone(), core.td.clone()
    b1["obs"].zero_()                                          # make sure the fused path cannot read stale raw frames
    b2, idx2, _ = dev.sample(r, fuse_into=core)
    assert b2["obs"] is None and torch.equal(idx1, idx2)
    core.compute_forward(b2)
    torch.cuda.synchronize()
    assert torch.equal(core.q, q1) and torch.equal(core.td, td1)
