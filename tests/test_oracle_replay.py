"""Replay-side restatement (LocalBuffer.finish / ReplayBuffer.add / sample_batch /
update_priorities) vs golden vectors produced by the unmodified reference."""
import os

import numpy as np

from oracle import synth
from helpers import build_oracle_replay, sample_with_seed, crc


def _check_batch(g, prefix, d):
    assert crc(d["obs"]) == int(g[f"{prefix}obs_crc"])
    assert crc(d["last_action"]) == int(g[f"{prefix}last_action_crc"])
    # reference hands out hidden as (2,B,H) transpose view of stack (B,2,H): crc was of the (2,B,H) copy
    assert crc(np.ascontiguousarray(d["hidden"].transpose(1, 0, 2))) == int(g[f"{prefix}hidden_crc"])
    for name in ("last_reward", "n_step_reward", "gamma", "burn_in", "learning", "forward", "idxes",
                 "is_weights"):
        np.testing.assert_array_equal(d[name], g[f"{prefix}{name}"], err_msg=name)
    np.testing.assert_array_equal(d["action"], g[f"{prefix}action"].reshape(-1))
    assert d["old_ptr"] == int(g[f"{prefix}old_ptr"])
    assert d["env_steps"] == int(g[f"{prefix}env_steps"])


def test_blocks_and_sample_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "replay_ragged.npz"))
    rb, blocks = build_oracle_replay(synth.RAGGED_SCRIPT, 8, 8)
    assert rb.block_ptr == int(g["block_ptr"])
    for i, (blk, prio, ep) in enumerate(blocks):
        assert crc(blk.obs) == int(g[f"blk{i}_obs_crc"])
        assert crc(blk.last_action) == int(g[f"blk{i}_last_action_crc"])
        assert crc(blk.hidden) == int(g[f"blk{i}_hidden_crc"])
        np.testing.assert_array_equal(blk.last_reward, g[f"blk{i}_last_reward"])
        np.testing.assert_array_equal(blk.action, g[f"blk{i}_action"])
        np.testing.assert_array_equal(blk.n_step_reward, g[f"blk{i}_n_step_reward"])
        np.testing.assert_array_equal(blk.gamma, g[f"blk{i}_gamma"])
        np.testing.assert_array_equal(
            np.stack([blk.burn_in_steps, blk.learning_steps, blk.forward_steps]), g[f"blk{i}_steps"])
        np.testing.assert_array_equal(prio, g[f"blk{i}_prio"])
        assert (-1.0 if ep is None else ep) == float(g[f"blk{i}_ep"])
    np.testing.assert_array_equal(rb.tree.ptree, g["tree_after_add"])
    _check_batch(g, "s0_", sample_with_seed(rb, 7))


def test_stale_masked_priority_update(golden_dir):
    g = np.load(os.path.join(golden_dir, "replay_ragged.npz"))
    for tag in ("eq", "gt", "lt"):
        rb, _ = build_oracle_replay(synth.RAGGED_SCRIPT, 8, 8)
        rb.update_priorities(g["upd_idx"], g["upd_td"], int(g[f"upd_{tag}_old_ptr"]))
        np.testing.assert_array_equal(rb.tree.ptree, g[f"upd_{tag}_tree"])
