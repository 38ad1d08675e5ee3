"""torch-CPU learner restatement vs golden vectors from the unmodified reference
``Learner.run`` (K consecutive updates: TD, priorities, loss, Q tensors,
gradients, post-Adam parameters).

Tolerances: the restatement evaluates the same fp32 math with a different op
grouping (explicit LSTM time loop instead of the packed ``nn.LSTM`` kernel), so
results agree to fp32 round-off, not bit-for-bit: 2e-5 absolute on Q/TD (values
are O(1)); the 1e-4 bar of the north star is applied to the CUDA path against
this oracle in the GPU tests.
"""
import os

import numpy as np
import pytest
import torch

from oracle import synth
from oracle.gen_golden import CFG0_SCRIPT
from oracle.learner import LearnerState, init_params, learner_update
from helpers import build_oracle_replay, sample_with_seed, A

CASES = {
    "learner_ragged.npz": (synth.RAGGED_SCRIPT,),
    "learner_cfg0.npz": (CFG0_SCRIPT,),
}


@pytest.mark.parametrize("name", list(CASES))
def test_learner_matches_reference_golden(golden_dir, name):
    torch.set_num_threads(8)
    g = np.load(os.path.join(golden_dir, name))
    batch_size, K, bl, ls, bi, fs, seed0, num_blocks = (int(x) for x in g["meta"])
    rb, _ = build_oracle_replay(CASES[name][0], num_blocks, batch_size, bl, ls, bi, fs)
    params = init_params(A, seed=3)
    st = LearnerState(online={k: v.clone() for k, v in params.items()},
                      target={k: v.clone() for k, v in params.items()})
    for k in range(K):
        d = sample_with_seed(rb, seed0 + k)
        np.testing.assert_array_equal(d["idxes"], g[f"k{k}_idxes"])
        np.testing.assert_array_equal(d["is_weights"], g[f"k{k}_is_weights"])
        out = learner_update(st, synth.to_torch_batch(d), max_forward=fs)
        np.testing.assert_allclose(out["qn_online"].numpy(), g[f"k{k}_out_qn_online"], atol=2e-5, rtol=0)
        np.testing.assert_allclose(out["qn_target"].numpy(), g[f"k{k}_out_qn_target"], atol=2e-5, rtol=0)
        np.testing.assert_allclose(out["q"].numpy(), g[f"k{k}_out_q"], atol=2e-5, rtol=0)
        np.testing.assert_allclose(out["td"], g[f"k{k}_out_td"], atol=2e-5, rtol=0)
        np.testing.assert_allclose(out["priorities"], g[f"k{k}_out_priorities"], atol=2e-5, rtol=0)
        assert abs(out["loss"] - float(g[f"k{k}_out_loss"])) < 1e-5 * max(1.0, abs(out["loss"]))
        for n, gr in out["grads"].items():
            ref_norm = float(g[f"k{k}_gradnorm_{n}"])
            assert abs(gr.double().norm().item() - ref_norm) <= 1e-3 * ref_norm + 1e-7, n
            np.testing.assert_allclose(gr.flatten()[:16].numpy(), g[f"k{k}_gradhead_{n}"],
                                       atol=1e-3 * max(ref_norm, 1e-4), rtol=1e-3)
        for n, p in st.online.items():
            np.testing.assert_allclose(p.flatten()[:16].numpy(), g[f"k{k}_phead_{n}"], atol=3e-6, rtol=0)
            assert abs(p.double().abs().sum().item() - float(g[f"k{k}_pabs_{n}"])) <= \
                2e-6 * p.numel() + 1e-6


def test_packed_lstm_baseline_mode_agrees_with_specification():
    """The fast (torch packed LSTM) mode used for the timed CPU baseline computes the same update."""
    import oracle.learner as ol
    d = synth.synthetic_batch(4, A, burn_in=6, learning=5, forward=3, seed=9, ragged=True)
    outs = {}
    for mode in ("loop", "packed"):
        ol.LSTM_MODE = mode
        try:
            params = init_params(A, seed=2)
            st = LearnerState(online={k: v.clone() for k, v in params.items()}, target={k: v.clone() for k, v in params.items()})
            outs[mode] = learner_update(st, synth.to_torch_batch(d), max_forward=3)
        finally:
            ol.LSTM_MODE = "loop"
    np.testing.assert_allclose(outs["packed"]["td"], outs["loop"]["td"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(outs["packed"]["q"].numpy(), outs["loop"]["q"].numpy(), atol=2e-6, rtol=0)
    for k in outs["loop"]["grads"]:
        np.testing.assert_allclose(outs["packed"]["grads"][k].numpy(), outs["loop"]["grads"][k].numpy(), atol=1e-6, rtol=1e-4)
