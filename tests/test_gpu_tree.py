"""K3 GPU sum tree vs the oracle / reference golden vectors (runs on the B200 box)."""
import os

import numpy as np
import pytest

from oracle.gen_golden import tree_script
from oracle.sumtree import SumTreeOracle

pytestmark = pytest.mark.gpu


def _ulp_close_f32_leaves(got_f64, want_f64, max_ulp=2):
    g, w = got_f64.astype(np.float32), want_f64.astype(np.float32)
    assert np.array_equal(g.astype(np.float64), got_f64), "leaves must be float32-representable"
    d = np.abs(g.view(np.int32).astype(np.int64) - w.view(np.int32).astype(np.int64))
    assert d.max(initial=0) <= max_ulp, f"leaf pow differs by {d.max()} ulp"


def _check_internal_sums(nodes, L):
    """Every ancestor that was touched equals left+right in float64 (bit-exact)."""
    n_int = (1 << (L - 1)) - 1
    i = np.arange(n_int)
    np.testing.assert_array_equal(nodes[i], nodes[2 * i + 1] + nodes[2 * i + 2])


@pytest.mark.parametrize("tag,cap", [("p2", 1024), ("np2", 1000), ("tiny", 3)])
def test_tree_golden_script(golden_dir, tag, cap):
    import torch
    from r2d2_b200.priority_tree import PriorityTree
    g = np.load(os.path.join(golden_dir, "tree_small.npz"))
    gpu = PriorityTree(cap, 0.9, 0.6)
    cpu = SumTreeOracle(cap, 0.9, 0.6)
    assert gpu.num_layers == int(g[f"{tag}_num_layers"])
    for k, op in enumerate(tree_script(cap, 5)):
        if op[0] == "update":
            gpu.update(op[1], op[2])
            cpu.update(op[1], op[2])
            nodes = gpu.ptree
            # update path: leaf = float32 pow; CUDA and NumPy pow may differ in the last ulp
            _ulp_close_f32_leaves(nodes[cpu.leaf_base:], g[f"{tag}_op{k}_tree"][cpu.leaf_base:])
            _check_internal_sums(nodes, gpu.num_layers)
            np.testing.assert_allclose(nodes, g[f"{tag}_op{k}_tree"], rtol=1e-6, atol=0)
            # then pin the GPU tree to the reference's exact leaves so sampling can be compared bit for bit
            gpu.set_leaves_device(torch.arange(1 << (gpu.num_layers - 1),
                                               device="cuda"),
                                  torch.from_numpy(g[f"{tag}_op{k}_tree"][cpu.leaf_base:].copy()).cuda())
            np.testing.assert_array_equal(gpu.ptree, g[f"{tag}_op{k}_tree"])
        else:
            np.random.seed(op[2])
            idx, w = gpu.sample(op[1])
            np.testing.assert_array_equal(idx, g[f"{tag}_op{k}_idx"])          # bit-exact indices
            np.testing.assert_allclose(w, g[f"{tag}_op{k}_isw"], rtol=1e-12)


@pytest.mark.parametrize("cap", [50_000, 1 << 16, 1 << 20])
@pytest.mark.parametrize("n", [64, 1024, 4096, 65536])
def test_tree_large_matches_oracle(cap, n):
    import torch
    from r2d2_b200.priority_tree import PriorityTree
    rng = np.random.default_rng(cap + n)
    gpu = PriorityTree(cap, 0.9, 0.6)
    cpu = SumTreeOracle(cap, 0.9, 0.6)
    leaves = rng.uniform(1e-3, 1.0, cap).astype(np.float32).astype(np.float64)
    cpu.set_leaves(np.arange(cap), leaves)
    gpu.set_leaves_device(torch.arange(cap, device="cuda"), torch.from_numpy(leaves).cuda())
    np.testing.assert_array_equal(gpu.ptree, cpu.ptree)
    for rep in range(2):
        r = rng.random(n)
        ic, wc = cpu.sample(n, r)
        ig, wg32, wg = gpu.sample_device(n, torch.from_numpy(r).cuda(), want_f64=True)
        np.testing.assert_array_equal(ig.cpu().numpy(), ic)
        np.testing.assert_allclose(wg.cpu().numpy(), wc, rtol=1e-12)
        np.testing.assert_allclose(wg32.cpu().numpy(), wc.astype(np.float32), rtol=1e-6)
        # update at the sampled (duplicate-bearing) indices with exact leaf values -> identical trees
        newleaf = rng.uniform(0, 2, n).astype(np.float32).astype(np.float64)
        cpu.set_leaves(ic, newleaf)
        gpu.set_leaves_device(ig, torch.from_numpy(newleaf).cuda())
        np.testing.assert_array_equal(gpu.ptree, cpu.ptree)
        # and through the float32-pow path: within 2 ulp of NumPy's leaves, sums exact
        td = rng.uniform(0, 2, n).astype(np.float32)
        cpu.update(ic, td)
        gpu.update_device(ig, torch.from_numpy(td).cuda())
        nodes = gpu.ptree
        _ulp_close_f32_leaves(nodes[cpu.leaf_base:], cpu.ptree[cpu.leaf_base:])
        _check_internal_sums(nodes, gpu.num_layers)
        cpu.set_leaves(np.arange(cap), nodes[cpu.leaf_base:cpu.leaf_base + cap])


def test_tree_stale_mask_matches_reference_golden(golden_dir):
    """worker.py:247-256 mask fused into the update kernel."""
    import torch
    from r2d2_b200.priority_tree import PriorityTree
    g = np.load(os.path.join(golden_dir, "replay_ragged.npz"))
    base = g["tree_after_add"]
    cap = 80
    L = 8
    leaf_base = (1 << (L - 1)) - 1
    for tag in ("eq", "gt", "lt"):
        gpu = PriorityTree(cap, 0.9, 0.6)
        assert gpu.num_layers == L
        gpu.set_leaves_device(torch.arange(1 << (L - 1), device="cuda"), torch.from_numpy(base[leaf_base:].copy()).cuda())
        np.testing.assert_array_equal(gpu.ptree, base)
        gpu.update_device(torch.from_numpy(g["upd_idx"]).cuda(), torch.from_numpy(g["upd_td"]).cuda(),
                          old_ptr=int(g[f"upd_{tag}_old_ptr"]), cur_ptr=int(g["block_ptr"]), seq_per_block=10)
        want = g[f"upd_{tag}_tree"]
        got = gpu.ptree
        _ulp_close_f32_leaves(got[leaf_base:], want[leaf_base:])
        np.testing.assert_allclose(got, want, rtol=1e-6)
        # untouched leaves stay bit-identical
        same = want[leaf_base:] == base[leaf_base:]
        np.testing.assert_array_equal(got[leaf_base:][same], base[leaf_base:][same])


def test_device_philox_sampling_is_stratified():
    import torch
    from r2d2_b200.priority_tree import PriorityTree
    cap, n = 4096, 512
    gpu = PriorityTree(cap, 0.9, 0.6, seed=1234)
    leaves = torch.ones(cap, dtype=torch.float64, device="cuda")
    gpu.set_leaves_device(torch.arange(cap, device="cuda"), leaves)
    i1, w1 = gpu.sample_device(n)
    i2, _ = gpu.sample_device(n)
    i1, i2 = i1.cpu().numpy(), i2.cpu().numpy()
    # uniform leaves: stratum k covers leaves [8k, 8k+8)
    assert np.all(i1 // 8 == np.arange(n)) and np.all(i2 // 8 == np.arange(n))
    assert not np.array_equal(i1, i2)
    np.testing.assert_allclose(w1.cpu().numpy(), 1.0)
