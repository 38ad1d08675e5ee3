#!/usr/bin/env python
"""BASELINE config #5: GPU sum-tree throughput sweep vs the CPU (NumPy) tree.

capacity 2^16..2^22 x n in {1K, 4K, 16K, 64K}: sample(n) and update(n) per call, GPU timed with CUDA events
(tree resident in HBM/L2, uniforms drawn on the device), CPU = oracle/sumtree.py (the reference algorithm,
priority_tree.py:4-45) single-threaded.  Identical leaves and identical uniforms: sampled indices are compared
bit for bit on every cell.  Writes one JSON line per cell + a roofline column (algorithmic bytes SURVEY 8d:
sample 8*L per draw + 12 out, update 8 + 24*(L-1) per index)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.sumtree import SumTreeOracle            # noqa: E402
from r2d2_b200.priority_tree import PriorityTree    # noqa: E402


def gpu_time(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters      # us


def main():
    rng = np.random.default_rng(0)
    out = []
    for logc in (16, 18, 20, 22):
        cap = 1 << logc
        gpu = PriorityTree(cap, 0.9, 0.6)
        cpu = SumTreeOracle(cap, 0.9, 0.6)
        L = cpu.num_layers
        leaves = rng.uniform(1e-3, 1.0, cap).astype(np.float32).astype(np.float64)
        cpu.set_leaves(np.arange(cap), leaves)
        gpu.set_leaves_device(torch.arange(cap, device="cuda"), torch.from_numpy(leaves).cuda())
        for n in (1024, 4096, 16384, 65536):
            r = rng.random(n)
            rg = torch.from_numpy(r).cuda()
            ic, _ = cpu.sample(n, r)
            ig, _ = gpu.sample_device(n, rg)
            exact = bool(np.array_equal(ig.cpu().numpy(), ic))
            td = torch.rand(n, device="cuda") + 1e-3
            t_s = gpu_time(lambda: gpu.sample_device(n))
            t_u = gpu_time(lambda: gpu.update_device(ig, td))
            tdc = td.cpu().numpy()
            t0 = time.perf_counter(); reps = 3
            for _ in range(reps):
                cpu.sample(n, r)
            c_s = (time.perf_counter() - t0) / reps * 1e6
            t0 = time.perf_counter()
            for _ in range(reps):
                cpu.update(ic, tdc)
            c_u = (time.perf_counter() - t0) / reps * 1e6
            bytes_s, bytes_u = n * (8 * L + 12), n * (8 + 24 * (L - 1))
            row = dict(capacity=cap, n=n, layers=L, indices_bit_exact=exact, gpu_sample_us=round(t_s, 1), gpu_update_us=round(t_u, 1),
                       cpu_sample_us=round(c_s, 1), cpu_update_us=round(c_u, 1), sample_speedup=round(c_s / t_s, 1),
                       update_speedup=round(c_u / t_u, 1), sample_GBps=round(bytes_s / t_s / 1e3, 2),
                       update_GBps=round(bytes_u / t_u / 1e3, 2))
            print(json.dumps(row))
            out.append(row)
            cpu.set_leaves(np.arange(cap), gpu.ptree[cpu.leaf_base:cpu.leaf_base + cap])
    return out


if __name__ == "__main__":
    main()
