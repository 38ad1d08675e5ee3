#!/usr/bin/env python
"""Learner throughput benchmark (BASELINE.json metric: learner sequences/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--channels C] [--impl reference]

One "step" = one learner update of one batch of 64 synthetic replay sequences
(b/l/f = 40/40/5 -> 85 frames of C x 84 x 84 each): unroll(online) + unroll(target)
-> fused TD -> BPTT -> [NCCL gradient all-reduce] -> clip+Adam -> priority update.
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the
definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A = 9                      # MsPacman action count (train.py:22 takes it from the env)
B = 64                     # config.batch_size
BURN, LEARN, FWD = 40, 40, 5
T = BURN + LEARN + FWD
TREE_CAPACITY = 1 << 20    # BASELINE config #3


def flops_per_sequence(C: int) -> float:
    """Algorithmic FLOPs (2*MAC) of one sequence's unroll fwd (2 nets) + bwd, SURVEY.md 8(d)."""
    enc = 2 * (819_200 * C + 2_654_208 + 1_806_336 + 1_605_632)
    lstm = 2 * 4 * 512 * (512 + A + 1 + 512)
    head = 2 * (2 * 512 * 512 + 512 * A + 512)
    fwd = 2 * (T * enc + T * lstm) + 3 * LEARN * head
    enc_b = 2 * (2 * (2_654_208 + 1_806_336 + 1_605_632) + 819_200 * C) - 2 * 0   # dgrad+wgrad, no conv1 dgrad
    bwd = (BURN + LEARN) * (enc_b + 2 * lstm) + LEARN * 2 * head
    return float(fwd + bwd)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_batches(n: int, C: int, seed0: int = 0):
    from oracle import synth          # bench-only use of the synthetic generator (inputs, not compute)
    return [synth.synthetic_batch(B, A, BURN, LEARN, FWD, channels=C, seed=seed0 + i) for i in range(n)]


def to_pinned(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            t = torch.from_numpy(v)
            out[k] = t.pin_memory() if t.numel() else t
        else:
            out[k] = v
    return out


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path, timed on the host cores: the oracle
    port (oracle/learner.py -- a torch-CPU restatement with the reference's three-pass structure;
    the reference itself cannot travel to the GPU box).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import synth
    from oracle.learner import LearnerState, init_params, learner_update
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    C = args.channels
    params = init_params(A, in_channels=C, seed=0)
    st = LearnerState(online={k: v.clone() for k, v in params.items()}, target={k: v.clone() for k, v in params.items()})
    batches = [synth.to_torch_batch(b) for b in host_batches(2, C)]
    for w in range(max(1, min(args.warmup, 1))):
        learner_update(st, batches[w % 2])
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for k in range(steps):
        learner_update(st, batches[k % 2])
    dt = (time.perf_counter() - t0) / steps
    val = B / dt
    line = {"impl": "reference", "metric": "learner sequences/sec", "value": val, "unit": "sequences/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: batch {B}, b/l/f {BURN}/{LEARN}/{FWD}, {C}x84x84 u8 frames, A={A}", "channels": C},
            "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} full updates of batch {B} after 1 warm-up (oracle/learner.py, torch CPU fp32, {cores} threads)"},
            "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- our arm
def count_kernels(fn):
    """Count kernels launched by one call of fn (CUPTI via torch.profiler), split ours/others."""
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    ours = other = 0
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in ev.name.lower() and "memset" not in ev.name.lower():
            if "r2d2" in ev.name:
                ours += 1
            else:
                other += 1
    return ours, other


def run_ours(args):
    import torch.distributed as dist
    from r2d2_b200.learner_core import DeviceLearner
    from r2d2_b200.priority_tree import PriorityTree
    from oracle.learner import init_params     # seeded init only (numpy RNG), no compute

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    C = args.channels

    dl = DeviceLearner(A, B, T, in_channels=C, max_learning=LEARN, max_forward=FWD, device=dev)
    dl.load_state_dict(init_params(A, in_channels=C, seed=0))
    tree = PriorityTree(TREE_CAPACITY, 0.9, 0.6, device=dev, seed=rank)
    tree.update_device(torch.arange(TREE_CAPACITY, device=dev),
                       torch.rand(TREE_CAPACITY, device=dev) + 1e-3)

    if world > 1:
        rows_g = torch.zeros(1, device=dev)

        def hook(l):
            # data-parallel exchange: SUM of d(loss_sum) and of the row counts, then one global mean
            dist.all_reduce(l.grads.flat)
            rows_g.copy_(l.rows)
            dist.all_reduce(rows_g)
            torch.reciprocal(rows_g, out=l.grad_scale)
        dl.grad_hook = hook

    nres = 4
    hb = host_batches(nres, C, seed0=100 * rank)
    resident = [dl.prepare({k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in b.items()}) for b in hb]
    pinned = [to_pinned(b) for b in hb]
    in_bytes = sum(v.numel() * v.element_size() for v in resident[0].values() if isinstance(v, torch.Tensor))
    idx_res = [torch.randint(0, TREE_CAPACITY, (B,), device=dev) for _ in range(nres)]

    def step_resident(i):
        b = resident[i % nres]
        idx, isw = tree.sample_device(B)            # K3 sample (indices drive the priority update below)
        dl.update(b)
        tree.update_device(idx, dl.prio)            # K3 update with the new priorities

    prio_host = torch.empty(B, dtype=torch.float32).pin_memory()
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def step_e2e(i):
        hbk = pinned[i % nres]
        b = dl.prepare(hbk)                          # H2D of the 14-tuple payload from pinned host memory
        dl.update(b)
        prio_host.copy_(dl.prio, non_blocking=True)  # worker.py:357,369: priorities + loss back to the host
        loss_host.copy_(dl.loss_sum, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_host[0])

    def timed(fn, steps, warmup):
        for w in range(warmup):
            fn(w)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            fn(warmup + k)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    sampler = ClockSampler(local)
    sampler.start()
    ms_step = timed(step_resident, args.steps, args.warmup)
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    sampler.stop_flag = True

    # unroll-only timing (K1 fwd x2 + K1b bwd) on the launching stream for the roofline
    b0 = resident[0]
    def unroll_only(i):
        dl.forward(0, b0, dl.q, dl.qn_online)
        dl.forward(1, b0, None, dl.qn_target)
        dl.backward(dl.dq)
    ms_unroll = timed(unroll_only, max(3, args.steps // 2), 2) if world == 1 else None

    if rank == 0:
        ours, other = count_kernels(lambda: step_resident(0))
        peaks = measured_peaks()
        value = world * B / (ms_step * 1e-3)
        e2e = world * B / (ms_e2e * 1e-3)
        line = {"metric": "learner sequences/sec", "value": value, "unit": "sequences/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"configs[1]: 1xB200 learner, batch {B}/GPU, b/l/f {BURN}/{LEARN}/{FWD} (T={T}), "
                                       f"{C}x84x84 u8 frames, A={A}, sum tree 2^20", "channels": C, "global_batch": world * B,
                           "parallelism": f"dp{world}", "l2": f"inputs larger than L2: {nres} rotating resident batches "
                                                              f"({nres * in_bytes / 1e6:.0f} MB) + ~1 GB of streamed activations per step"},
                "clocks": sampler.summary(),
                "e2e": {"value": e2e, "unit": "sequences/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": in_bytes,
                        "d2h_bytes_per_step": B * 4 + 4},
                "gpu_launches": ours, "other_launches": other}
        if ms_unroll is not None:
            fl = flops_per_sequence(C) * B
            ach = fl / (ms_unroll * 1e-3) / 1e12
            line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                                "frac": ach / peaks["tflops"], "traffic": None, "peak_source": peaks["source"],
                                "kernel": "K1+K1b unroll group (forward online+target, backward)", "ms": ms_unroll,
                                "algorithmic_gflop_per_launch": fl / 1e9}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(C)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(C):
    from oracle import synth
    from oracle.learner import LearnerState, init_params, learner_update
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = init_params(A, in_channels=C, seed=0)
    st = LearnerState(online={k: v.clone() for k, v in params.items()}, target={k: v.clone() for k, v in params.items()})
    batches = [synth.to_torch_batch(b) for b in host_batches(2, C)]
    learner_update(st, batches[0])
    t0 = time.perf_counter()
    n = 2
    for k in range(n):
        learner_update(st, batches[k % 2])
    dt = (time.perf_counter() - t0) / n
    return {"value": B / dt, "unit": "sequences/s", "cores": cores, "kind": "port",
            "sample": f"{n} full updates of batch {B} after 1 warm-up (oracle/learner.py, torch CPU fp32, {cores} threads)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--channels", type=int, default=4, help="frame channels: 4 = BASELINE.json shape, 1 = reference obs_shape")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
