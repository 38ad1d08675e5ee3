"""Checks of the data-parallel exchange step (the two-GPU one is skipped on single-GPU boxes; the host logic is covered on CPU by
tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_overlapped_hook_matches_plain_hook_on_two_gpus():
    """tools/check_overlap_hook.py under torchrun: the hook that starts the dense-layer all-reduce mid-backward (row count
    riding in a padding slot) must give bit-identical parameters to the plain single all-reduce, on every rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tools", "check_overlap_hook.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    # NCCL hooks (plain / overlapped / overlapped between two graphs) and our peer-memory kernels (multicast and P2P)
    assert r.stdout.count("overlapped == plain: True; peer kernels ok: True; IS factor ok: True; ranks agree: True") == 2, tail


def test_peer_exchange_kernels_two_emulated_ranks_on_one_gpu():
    """tools/check_dp_single_gpu.py: csrc/dp.cu's all-reduce (row-count slot, both channels, several epochs) and its
    importance-weight exchange with two ranks emulated on one GPU (own buffers, control blocks and streams; P2P path), in a
    child process because a rank that loses its peer traps."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_dp_single_gpu.py")], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "two emulated ranks: OK" in r.stdout, tail
