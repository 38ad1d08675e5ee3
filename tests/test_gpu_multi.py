"""Two-GPU checks of the data-parallel exchange step (skipped on single-GPU boxes; the host logic is covered on CPU by
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
