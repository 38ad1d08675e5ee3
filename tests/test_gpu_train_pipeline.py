"""End-to-end drop-in topology on a GPU box: forked CPU actors + forked stager + CUDA learner in the main process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_pipeline_runs_to_completion():
    cmd = [sys.executable, os.path.join(ROOT, "examples", "train.py"), "--synthetic-env", "--actors", "2", "--training-steps", "6",
           "--learning-starts", "400", "--buffer-capacity", "3200", "--batch-size", "8", "--log-interval", "2",
           "--save-interval", "1000000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "done: 6 updates" in r.stdout, tail
    assert "number of training steps:" in r.stdout and "buffer size:" in r.stdout, tail


def test_train_pipeline_with_batched_gpu_actors():
    """same topology with all actors behind one batched GPU inference process (worker.VectorActor, --gpu-actors)"""
    cmd = [sys.executable, os.path.join(ROOT, "examples", "train.py"), "--synthetic-env", "--actors", "4", "--gpu-actors", "--training-steps", "6",
           "--learning-starts", "400", "--buffer-capacity", "3200", "--batch-size", "8", "--log-interval", "2",
           "--save-interval", "1000000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "done: 6 updates" in r.stdout, tail


def test_reference_train_py_runs_unmodified_against_the_dropin_modules(tmp_path):
    """train.py:1-49 of the upstream repository, byte for byte (baseline/_ref copy), with dropin/ on PYTHONPATH: its
    `from worker import Learner, Actor, ReplayBuffer`, `from model import Network`, `from environment import create_env`
    and `import config` bind this package.  A smaller run is requested the way a user would edit config.py
    (R2D2_CONFIG_OVERRIDES); the emulator is absent on this box, so the synthetic environment is opted into."""
    import json
    train_py = os.path.join(ROOT, "baseline", "_ref", "train.py")
    if not os.path.isfile(train_py):
        train_py = "/root/reference/train.py"
    if not os.path.isfile(train_py):
        pytest.skip("reference train.py not present (baseline/_ref is created by __graft_entry__.build())")
    import filecmp
    import shutil
    script = tmp_path / "train.py"               # the script's own directory is sys.path[0]: it must not contain the reference's
    shutil.copyfile(train_py, script)            # worker.py / model.py, or those would shadow the drop-in modules
    assert filecmp.cmp(train_py, script, shallow=False)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT, env.get("PYTHONPATH", "")])
    env["R2D2_SYNTHETIC_ENV"] = "1"
    env["R2D2_CONFIG_OVERRIDES"] = json.dumps({"training_steps": 6, "learning_starts": 400, "buffer_capacity": 3200, "batch_size": 8,
                                               "log_interval": 2, "num_actors": 2, "save_interval": 1000000})
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, cwd=str(tmp_path), env=env)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "number of training steps: 6" in r.stdout, tail
