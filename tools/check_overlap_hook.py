"""torchrun --nproc-per-node 2 tools/check_overlap_hook.py : one learner update with the plain gradient hook and one with the
overlapped hook (dense-layer all-reduce issued mid-backward) from identical states must give identical parameters."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth                                   # noqa: E402  (test infrastructure: synthetic batch + init only)
from oracle.learner import init_params                     # noqa: E402
from r2d2_b200 import dist as r2dist                       # noqa: E402
from r2d2_b200.learner_core import DeviceLearner           # noqa: E402

rank, world, local = r2dist.init_from_env("nccl")
torch.cuda.set_device(local)
A, C, B = 9, 1, 8
res = []
for mode in ("plain", "overlap"):
    core = DeviceLearner(A, B, 85, in_channels=C, device=torch.device("cuda", local))
    core.load_state_dict(init_params(A, in_channels=C, seed=0))
    core.grad_hook = r2dist.make_grad_hook() if mode == "plain" else r2dist.make_overlapped_grad_hook(core)
    d = synth.synthetic_batch(B, A, channels=C, seed=50 + rank, ragged=True)
    batch = {k: (torch.from_numpy(v) if hasattr(v, "dtype") and not isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    for _ in range(2):
        core.update(core.prepare(batch))
    torch.cuda.synchronize()
    res.append(core.online.flat.clone())
same = torch.equal(res[0], res[1])
other = res[1].clone()
dist.broadcast(other, src=0)
ranks_agree = torch.equal(other, res[1])
print(f"rank {rank}: overlapped == plain: {same}; ranks agree: {ranks_agree}; max |dp| {float((res[0] - res[1]).abs().max()):.3e}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if same and ranks_agree else 1)
