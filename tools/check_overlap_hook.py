"""torchrun --nproc-per-node 2 tools/check_overlap_hook.py : learner updates with the plain gradient hook, with the overlapped hook
(dense-layer all-reduce issued mid-backward) and with the overlapped hook between two CUDA graphs (gradients / optimizer) from
identical states must give identical parameters."""
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
for mode in ("plain", "overlap", "overlap+graphs"):
    core = DeviceLearner(A, B, 85, in_channels=C, device=torch.device("cuda", local))
    core.use_graph = mode == "overlap+graphs"            # two CUDA graphs (gradients, optimizer) around the eager hook
    core.load_state_dict(init_params(A, in_channels=C, seed=0))
    core.grad_hook = r2dist.make_grad_hook() if mode == "plain" else r2dist.make_overlapped_grad_hook(core)
    d = synth.synthetic_batch(B, A, channels=C, seed=50 + rank, ragged=True)
    batch = {k: (torch.from_numpy(v) if hasattr(v, "dtype") and not isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    prepared = core.prepare(batch)                        # the same device buffers every time: the third update replays the graphs
    for _ in range(4):
        core.update(prepared)
    torch.cuda.synchronize()
    res.append(core.online.flat.clone())
same = torch.equal(res[0], res[1]) and torch.equal(res[0], res[2])
other = res[1].clone()
dist.broadcast(other, src=0)
ranks_agree = torch.equal(other, res[1])
print(f"rank {rank}: overlapped == plain: {same}; ranks agree: {ranks_agree}; max |dp| {float((res[0] - res[1]).abs().max()):.3e} "
      f"(graphs: {float((res[0] - res[2]).abs().max()):.3e})", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if same and ranks_agree else 1)
