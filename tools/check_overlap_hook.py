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
MODES = ("plain", "overlap", "overlap+graphs", "peer-multicast", "peer-p2p+graphs")
exchanges = []
for mode in MODES:
    core = DeviceLearner(A, B, 85, in_channels=C, device=torch.device("cuda", local))
    core.use_graph = mode.endswith("graphs")             # two CUDA graphs (gradients, optimizer) around the eager hook
    core.load_state_dict(init_params(A, in_channels=C, seed=0))
    if mode.startswith("peer"):                          # our own all-reduce kernels over NVLink peer memory (csrc/dp.cu)
        ex = r2dist.PeerExchange(core, use_multicast=mode == "peer-multicast")
        exchanges.append(ex)
        core.grad_hook = r2dist.make_peer_grad_hook(core, ex)
    else:
        core.grad_hook = r2dist.make_grad_hook() if mode == "plain" else r2dist.make_overlapped_grad_hook(core)
    d = synth.synthetic_batch(B, A, channels=C, seed=50 + rank, ragged=True)
    batch = {k: (torch.from_numpy(v) if hasattr(v, "dtype") and not isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    prepared = core.prepare(batch)                        # the same device buffers every time: the third update replays the graphs
    for _ in range(4):
        core.update(prepared)
    torch.cuda.synchronize()
    res.append(core.online.flat.clone())
# the importance-weight exchange of the peer kernels against the NCCL formulation (dist.global_is_factor), three rounds
from r2d2_b200 import _lib                                # noqa: E402
is_ok = True
ex = exchanges[0]
for rnd in range(3):
    g = torch.Generator().manual_seed(1000 * rnd + rank)
    nodes = (torch.rand(1 + 64, generator=g, dtype=torch.float64) + 0.01).cuda()
    nodes[0] = nodes[1:].sum()
    idx = torch.randint(0, 64, (8,), generator=g).cuda()
    w = torch.rand(40, generator=g).cuda()
    want = w * r2dist.global_is_factor((nodes[1 + idx].min() / nodes[0]).reshape(1), 0.6).to(torch.float32)
    s_ = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().r2d2_dp_is_post(ex._h, nodes.data_ptr(), 1, idx.data_ptr(), idx.numel(), s_))
    _lib.check(_lib.lib().r2d2_dp_is_apply(ex._h, 0.6, w.data_ptr(), w.numel(), None, s_))
    is_ok = is_ok and bool(torch.allclose(w, want, rtol=1e-6, atol=0))
# NCCL variants: bit-identical.  Peer kernels: bit-identical at world size 2 (a + b has one rounding whatever the order), within
# fp32 summation-order noise beyond that; all ranks must agree bit for bit in every mode.
diffs = [float((res[0] - r).abs().max()) for r in res]
same = all(torch.equal(res[0], r) if world == 2 else d < 1e-6 for r, d in zip(res[1:3], diffs[1:3]))   # NCCL re-orders sums beyond 2 ranks
peer_ok = all(torch.equal(res[0], r) if world == 2 else d < 1e-6 for r, d in zip(res[3:], diffs[3:]))
ranks_agree = True
for r in res:
    other = r.clone()
    dist.broadcast(other, src=0)
    ranks_agree = ranks_agree and torch.equal(other, r)
print(f"rank {rank}: overlapped == plain: {same}; peer kernels ok: {peer_ok}; IS factor ok: {is_ok}; ranks agree: {ranks_agree}; "
      f"max |dp| vs plain {dict(zip(MODES, diffs))}; multicast available: {exchanges[0].multicast}", flush=True)
same = same and peer_ok and is_ok
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if same and ranks_agree else 1)
