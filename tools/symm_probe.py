"""torchrun --nproc-per-node 2 tools/symm_probe.py : does symmetric memory (peer pointers, multicast) rendezvous on this box?"""
import os
import sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as sm
try:
    t = sm.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", local))
    h = sm.rendezvous(t, dist.group.WORLD)
    print(rank, "symm ok: ptrs", [hex(p) for p in h.buffer_ptrs], "multicast", h.has_multicast_support(local and 0 or 0) if False else None,
          "mc_ptr", hex(h.multicast_ptr) if h.multicast_ptr else None, "signal", [hex(p) for p in h.signal_pad_ptrs], "pad size", h.signal_pad_size, flush=True)
    t.fill_(rank + 1)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer value", float(peer[0]), flush=True)
    h.barrier()
except Exception as e:                                     # noqa: BLE001
    print(rank, "symm FAILED:", repr(e)[:400], flush=True)
try:
    print(rank, "can access peer:", torch.cuda.can_device_access_peer(local, (local + 1) % world), flush=True)
    x = torch.full((1024,), float(rank + 1), device="cuda")
    hd = x.untyped_storage()._share_cuda_()
    objs = [None] * world
    dist.all_gather_object(objs, hd)
    o = objs[(rank + 1) % world]
    st = torch.UntypedStorage._new_shared_cuda(*o)
    y = torch.empty(0, dtype=torch.float32, device=torch.device("cuda", o[0])).set_(st, 0, (1024,))
    print(rank, "ipc peer value", float(y[0].item()), "ptr", hex(y.data_ptr()), "dev", y.device, flush=True)
except Exception as e:                                     # noqa: BLE001
    print(rank, "ipc FAILED:", repr(e)[:400], flush=True)
dist.barrier()
dist.destroy_process_group()
