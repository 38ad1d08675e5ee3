"""python tools/check_dp_single_gpu.py : the peer-memory exchange kernels (csrc/dp.cu) with TWO ranks emulated on ONE GPU.

Each "rank" is a handle with its own gradient buffer, control block and stream; the peers' pointers are simply the other
rank's buffers on the same device (P2P path; there is no multicast mapping inside one process).  The two all-reduce launches
barrier against each other from two streams, exactly as two GPUs would.  Checked: sums bit-exact against torch, both ranks
identical, row-count slot -> grad_scale and back to zero, several epochs, both channels, and the importance-weight factor
against dist.global_is_factor's formula."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2d2_b200 import _lib                                   # noqa: E402

lib = _lib.lib()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
W, n = 2, 4 * 50_001                                         # floats; slices of unequal length
torch.manual_seed(0)
grads = [torch.zeros(n, device=dev) for _ in range(W)]
ctl = [torch.zeros(lib.r2d2_dp_ctl_bytes() // 4, dtype=torch.int32, device=dev) for _ in range(W)]
rows = [torch.tensor([1000 + 37 * r], dtype=torch.int32, device=dev) for r in range(W)]
scale = [torch.zeros(1, device=dev) for _ in range(W)]
streams = [torch.cuda.Stream(device=dev) for _ in range(W)]
gp = (C.c_ulonglong * W)(*[g.data_ptr() for g in grads])
cp = (C.c_ulonglong * W)(*[c.data_ptr() for c in ctl])
handles = []
for r in range(W):
    h = C.c_void_p()
    _lib.check(lib.r2d2_dp_create(r, W, gp, 0, cp, C.byref(h)))
    handles.append(h)
torch.cuda.synchronize()

ok = True
slot = n - 1
for epoch in range(4):
    src = [torch.randn(n, device=dev) * (r + 1) for r in range(W)]
    for r in range(W):
        src[r][slot] = 0.0                                   # the padding slot is zero between updates
        grads[r].copy_(src[r])
    want = src[0] + src[1]
    torch.cuda.synchronize()
    half = 4 * 20_000
    for r in range(W):                                       # dense range with the row count on channel 0, the rest on channel 1
        with torch.cuda.stream(streams[r]):
            _lib.check(lib.r2d2_dp_allreduce(handles[r], half, n - half, 0, rows[r].data_ptr(), slot, scale[r].data_ptr(),
                                             8, 64, 0, streams[r].cuda_stream))
            _lib.check(lib.r2d2_dp_allreduce(handles[r], 0, half, 1, None, 0, None, 4, 256, 0, streams[r].cuda_stream))
    torch.cuda.synchronize()
    for r in range(W):
        ok &= bool(torch.equal(grads[r], want))
        ok &= float(scale[r]) == float(torch.tensor(1.0) / torch.tensor(2037.0)) and float(grads[r][slot]) == 0.0
    ok &= bool(torch.equal(grads[0], grads[1]))

# importance-weight factor: post on every rank, then apply
beta = 0.6
for rnd in range(3):
    g = torch.Generator().manual_seed(7 + rnd)
    nodes = [(torch.rand(65, generator=g, dtype=torch.float64) + 0.01).to(dev) for _ in range(W)]
    for t in nodes:
        t[0] = t[1:].sum()
    idx = [torch.randint(0, 64, (8,), generator=g).to(dev) for _ in range(W)]
    w = [torch.rand(40, generator=g).to(dev) for _ in range(W)]
    fac = [torch.zeros(1, device=dev) for _ in range(W)]
    local = [float(nodes[r][1 + idx[r]].min() / nodes[r][0]) for r in range(W)]
    want = [(local[r] / min(local)) ** -beta for r in range(W)]
    w0 = [t.clone() for t in w]
    for r in range(W):
        with torch.cuda.stream(streams[r]):
            _lib.check(lib.r2d2_dp_is_post(handles[r], nodes[r].data_ptr(), 1, idx[r].data_ptr(), 8, streams[r].cuda_stream))
            _lib.check(lib.r2d2_dp_is_apply(handles[r], beta, w[r].data_ptr(), 40, fac[r].data_ptr(), streams[r].cuda_stream))
    torch.cuda.synchronize()
    for r in range(W):
        ok &= abs(float(fac[r]) - want[r]) <= 1e-6 * want[r]
        ok &= bool(torch.allclose(w[r], w0[r] * float(fac[r]), rtol=1e-6, atol=0))
err = C.c_uint(0)
for h in handles:
    _lib.check(lib.r2d2_dp_error(h, C.byref(err)))
    ok &= err.value == 0
    lib.r2d2_dp_destroy(h)
print("dp kernels on one GPU, two emulated ranks:", "OK" if ok else "MISMATCH", flush=True)
sys.exit(0 if ok else 1)
