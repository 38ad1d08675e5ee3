"""Does a row-shifted UMMA descriptor read the right rows of a swizzled smem buffer? (planning experiment for
halo-reusing 'window' convolutions)"""
import sys
import torch
sys.path.insert(0, ".")
from r2d2_b200 import _lib
_lib.require_device()
torch.manual_seed(0)
A = torch.randn(144, 64, device="cuda").bfloat16().contiguous()
B = torch.randn(32, 64, device="cuda").bfloat16().contiguous()
for mode in (0, 1):
    line = []
    for s in range(0, 12):
        D = torch.zeros(128, 32, device="cuda")
        _lib.check(_lib.lib().r2d2_debug_shift_probe(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), s, mode, _lib.stream_ptr()))
        torch.cuda.synchronize()
        ref = A[s:s + 128].float() @ B.float().t()
        err = (D - ref).abs().max().item() / ref.abs().max().item()
        line.append(f"{s}:{'ok' if err < 1e-2 else f'{err:.1e}'}")
    print(f"base_offset mode {mode}: " + " ".join(line))

# mode 2: A consumed MN-major (rows of the smem buffer are the reduction index), M = 64, K = 64, N = 32
line = []
for s in range(0, 12):
    D = torch.zeros(128, 32, device="cuda")
    _lib.check(_lib.lib().r2d2_debug_shift_probe(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), s, 2, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = A[s:s + 64].float().t() @ B.float().t()                       # [64 m][32 n]
    lanes = torch.tensor([(m & 15) + 32 * (m >> 4) for m in range(64)], device="cuda")
    err = (D[lanes] - ref).abs().max().item() / ref.abs().max().item()
    line.append(f"{s}:{'ok' if err < 1e-2 else f'{err:.1e}'}")
print("MN-major shifted (mode 2): " + " ".join(line))

# mode 2+d: M = 128 built from two overlapping MN-major atoms d lines apart (LBO = 128*d bytes): rows 0-63 = window at
# shift, rows 64-127 = window at shift + d  (pairs two kernel taps in one MMA)
for d in (1, 2, 9, 21):
    line = []
    for s in range(0, 12, 3):
        D = torch.zeros(128, 32, device="cuda")
        _lib.check(_lib.lib().r2d2_debug_shift_probe(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), s, 2 + d, _lib.stream_ptr()))
        torch.cuda.synchronize()
        ref = torch.cat([A[s:s + 64].float().t() @ B.float().t(), A[s + d:s + d + 64].float().t() @ B.float().t()])
        err = (D - ref).abs().max().item() / ref.abs().max().item()
        line.append(f"{s}:{'ok' if err < 1e-2 else f'{err:.1e}'}")
    print(f"paired taps, atom distance {d} lines: " + " ".join(line))
