"""tcgen05.mma issue rate in SS mode (both operands in shared memory): cycles per M x N x 16 bf16 instruction."""
import sys
import torch
sys.path.insert(0, ".")
from r2d2_b200 import _lib
_lib.require_device()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
reps = 2000
print("M    N   mode(1=two accumulators, 2=A MN-major)  ctas  clk/MMA   floor(128*N/256)")
for ctas in (1, 148):
    for M in (128, 64):
        for N in (32, 64, 128, 256):
            for mode in (0, 1, 2):
                if mode == 1 and 2 * N > 512:
                    continue
                _lib.check(_lib.lib().r2d2_debug_mma_rate(M, N, reps, mode, ctas, _lib.ptr(out), _lib.stream_ptr()))
                torch.cuda.synchronize()
                print(f"{M:4d} {N:4d} {mode:3d} {ctas:5d} {out.item() / (4 * reps):8.1f} {128 * N / 256:8.1f}")
