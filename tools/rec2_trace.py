"""Per-phase timing of the cluster recurrence kernel (CTA 0, thread 0) + cluster capacity of the device."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from r2d2_b200 import _lib
from r2d2_b200.learner_core import DeviceLearner
from oracle import synth
from oracle.learner import init_params

A, B, T = 9, 64, 85
cap = _lib.lib().r2d2_debug_cluster_capacity()
print("max active 16-CTA clusters: NS=16: %d, NS=32: %d" % (cap // 100, cap % 100))
dl = DeviceLearner(A, B, T)
dl.load_state_dict(init_params(A, seed=0))
d = synth.synthetic_batch(B, A, seed=1)
b = dl.prepare({k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in d.items()})
trace = torch.zeros(T * 8, dtype=torch.int64, device="cuda")
for _ in range(3):
    dl.compute_forward(b)
_lib.check(_lib.lib().r2d2_debug_rec_trace(trace.data_ptr()))
dl.compute_forward(b)
torch.cuda.synchronize()
_lib.lib().r2d2_debug_rec_trace(None)
tr = trace.view(T, 8).cpu().numpy().astype(np.int64)
names = ["loop start->acc ready", "acc ready->cell done", "cell done->published"]
sl = slice(5, 80)
print("MMA warp: loop start(epi) -> first tile landed  %.0f ns" % (tr[sl, 6] - tr[sl, 0]).mean())
print("MMA warp: first tile -> all issued+commit       %.0f ns" % (tr[sl, 7] - tr[sl, 6]).mean())
print("MMA warp: commit -> acc ready seen by epilogue  %.0f ns" % (tr[sl, 1] - tr[sl, 7]).mean())
print("published(prev step) -> first tile landed       %.0f ns" % (tr[6:81, 6] - tr[5:80, 3]).mean())
dd = np.diff(tr[:, :4], axis=1)[5:80]
step = np.diff(tr[:, 0])[5:80]
print("step period ns: mean %.0f  p50 %.0f" % (step.mean(), np.median(step)))
for i, n in enumerate(names):
    print(f"{n:26s} mean {dd[:, i].mean():7.0f} ns   p50 {np.median(dd[:, i]):7.0f}")
# one network only (4 clusters) vs both (8 clusters)
import time
q = torch.zeros(dl.rows_cap, A, device="cuda")
for which, label in ((None, "both networks"),):
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for fn, label in ((lambda: dl.compute_forward(b), "forward_pair (8 clusters)"), (lambda: dl.forward(0, b, q, q), "forward online only (4 clusters)")):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(label, "%.3f ms" % (e0.elapsed_time(e1) / 5))
