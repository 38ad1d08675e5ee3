"""Per-phase timing of the persistent recurrence kernel (CTA 0), from globaltimer stamps."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from r2d2_b200 import _lib
from r2d2_b200.learner_core import DeviceLearner
from oracle import synth
from oracle.learner import init_params

A, B, T = 9, 64, 85
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
names = ["poll->acquired", "acquired->staged", "staged->mma done", "mma done->epilogue done", "epilogue->released"]
d = np.diff(tr[:, :6], axis=1)[5:80]
step = np.diff(tr[:, 0])[5:80]
print("step period ns: mean %.0f  p50 %.0f" % (step.mean(), np.median(step)))
for i, n in enumerate(names):
    print(f"{n:26s} mean {d[:, i].mean():7.0f} ns   p50 {np.median(d[:, i]):7.0f}")
