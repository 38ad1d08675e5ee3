"""Per-phase timing of the cluster BPTT kernel (CTA 0, thread 0)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from r2d2_b200 import _lib
from r2d2_b200.learner_core import DeviceLearner
from r2d2_b200.synthetic import init_state_dict, synthetic_batch

A, B, T = 9, 64, 85
dl = DeviceLearner(A, B, T)
dl.use_graph = False
dl.load_state_dict(init_state_dict(A, seed=0))
d = synthetic_batch(B, A, seed=1)
b = dl.prepare({k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in d.items()})
trace = torch.zeros(T * 8, dtype=torch.int64, device="cuda")
for _ in range(3):
    dl.compute_gradients(b)
_lib.check(_lib.lib().r2d2_debug_rec_trace_bwd(trace.data_ptr()))
dl.compute_gradients(b)
torch.cuda.synchronize()
_lib.lib().r2d2_debug_rec_trace_bwd(None)
tr = trace.view(T, 8).cpu().numpy().astype(np.int64)
sl = slice(5, 75)
print("step period ns: %.0f" % np.diff(tr[:, 0])[sl].mean())
for a, b_, name in ((0, 1, "start -> partials received + summed"), (1, 2, "cell backward + dgates staged (+proxy fence)"), (2, 3, "CTA barrier"),
                    (3, 4, "barrier -> accumulator ready (MMAs)"), (4, 5, "tmem load + st.async push"), ):
    print(f"{name:48s} {(tr[sl, b_] - tr[sl, a]).mean():7.0f} ns")
print("push -> next step start                          %7.0f ns" % (tr[6:76, 0] - tr[5:75, 5]).mean())
