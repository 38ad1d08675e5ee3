"""CPU oracle for the R2D2 learner hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, on the CPU, the algorithm of the reference's learner
inner loop (sample -> burn-in -> unroll -> n-step TD -> priority update).  Each
function cites the reference file:line it follows (paths relative to the
upstream repository root).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs -- as the checker or
the timed CPU baseline, never as the thing shipped.  Nothing under
``r2d2_b200/`` imports it; the product path fails loudly when the CUDA
extension is missing.

Parity pinning: the reference has no tests, golden vectors or fixtures for this
path, so the oracle is pinned against outputs of the
reference itself: ``oracle/gen_golden.py`` imports the unmodified reference in
the build container, drives its ``LocalBuffer -> ReplayBuffer -> Learner`` and
``PriorityTree`` and writes ``tests/golden/*.npz``; ``tests/test_oracle_*``
check the restatement against those fixtures (and live against the reference
when it is present).
"""
