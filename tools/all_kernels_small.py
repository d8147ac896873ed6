#!/usr/bin/env python
"""Small invocation of every kernel of libkfbo.so (a target for compute-sanitizer where the pool allows it -- this one
does not: "compute-sanitizer is closed on this pool" -- and a quick all-paths check otherwise): the Philox rollout
(one candidate / a batch with device-built constants and device-side cost assembly, with and without per-trial outputs),
both supplied-noise kernels (cp.async.bulk + mbarrier ring, per-thread loads; ragged last tile), the single-trial trace,
and the state-augmented kernels (N = 3, 4; Philox and supplied noise)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kf_bayesopt_b200 as kf
from kf_bayesopt_b200 import _lib as L

rng = np.random.default_rng(3)
times = [(0.1, 37), (0.5, 42)]
rv = kf.ReadParaVehicle(nsimruns_=300, cpu_core_number_=1)
with kf.CostEvaluator(rv, times, seed=11, max_candidates=5) as ev:
    a = ev.eval([1.0], [0.1])
    b = ev.eval_batch_raw(np.linspace(0.2, 3.0, 5), np.linspace(0.02, 0.5, 5))
    c, mx = ev.eval_batch_costs(np.linspace(0.2, 3.0, 5), np.linspace(0.02, 0.5, 5))
    pt = ev.eval_philox_trials(1, 0.7, 0.2, 3, 17, 301)
    assert np.all(np.isfinite(a.cost)) and np.all(np.isfinite(b["cost"][:, :2])) and np.all(np.isfinite(c)) and np.all(np.isfinite(pt))
    tr = ev.trace_trial(0, 1.0, 0.1, 0, 5)
    assert tr.shape == (37, 9) and np.all(np.isfinite(tr))
    for B in (300, 258, 131):                                    # whole tiles + ragged tile (even: bulk kernel), odd: per-thread loads
        for k, (dt, T) in enumerate(times):
            x0n, w, v = rng.normal(size=(2, B)), rng.normal(size=(T, 2, B)) * 0.1, rng.normal(size=(T, B))
            for variant in (L.SUPPLIED_AUTO, L.SUPPLIED_LDG):
                ev.set_option(L.OPT_SUPPLIED_KERNEL, variant)
                per, sums = ev.eval_supplied(k, [1.0, 0.4], [0.1, 0.3], x0n, w, v)
                assert np.all(np.isfinite(per)) and np.all(np.isfinite(sums))
for N in (3, 4):
    rvn = kf.ReadParaVehicle(nsimruns_=77, cpu_core_number_=1, state_dof_=N)
    with kf.CostEvaluator(rvn, times, seed=11, max_candidates=3, state_dim=N) as ev:
        a = ev.eval([1.0], [0.1])
        b = ev.eval_batch_raw([0.5, 1.0, 2.0], [0.1, 0.1, 0.2])
        pt = ev.eval_philox_trials(1, 0.7, 0.2, 3, 17, 50)
        B, (dt, T) = 45, times[0]
        per, sums = ev.eval_supplied(0, [1.0, 0.4], [0.1, 0.3], rng.normal(size=(N, B)), rng.normal(size=(T, N, B)) * 0.1, rng.normal(size=(T, B)))
        assert np.all(np.isfinite(a.cost)) and np.all(np.isfinite(b["cost"][:, :2])) and np.all(np.isfinite(pt)) and np.all(np.isfinite(per))
print("all_kernels_small ok")
