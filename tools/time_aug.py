#!/usr/bin/env python
"""Time the state-augmented evaluation (262,144 trials x 1000 steps x 2 sample times, on-device noise; kernel ms via CUDA
events inside libkfbo) for the library selected by KFBO_LIB_PATH, both layouts -- used to compare build variants."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kf_bayesopt_b200 as kf
from kf_bayesopt_b200 import _lib as L

B, T = 1 << 18, 1000
for N in (3, 4):
    rv = kf.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1, state_dof_=N)
    with kf.CostEvaluator(rv, [(0.1, T), (0.5, T)], seed=1, state_dim=N) as ev:
        out = []
        for name, layout in (("thread", L.AUG_LAYOUT_THREAD), ("lanes", L.AUG_LAYOUT_LANES)):
            ev.set_option(L.OPT_AUG_LAYOUT, layout)
            ms = []
            for i in range(5):
                ev.set_eval_counter(0)
                r = ev.eval([1.0], [0.1])
                ms.append(ev.last_kernel_ms()[0])
            out.append(f"{name} {min(ms[1:]):.3f} ms nees={r.average_nees}")
    print(f"{os.path.basename(os.environ.get('KFBO_LIB_PATH', 'default')):28s} N={N} " + " | ".join(out))
