#!/usr/bin/env python
"""Time the cfg3 evaluation (kernel ms via CUDA events inside libkfbo) for the library selected by
KFBO_LIB_PATH -- used to compare kernel build variants on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kf_bayesopt_b200 as kf

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
rv = kf.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
with kf.CostEvaluator(rv, [(0.1, T), (0.5, T)], seed=1) as ev:
    form = 0
    ms = []
    for i in range(6):
        r = ev.eval([1.0], [0.1])
        ms.append(ev.last_kernel_ms()[0])
    best = min(ms[2:])
    print(f"{os.environ.get('KFBO_LIB_PATH', 'default'):40s} form={form} B={B} T={T} kernel_ms={best:.3f} steps/s={2 * B * T / best * 1e3:.4e} nees={r.average_nees}")
