#!/usr/bin/env python
"""Latency of one 'noise' -> 'cost' evaluation of small workloads with the latency form of the rollout (KFBO_OPT_SMALL_LAUNCH) on and
off, over a sweep of step counts -- separates the fixed cost of a launch from the cost per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kf_bayesopt_b200 as kf
from kf_bayesopt_b200 import _lib as L

for B in (200, 32, 2000):
    for T in (2, 51, 201, 801, 2000):
        row = []
        for small in (1, 0):
            rv = kf.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
            with kf.CostEvaluator(rv, [(0.1, T), (0.5, T)], seed=1) as ev:
                ev.set_option(L.OPT_SMALL_LAUNCH, small)
                for _ in range(20):
                    ev.eval([1.0], [0.1])
                ev.kernel_ms_total(reset=True)
                t = time.perf_counter()
                n = 200
                for _ in range(n):
                    ev.eval([1.0], [0.1])
                wall = (time.perf_counter() - t) / n * 1e6
                kms, cnt = ev.kernel_ms_total()
                row.append((wall, kms / cnt * 1e3))
        print(f"B={B:5d} T={T:5d}  small: wall {row[0][0]:7.1f} us kernel {row[0][1]:7.1f} us | plain: wall {row[1][0]:7.1f} us kernel {row[1][1]:7.1f} us", flush=True)
