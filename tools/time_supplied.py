#!/usr/bin/env python
"""Time the supplied-noise (HBM-bound) kernel for the library selected by KFBO_LIB_PATH:
kernel ms via CUDA events inside libkfbo, achieved algorithmic GB/s.
  time_supplied.py B T [pad_kb ...]   -- LDG variant once, bulk variant for every smem pad given"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kf_bayesopt_b200 as kf
from kf_bayesopt_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
pads = [int(a) for a in sys.argv[3:]] or [0]
g = torch.Generator(device="cuda").manual_seed(1)
x0n = torch.randn((2, B), dtype=torch.float64, device="cuda", generator=g)
w = torch.randn((T, 2, B), dtype=torch.float64, device="cuda", generator=g) * 0.1
v = torch.randn((T, B), dtype=torch.float64, device="cuda", generator=g)
rv = kf.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
byts = 24.0 * B * T + 48.0 * B
tag = os.path.basename(os.environ.get('KFBO_LIB_PATH', 'default'))
with kf.CostEvaluator(rv, [(0.1, T)]) as ev:
    ref = None
    for name, variant, pad in [("ldg", _lib.SUPPLIED_LDG, 0)] + [("bulk", _lib.SUPPLIED_BULK, p) for p in pads]:
        ev.set_option(_lib.OPT_SUPPLIED_KERNEL, variant)
        ev.set_option(_lib.OPT_BULK_SMEM_PAD, pad << 10)
        ms = []
        for i in range(8):
            s = ev.eval_supplied_dev(0, 1.0, 0.1, B, x0n.data_ptr(), w.data_ptr(), v.data_ptr())
            ms.append(ev.last_kernel_ms()[0])
        best = min(ms[3:])
        if ref is None:
            ref = s.copy()
        assert (ref == s).all(), "variants disagree"
        print(f"{tag:28s} {name:5s} pad={pad:3d}KB B={B} T={T} kernel_ms={best:.4f} GB/s={byts / best / 1e6:.1f} "
              f"steps/s={B * T / best * 1e3:.4e}", flush=True)
