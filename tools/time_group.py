#!/usr/bin/env python
"""ms per evaluation of the cfg3 / cfg4 workloads with ONE process driving G GPUs (kfbo_group_*), G = 1, 2, 4, 8 as available:
    python tools/time_group.py [--steps 20]
One JSON line per (workload, G).  The multi-process numbers (torchrun, bench.py --gpus N) are the headline; this is the
single-process mode the ROS node uses."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import kf_bayesopt_b200 as kf

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--reduce", default="both", choices=["auto", "nccl", "both"])
ap.add_argument("--only", type=int, default=0, help="only this group size")
args = ap.parse_args()
from kf_bayesopt_b200 import _lib
ndev = torch.cuda.device_count()
n = 64
q = np.repeat(np.geomspace(0.1, 5.0, n), n); r = np.tile(np.geomspace(0.01, 1.0, n), n)
for G in [g for g in (1, 2, 4, 8) if g <= ndev and args.only in (0, g)]:
    for name, rv, times, mode, C, call in [
        ("cfg3", kf.ReadParaVehicle(nsimruns_=1 << 20, cpu_core_number_=1), [(0.1, 1000), (0.5, 1000)], kf.SHARD_TRIALS, 1,
         lambda ev: ev.eval([1.0], [0.1])),
        ("cfg4", kf.ReadParaVehicle(nsimruns_=1024, cpu_core_number_=1), [(0.1, 201), (0.5, 211)], kf.SHARD_CANDIDATES, n * n,
         lambda ev: ev.eval_batch_raw(q, r))]:
      for red, threads in ([(_lib.REDUCE_AUTO, 1), (_lib.REDUCE_AUTO, 0), (_lib.REDUCE_NCCL, 1)] if args.reduce == "both" and G > 1 else
                           [(_lib.REDUCE_NCCL if args.reduce == "nccl" else _lib.REDUCE_AUTO, 1)]):
        with kf.CostEvaluatorGroup(rv, times, n_gpus=G, shard_mode=mode, max_candidates=C) as ev:
            ev.set_option(_lib.OPT_REDUCE, red)
            ev.set_option(_lib.OPT_GROUP_THREADS, threads)     # one host thread per device (default) or the calling thread for all
            kms = []
            for _ in range(3):
                call(ev)
                kms.append(ev.last_kernel_ms()[0])          # the kernel time is read in the warm-up only: reading it costs 2 event queries per device
            t = time.perf_counter()
            for _ in range(args.steps):
                call(ev)
            ms = (time.perf_counter() - t) / args.steps * 1e3
            steps = float(rv.nsimruns_) * C * sum(T for _, T in times)
            print(json.dumps({"workload": name, "mode": "one process, device group", "n_gpus": G, "sums_meet_by": ev.last_reduce_path(), "host_threads": "one per device" if threads and G > 1 else "one", "ms_per_eval": ms,
                              "slowest_kernel_ms": float(np.mean(kms)), "filter_steps_per_s": steps / (ms * 1e-3)}), flush=True)
