#!/usr/bin/env python
"""Where the host time of one evaluation goes (kfbo_last_host_us) next to the kernel time, for the cfg3 and cfg4 shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kf_bayesopt_b200 as kf

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:          # torchrun: one rank per GPU, the evaluator's own NCCL communicator
    import torch
    import torch.distributed as dist
    from kf_bayesopt_b200.dist import init_comm
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

def show(name, ev, call, n=10):
    for _ in range(3):
        call()
    acc, wall, kms = np.zeros(5), 0.0, 0.0
    for _ in range(n):
        t = time.perf_counter(); call(); wall += time.perf_counter() - t
        acc += np.array(list(ev.last_host_us().values())); kms += ev.last_kernel_ms()[0]
    acc /= n
    print(f"[rank {rank}/{world}] {name}: wall {wall / n * 1e3:.3f} ms, kernel {kms / n:.3f} ms, host us: build {acc[0]:.1f} "
          f"enqueue {acc[1]:.1f} device_wait {acc[2]:.1f} cost_assembly {acc[3]:.1f}; all-reduce (device) {acc[4]:.1f} us", flush=True)

rv = kf.ReadParaVehicle(nsimruns_=1 << 20, cpu_core_number_=1)
with kf.CostEvaluator(rv, [(0.1, 1000), (0.5, 1000)], device=local) as ev:
    if world > 1:
        init_comm(ev, kf.SHARD_TRIALS)
    show("cfg3", ev, lambda: ev.eval([1.0], [0.1]))
n = 64
q = np.repeat(np.geomspace(0.1, 5.0, n), n); r = np.tile(np.geomspace(0.01, 1.0, n), n)
rv = kf.ReadParaVehicle(nsimruns_=1024, cpu_core_number_=1)
with kf.CostEvaluator(rv, [(0.1, 201), (0.5, 211)], max_candidates=n * n, device=local) as ev:
    if world > 1:
        init_comm(ev, kf.SHARD_CANDIDATES)
    show("cfg4", ev, lambda: ev.eval_batch_raw(q, r))
rv = kf.ReadParaVehicle(nsimruns_=200, cpu_core_number_=10)
with kf.CostEvaluator(rv, [(0.1, 201), (1.0, 201)], device=local) as ev:
    if world > 1:
        init_comm(ev, kf.SHARD_TRIALS)
    show("cfg1/2 (200 trials x 201 steps x 2)", ev, lambda: ev.eval([1.0], [0.1]), n=50)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
