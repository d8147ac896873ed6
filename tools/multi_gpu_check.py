#!/usr/bin/env python
"""Multi-GPU correctness check, launched with torchrun (one rank per GPU):
  * SHARD_TRIALS: the all-reduced cost sums of G ranks equal the single-GPU sums (Philox counters use
    the global trial index, so the draws do not depend on G) to 1e-12, costs to 1e-12, argmax identical;
  * SHARD_CANDIDATES: bit-identical (every slot is written by one rank, the others add 0.0).
Prints 'multi_gpu_check ok' on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kf_bayesopt_b200 as kf
from kf_bayesopt_b200.dist import init_comm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
times = [(0.1, 201), (0.5, 211)]
B, C = 4099, 13                         # neither divides by the rank count
rv = kf.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
q = np.linspace(0.1, 5.0, C)
r = np.linspace(0.01, 1.0, C)
seed = 77
res = {}
for mode in (kf.SHARD_TRIALS, kf.SHARD_CANDIDATES):
    with kf.CostEvaluator(rv, times, seed=seed, max_candidates=C, device=local) as ev:
        init_comm(ev, mode)
        out = ev.eval_batch(q, r)
        res[mode] = (ev.last_sums(C).copy(), np.stack([o.cost for o in out]), [o.index_max for o in out])
        one = ev.eval([1.0], [0.1])     # a second evaluation on the same communicator
        res[mode] += (one.cost.copy(),)
if rank == 0:
    with kf.CostEvaluator(rv, times, seed=seed, max_candidates=C, device=local) as ev:
        out = ev.eval_batch(q, r)
        sums, cost, idx = ev.last_sums(C), np.stack([o.cost for o in out]), [o.index_max for o in out]
        one = ev.eval([1.0], [0.1]).cost
    s_t, c_t, i_t, o_t = res[kf.SHARD_TRIALS]
    np.testing.assert_allclose(s_t, sums, rtol=1e-12)
    np.testing.assert_allclose(c_t, cost, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(o_t, one, rtol=1e-10, atol=1e-13)
    assert i_t == idx
    s_c, c_c, i_c, o_c = res[kf.SHARD_CANDIDATES]
    assert np.array_equal(s_c, sums) and np.array_equal(c_c, cost) and i_c == idx and np.array_equal(o_c, one)
    print(f"multi_gpu_check ok: {world} ranks, trials-sharded max rel diff "
          f"{np.max(np.abs(s_t - sums) / np.abs(sums)):.2e}, candidates-sharded bit-identical", flush=True)
dist.barrier()
dist.destroy_process_group()
