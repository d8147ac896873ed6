#!/usr/bin/env python
"""Multi-GPU correctness check, launched with torchrun (one rank per GPU):
  * SHARD_TRIALS: the all-reduced cost sums of G ranks equal the single-GPU sums (Philox counters use
    the global trial index, so the draws do not depend on G) to 1e-12, costs to 1e-12, argmax identical;
  * SHARD_CANDIDATES: bit-identical (every slot is written by one rank, the others add 0.0).
  * the cost sums meet over peer memory (fused into the rollout kernel / as a kernel behind it) or in an
    ncclAllReduce: the three give the same numbers (bit for bit between the two peer-memory forms, and against NCCL
    where the sum has one order: candidates sharded, or two ranks), also over a run of back-to-back evaluations
    (both generations of the exchange buffers) and for batches with fewer candidates than ranks;
  * a rank that drops out of an evaluation: KFBO_EPEER on the others after the timeout, then KFBO_ESTATE everywhere.
Prints 'multi_gpu_check ok' on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kf_bayesopt_b200 as kf
from kf_bayesopt_b200 import _lib
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
res, paths = {}, {}
for mode in (kf.SHARD_TRIALS, kf.SHARD_CANDIDATES):
    with kf.CostEvaluator(rv, times, seed=seed, max_candidates=C, device=local) as ev:
        peer = init_comm(ev, mode)
        out = ev.eval_batch(q, r)
        res[mode] = (ev.last_sums(C).copy(), np.stack([o.cost for o in out]), [o.index_max for o in out])
        paths[mode] = [ev.last_reduce_path()]
        one = ev.eval([1.0], [0.1])     # a second evaluation on the same communicator
        res[mode] += (one.cost.copy(),)
        paths[mode].append(ev.last_reduce_path())
        # the same two evaluations again (same Philox streams) with every way the sums can meet
        variants = {}
        settings = [("nccl", _lib.REDUCE_NCCL, _lib.PEER_FUSE_AUTO)]
        if peer:
            settings += [("peer", _lib.REDUCE_AUTO, _lib.PEER_FUSE_AUTO), ("peer-never-fused", _lib.REDUCE_AUTO, _lib.PEER_FUSE_NEVER),
                         ("peer-always-fused", _lib.REDUCE_AUTO, _lib.PEER_FUSE_ALWAYS)]
        for name, red, fuse in settings:
            ev.set_option(_lib.OPT_REDUCE, red)
            ev.set_option(_lib.OPT_PEER_FUSE, fuse)
            ev.set_eval_counter(0)
            ev.eval_batch(q, r)
            a = ev.last_sums(C).copy()
            pa = ev.last_reduce_path()
            ev.eval([1.0], [0.1])
            variants[name] = (a, ev.last_sums(1).copy(), pa, ev.last_reduce_path())
        for name, (a, b, pa, pb) in variants.items():
            np.testing.assert_allclose(a, res[mode][0], rtol=1e-12, err_msg=name)
            if name.startswith("peer"):
                assert np.array_equal(a, variants["peer"][0]) and np.array_equal(b, variants["peer"][1]), name
                assert "peer" in pa and "peer" in pb, (name, pa, pb)
            if mode == kf.SHARD_CANDIDATES or world == 2:
                assert np.array_equal(a, variants["nccl"][0]) and np.array_equal(b, variants["nccl"][1]), name
        if peer:
            if mode == kf.SHARD_TRIALS:   # every rank has work in every evaluation, so the path is the option's
                assert "fused" in variants["peer-always-fused"][2] and "fused" in variants["peer-always-fused"][3]
                assert "behind" in variants["peer-never-fused"][3] and "behind" in variants["peer"][2] and "behind" in variants["peer"][3]
            # a run of back-to-back single-candidate evaluations: both generations of the buffers, no host barrier between
            ev.set_option(_lib.OPT_REDUCE, _lib.REDUCE_AUTO)
            ev.set_option(_lib.OPT_PEER_FUSE, _lib.PEER_FUSE_AUTO)
            ev.set_eval_counter(100)
            run_peer = [ev.eval([0.5 + 0.01 * i], [0.1]).cost.copy() for i in range(40)]
            # the launch-path features one at a time -- graph replay off, host polling off: neither touches the arithmetic,
            # so the same streams give the same bits
            for opt, val in ((_lib.OPT_CUDA_GRAPH, 0), (_lib.OPT_HOST_POLL, 0)):
                ev.set_option(opt, val)
                ev.set_eval_counter(100)
                again = [ev.eval([0.5 + 0.01 * i], [0.1]).cost.copy() for i in range(12)]
                assert all(np.array_equal(a, b) for a, b in zip(again, run_peer)), (opt, val)
            ev.set_option(_lib.OPT_CUDA_GRAPH, 1)
            ev.set_option(_lib.OPT_HOST_POLL, 1)
            ev.set_eval_counter(140)
            # fewer candidates than ranks: ranks without a candidate (or with one) still take part in the exchange
            few_peer = ev.eval_batch_raw(q[:1], r[:1])["cost"].copy()
            ev.set_option(_lib.OPT_REDUCE, _lib.REDUCE_NCCL)
            ev.set_eval_counter(100)
            run_nccl = [ev.eval([0.5 + 0.01 * i], [0.1]).cost.copy() for i in range(40)]
            ev.set_eval_counter(140)
            few_nccl = ev.eval_batch_raw(q[:1], r[:1])["cost"].copy()
            np.testing.assert_allclose(np.stack(run_peer), np.stack(run_nccl), rtol=1e-10, atol=1e-13)
            np.testing.assert_allclose(few_peer, few_nccl, rtol=1e-10, atol=1e-13)
# the state-augmented variants (both layouts) shard the same way: trials-sharded sums == one GPU's to rounding
aug = {}
for N in (3, 4):
    rva = kf.ReadParaVehicle(nsimruns_=1031, cpu_core_number_=1, state_dof_=N)
    for layout in (_lib.AUG_LAYOUT_THREAD, _lib.AUG_LAYOUT_LANES):
        with kf.CostEvaluator(rva, times, seed=seed, max_candidates=3, device=local, state_dim=N) as ev:
            ev.set_option(_lib.OPT_AUG_LAYOUT, layout)
            init_comm(ev, kf.SHARD_TRIALS)
            ev.eval_batch(q[:3], r[:3])
            aug[(N, layout)] = ev.last_sums(3).copy()
        if rank == 0:
            with kf.CostEvaluator(rva, times, seed=seed, max_candidates=3, device=local, state_dim=N) as ev1:
                ev1.set_option(_lib.OPT_AUG_LAYOUT, layout)
                ev1.eval_batch(q[:3], r[:3])
                np.testing.assert_allclose(aug[(N, layout)], ev1.last_sums(3), rtol=1e-12, err_msg=f"aug N={N} layout={layout}")
# a rank dropping out of an evaluation (fault injected before its launch): it reports its own error, the others KFBO_EPEER
# after the timeout instead of hanging, and every rank refuses further evaluations on that communicator (KFBO_ESTATE)
import time
with kf.CostEvaluator(rv, times, seed=seed, max_candidates=C, device=local) as ev:
    if init_comm(ev, kf.SHARD_TRIALS):
        ev.set_option(_lib.OPT_PEER_TIMEOUT_MS, 300)
        ev.eval([1.0], [0.1])
        if rank == world - 1:
            ev.set_option(_lib.OPT_INJECT_FAULT, 1)
        t = time.perf_counter()
        try:
            ev.eval([1.0], [0.1])
            code = 0
        except kf.KfboError as e:
            code = e.code
        assert time.perf_counter() - t < 5.0, "the exchange did not time out"
        assert code == (1 if rank == world - 1 else 3), (rank, code)      # KFBO_ECUDA on the rank that failed, KFBO_EPEER on the others
        try:
            ev.eval([1.0], [0.1])
            code = 0
        except kf.KfboError as e:
            code = e.code
        assert code == -4, (rank, code)                                    # KFBO_ESTATE
    dist.barrier()
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
    print(f"multi_gpu_check ok: {world} ranks (state-augmented N = 3, 4, both layouts, included), trials-sharded max rel diff "
          f"{np.max(np.abs(s_t - sums) / np.abs(sums)):.2e}, candidates-sharded bit-identical; the sums met by: {paths}", flush=True)
dist.barrier()
dist.destroy_process_group()
