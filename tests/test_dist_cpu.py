"""world_size-2 gloo test of the sharding logic: each rank evaluates its shard's partial cost sums
(on the CPU, with the oracle standing in for the kernel -- it is the checker here, not the product),
the sums meet in one all-reduce, and the finalised costs equal the unsharded evaluation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

T, B, C_CAND, D = 41, 64, 6, 2
SAMPLE_TIMES = [(0.1, T), (0.5, T)]
Q = np.linspace(0.2, 4.0, C_CAND)
R = np.linspace(0.02, 0.9, C_CAND)


def _local_sums(o, mode, rank, world, kf):
    sums = np.zeros((C_CAND, D, 4))
    if mode == "trials":
        b, e = kf.shard_range(B, rank, world)
        cands = range(C_CAND)
    else:
        b, e = 0, B
        cands = range(*kf.shard_range(C_CAND, rank, world))
    for c in cands:
        for k, (dt, t) in enumerate(SAMPLE_TIMES):
            if e > b:
                sums[c, k] = o.eval_philox(99, c * D + k, dt, t, 1.0, 0.1, Q[c], R[c], b, e - b).sum(axis=1)
    return sums


def _worker(rank, world, port, mode, out_path):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kf_bayesopt_b200 as kf
    from kf_bayesopt_b200.dist import allreduce_sums_host
    from oracle import oracle as o
    total = allreduce_sums_host(_local_sums(o, mode, rank, world, kf))
    if rank == 0:
        np.save(out_path, total)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["trials", "candidates"])
def test_two_rank_sharding_matches_unsharded(tmp_path, kfbo, oracle, mode):
    out = str(tmp_path / "sums.npy")
    mp.spawn(_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    total = np.load(out)
    full = _local_sums(oracle, "trials", 0, 1, kfbo)
    if mode == "candidates":
        assert np.array_equal(total, full)            # x + 0 is exact: candidate sharding is bit-identical
    else:
        np.testing.assert_allclose(total, full, rtol=1e-13)
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
    params = kfbo.make_params(rv, SAMPLE_TIMES, max_candidates=C_CAND)
    for c in range(C_CAND):
        a, b = kfbo.finalize(params, total[c]), kfbo.finalize(params, full[c])
        np.testing.assert_allclose(a.cost, b.cost, rtol=1e-12)
        assert a.index_max == b.index_max
