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


# ---------------------------------------------------------------------------------------------
# The attach protocol of the peer-memory exchange (dist.attach_peers): export -> all-gather -> attach -> agree.
# The CUDA side needs GPUs (tests/test_gpu_multi.py); here two gloo ranks run the protocol on a stand-in evaluator.
# ---------------------------------------------------------------------------------------------
class _FakeEvaluator:
    def __init__(self, rank, fail_export=False, fail_attach=False):
        self.rank, self.fail_export, self.fail_attach = rank, fail_export, fail_attach
        self.attached, self.options = None, []

    def peer_export(self):
        from kf_bayesopt_b200 import KfboError
        if self.fail_export:
            raise KfboError(1, "no IPC here")
        return bytes([self.rank]) * 64

    def peer_attach(self, handles):
        from kf_bayesopt_b200 import KfboError
        if self.fail_attach:
            raise KfboError(1, "cannot map the peer")
        self.attached = list(handles)

    def set_option(self, option, value):
        self.options.append((option, value))


def _attach_worker(rank, world, port, scenario, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kf_bayesopt_b200 import _lib
    from kf_bayesopt_b200.dist import attach_peers
    ev = _FakeEvaluator(rank, fail_export=(scenario == "export" and rank == 1), fail_attach=(scenario == "attach" and rank == 0))
    ok = attach_peers(ev, world)
    if scenario == "ok":
        assert ok and ev.attached == [bytes([r]) * 64 for r in range(world)] and ev.options == []
    else:       # one rank failed: EVERY rank falls back to the NCCL call
        assert not ok and ev.options == [(_lib.OPT_REDUCE, _lib.REDUCE_NCCL)]
    open(os.path.join(out_dir, f"done{rank}"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scenario", ["ok", "export", "attach"])
def test_peer_attach_protocol_agrees_across_ranks(tmp_path, scenario):
    mp.spawn(_attach_worker, args=(2, _free_port(), scenario, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["done0", "done1"]
