"""Multi-GPU test (needs >= 2 GPUs on the box, e.g. `gpurun --gpus 2`): one process per GPU under
torchrun, NCCL all-reduce of the cost sums inside libkfbo.so; sharded == unsharded."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_ranks_nccl_allreduce_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29731",
                          os.path.join(ROOT, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi_gpu_check ok" in out.stdout


# ---------------------------------------------------------------------------------------------
# Device group: ONE process driving several GPUs (kfbo_group_*, what the single-process ROS node uses)
# ---------------------------------------------------------------------------------------------
def test_group_of_one_device_equals_a_plain_handle(kfbo):
    import numpy as np
    rv = kfbo.ReadParaVehicle(nsimruns_=4099, cpu_core_number_=1)
    times, q, r = [(0.1, 201), (0.5, 211)], np.linspace(0.1, 5.0, 7), np.linspace(0.01, 1.0, 7)
    with kfbo.CostEvaluator(rv, times, seed=5, max_candidates=7) as ev, \
            kfbo.CostEvaluatorGroup(rv, times, n_gpus=1, seed=5, max_candidates=7) as gr:
        a, b = ev.eval_batch_raw(q, r), gr.eval_batch_raw(q, r)
        assert np.array_equal(ev.last_sums(7), gr.last_sums(7))
        assert np.array_equal(a["cost"], b["cost"]) and np.array_equal(a["index_max"], b["index_max"])
        one_a, one_b = ev.eval([1.0], [0.1]), gr.eval([1.0], [0.1])      # second evaluation: fresh streams on both
        assert np.array_equal(one_a.cost, one_b.cost)
    with pytest.raises(kfbo.KfboError):
        kfbo.CostEvaluatorGroup(rv, times, n_gpus=4096)                  # more devices than the box has


def test_device_group_matches_single_device(kfbo):
    import numpy as np
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    G = 2 if n < 4 else 4
    B, C = 4099, 13                                                     # neither divides by the device count
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
    times, q, r = [(0.1, 201), (0.5, 211)], np.linspace(0.1, 5.0, C), np.linspace(0.01, 1.0, C)
    with kfbo.CostEvaluator(rv, times, seed=77, max_candidates=C) as ev:
        want = ev.eval_batch_raw(q, r)
        want_sums = ev.last_sums(C)
        want_one = ev.eval([1.0], [0.1]).cost
    for mode in (kfbo.SHARD_TRIALS, kfbo.SHARD_CANDIDATES):
        with kfbo.CostEvaluatorGroup(rv, times, n_gpus=G, shard_mode=mode, seed=77, max_candidates=C) as gr:
            got = gr.eval_batch_raw(q, r)
            sums = gr.last_sums(C)
            one = gr.eval([1.0], [0.1]).cost                            # a second evaluation on the same communicators
            ms, launches = gr.last_kernel_ms()
            assert ms > 0 and 1 <= launches <= 2 * G                    # one candidate, candidate-sharded: one device has work
            path = gr.last_reduce_path()
            # the same two evaluations with the sums meeting in an ncclAllReduce instead of peer memory (or again in
            # NCCL on a box without peer access): the same numbers -- bit for bit where the sum has one order
            from kf_bayesopt_b200 import _lib
            gr.set_option(_lib.OPT_REDUCE, _lib.REDUCE_NCCL)
            gr.set_eval_counter(0)
            got_nccl = gr.eval_batch_raw(q, r)
            sums_nccl = gr.last_sums(C)
            assert gr.last_reduce_path() == "ncclAllReduce"
            if mode == kfbo.SHARD_CANDIDATES or G == 2:
                assert np.array_equal(sums, sums_nccl) and np.array_equal(got["cost"], got_nccl["cost"]), path
            else:
                np.testing.assert_allclose(sums, sums_nccl, rtol=1e-12)
            # back to the default path for a run of single-candidate evaluations (both generations of the exchange buffers)
            gr.set_option(_lib.OPT_REDUCE, _lib.REDUCE_AUTO)
            gr.set_eval_counter(50)
            run_a = [gr.eval([0.4 + 0.05 * i], [0.1]).cost.copy() for i in range(12)]
            gr.set_option(_lib.OPT_REDUCE, _lib.REDUCE_NCCL)
            gr.set_eval_counter(50)
            run_b = [gr.eval([0.4 + 0.05 * i], [0.1]).cost.copy() for i in range(12)]
            np.testing.assert_allclose(np.stack(run_a), np.stack(run_b), rtol=1e-10, atol=1e-13)
            # one host thread per device (the default, above) or the calling thread serving one device after the other: the
            # same launches in another order on the host -- the same bits
            gr.set_option(_lib.OPT_REDUCE, _lib.REDUCE_AUTO)
            gr.set_option(_lib.OPT_GROUP_THREADS, 0)
            gr.set_eval_counter(0)
            got_1t = gr.eval_batch_raw(q, r)
            assert np.array_equal(gr.last_sums(C), sums) and np.array_equal(got_1t["cost"], got["cost"])
            gr.set_eval_counter(50)
            run_c = [gr.eval([0.4 + 0.05 * i], [0.1]).cost.copy() for i in range(12)]
            assert all(np.array_equal(a, c) for a, c in zip(run_a, run_c))
            gr.set_option(_lib.OPT_GROUP_THREADS, 1)
        if mode == kfbo.SHARD_CANDIDATES:                               # every slot written by one device, the others add 0.0
            assert np.array_equal(sums, want_sums) and np.array_equal(got["cost"], want["cost"])
        else:                                                           # the same trials, summed in another order
            np.testing.assert_allclose(sums, want_sums, rtol=1e-12)
            np.testing.assert_allclose(got["cost"], want["cost"], rtol=1e-10, atol=1e-13)
        assert np.array_equal(got["index_max"], want["index_max"])
        np.testing.assert_allclose(one, want_one, rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("threads", [1, 0])
def test_a_device_dropping_out_fails_the_evaluation_and_poisons_the_group(kfbo, threads):
    # one device of a group fails before it launches anything (KFBO_OPT_INJECT_FAULT): the evaluation must come back with an
    # error -- not hang: the other device's exchange gives up after KFBO_OPT_PEER_TIMEOUT_MS --, every device must be drained,
    # and the group must refuse further evaluations (its devices' counters and exchange generations no longer agree) instead
    # of returning numbers from mismatched streams.  With one host thread per device and with one thread for all.
    import time
    import torch
    from kf_bayesopt_b200 import _lib
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    rv = kfbo.ReadParaVehicle(nsimruns_=4096, cpu_core_number_=1)
    with kfbo.CostEvaluatorGroup(rv, [(0.1, 201), (0.5, 211)], n_gpus=2, seed=3) as gr:
        gr.set_option(_lib.OPT_GROUP_THREADS, threads)
        gr.set_option(_lib.OPT_PEER_TIMEOUT_MS, 300)
        good = gr.eval([1.0], [0.1]).cost.copy()
        assert good.shape == (2,) and all(good > 0)
        gr.set_device_option(1, _lib.OPT_INJECT_FAULT, 1)
        t = time.perf_counter()
        with pytest.raises(kfbo.KfboError) as e1:
            gr.eval([1.0], [0.1])
        assert time.perf_counter() - t < 5.0                              # the timeout, not the 10 s default, and no hang
        assert e1.value.code in (1, 3)                                    # KFBO_ECUDA of the failed device or KFBO_EPEER of the waiting one
        with pytest.raises(kfbo.KfboError) as e2:
            gr.eval([1.0], [0.1])
        assert e2.value.code == -4                                        # KFBO_ESTATE: fails fast from now on
    # a new group on the same devices works again
    with kfbo.CostEvaluatorGroup(rv, [(0.1, 201), (0.5, 211)], n_gpus=2, seed=3) as gr:
        assert np.array_equal(gr.eval([1.0], [0.1]).cost, good)


def test_cpp_node_drives_several_gpus_from_one_process(kfbo, tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    exe = os.path.join(ROOT, "kf_bayesopt_b200", "bin", "kfbo_test_trial")
    if not os.path.exists(exe):
        pytest.skip("host programs not built")
    y = tmp_path / "vehicleParams.yaml"
    y.write_text("CPUcoreNumber: 10\nstateDOF: 2\nobservationDOF: 1\nt_start: 0.0\nt_end: 20.1\ndt: 0.1\nxi0: 0.0\nxidot0: 0.0\n"
                 "Nsimruns: 200000\npnoise: [1.0]\nonoise: [0.1]\noptimizationChoice: all\ncostChoice: JNEES\n")
    outs = []
    for gpus in (1, 2):
        out = subprocess.run([exe, "--vehicle", str(y), "--pnoise_est", "0.1", "--fused", "--gpus", str(gpus), "--seed", "9"],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        outs.append(float(next(l for l in out.stdout.splitlines() if l.startswith("cost JNEES")).split()[-1]))
    assert outs[0] == pytest.approx(1.613, abs=0.01)
    assert outs[1] == pytest.approx(outs[0], rel=1e-5)                  # same trials and streams; printed with 6 digits


def test_the_callers_current_cuda_device_is_preserved(kfbo):
    """The library sits next to other CUDA users of the calling thread (torch): a handle / a group on another device must
    leave the thread's current device where it was, on every entry point."""
    import numpy as np
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    torch.cuda.set_device(1)
    x = torch.ones(8, device="cuda")                                   # torch's work lives on device 1
    rv = kfbo.ReadParaVehicle(nsimruns_=512, cpu_core_number_=1)
    with kfbo.CostEvaluator(rv, [(0.1, 201), (0.5, 211)], device=0, max_candidates=3) as ev:
        assert torch.cuda.current_device() == 1
        ev.eval([1.0], [0.1])
        assert torch.cuda.current_device() == 1
        ev.eval_batch_raw([0.5, 1.0, 2.0], [0.1, 0.1, 0.2])
        ev.eval_batch_costs([0.5, 1.0, 2.0], [0.1, 0.1, 0.2])
        ev.eval_philox_trials(0, 1.0, 0.1, 0, 0, 64)
        ev.trace_trial(0, 1.0, 0.1, 0, 0)
        ev.last_kernel_ms()
        assert torch.cuda.current_device() == 1
    assert torch.cuda.current_device() == 1
    with kfbo.CostEvaluatorGroup(rv, [(0.1, 201)], n_gpus=2) as gr:
        gr.eval([1.0], [0.1])
        assert torch.cuda.current_device() == 1
    kfbo.measure_fp64_peak(0, 5.0)
    assert torch.cuda.current_device() == 1
    assert float((x + 1).sum().item()) == 16.0 and x.device.index == 1
    torch.cuda.set_device(0)
