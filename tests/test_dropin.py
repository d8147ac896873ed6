"""THE REFERENCE'S UNMODIFIED NODE on the B200 evaluator.  tests/native/dropin builds /root/reference/trial_node.cpp and
/root/reference/src/test_trial.cpp where they lie, with include/compat ahead of the reference's include/ -- so their
`#include "trial.h"` / `"read_para_vehicle.h"` are this repository's drop-ins over the C ABI -- and the ROS stand-in of
oracle/ref_shim.  Nothing of the callers changes: RunTrial::ProcessNoiseCallback still builds CPUcoreNumber Trial
objects per sample time, runs Trial::Run on std::threads, averages the getters, takes the max (trial_node.cpp:49-119);
each Trial::Run is now a kernel launch.  This is SURVEY.md 8(b)'s "header-only Trial-compatible adapter", exercised by
the reference's own code.

The library is prebuilt here (build() runs the Makefile when the reference tree is present) and travels to the GPU box."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "tests", "native", "dropin", "_dropin", "libkfbo_dropin.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="tests/native/dropin not built (needs /root/reference)")
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)

# config/vehicleParams.yaml + the one key of config/bayesParams.yaml the node reads (trial_node.cpp:122)
VEHICLE = {"CPUcoreNumber": 10, "stateDOF": 2, "observationDOF": 1, "t_start": 0.0, "t_end": 20.1, "dt": 0.1, "xi0": 0.0,
           "xidot0": 0.0, "Nsimruns": 200, "pnoise": [1.0], "onoise": [0.1], "optimizationChoice": "all", "costChoice": "JNEES",
           "n_init_samples": 20}


@pytest.fixture(scope="module")
def dropin(kfbo):
    l = C.CDLL(LIB)
    l.kfbo_dropin_param_clear.restype = None
    l.kfbo_dropin_param_set_int.argtypes = [C.c_char_p, C.c_longlong]
    l.kfbo_dropin_param_set_double.argtypes = [C.c_char_p, C.c_double]
    l.kfbo_dropin_param_set_string.argtypes = [C.c_char_p, C.c_char_p]
    l.kfbo_dropin_param_set_vector.argtypes = [C.c_char_p, _dp, C.c_int]
    l.kfbo_dropin_node_run.restype = C.c_int
    l.kfbo_dropin_node_run.argtypes = [C.c_int, _ip, _ip, _dp, _dp, C.c_int, C.c_char_p, C.c_int]
    l.kfbo_dropin_test_trial_run.restype = C.c_int
    l.kfbo_dropin_test_trial_run.argtypes = [C.c_char_p, C.c_int]
    return l


def set_params(l, **over):
    l.kfbo_dropin_param_clear()
    for k, v in dict(VEHICLE, **over).items():
        if isinstance(v, int):
            l.kfbo_dropin_param_set_int(k.encode(), v)
        elif isinstance(v, float):
            l.kfbo_dropin_param_set_double(k.encode(), v)
        elif isinstance(v, str):
            l.kfbo_dropin_param_set_string(k.encode(), v.encode())
        else:
            a = np.ascontiguousarray(v, dtype=np.float64)
            l.kfbo_dropin_param_set_vector(k.encode(), a.ctypes.data_as(_dp), a.size)


def node_run(l, messages):
    npn = np.array([len(pn) for pn, _ in messages], dtype=np.int32)
    non = np.array([len(on) for _, on in messages], dtype=np.int32)
    data = np.ascontiguousarray([x for pn, on in messages for x in list(pn) + list(on)], dtype=np.float64)
    costs, log = np.full(len(messages) + 4, np.nan), C.create_string_buffer(1 << 16)
    n = l.kfbo_dropin_node_run(len(messages), npn.ctypes.data_as(_ip), non.ctypes.data_as(_ip), data.ctypes.data_as(_dp),
                               costs.ctypes.data_as(_dp), costs.size, log, len(log))
    return costs[:n].copy(), log.value.decode()


def test_without_a_device_the_node_publishes_nan_not_a_cpu_answer(dropin):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    set_params(dropin, CPUcoreNumber=2, Nsimruns=4)
    costs, log = node_run(dropin, [([1.0], [0.1])])
    # one "cost" per "noise" as always; no device -> every Trial::Run records its error, the getters stay NaN and so
    # does the cost (the optimiser's wait loop, robot1d_bayesopt.cpp:107, ends on NaN): there is no CPU fallback
    assert len(costs) == 1 and np.isnan(costs[0])
    assert "get the noise" in log and "publish the cost" in log


@pytest.mark.gpu
def test_reference_node_unmodified_runs_on_the_gpu(dropin, kfbo):
    # truth q = 1, r = 0.1; 20000 trials over the yaml's 10 threads; the callback's own two passes (dt 0.1 / 1.0)
    set_params(dropin, Nsimruns=20000)
    msgs = [([1.0], [0.1]), ([0.1], [0.1]), ([5.0], [0.1]), ([1.0], [1.0])]
    costs, log = node_run(dropin, msgs)
    assert len(costs) == len(msgs) and np.all(np.isfinite(costs))
    # SURVEY 8c anchors (and tests/golden stat_*): matched -> ~0; q_est = 0.1 -> J = |log(NEES/2)| ~ 1.61 (dt 0.1) / 1.59 (dt 1.0);
    # q_est = 5 -> 0.474 / 0.455; r_est = 1 -> |log(1.124/2)| = 0.576 / |log(0.9865/2)| = 0.707
    assert costs[0] < 0.01
    assert costs[1] == pytest.approx(1.613, abs=0.01)
    assert costs[2] == pytest.approx(0.474, abs=0.01)
    assert costs[3] == pytest.approx(0.707, abs=0.01)
    # the node narrates both passes and the max it publishes (trial_node.cpp:84,117)
    pairs = re.findall(r"^([0-9.e+-]+),([0-9.e+-]+)$", log, flags=re.M)
    assert len(pairs) == len(msgs)
    for (a, b), c in zip(pairs, costs):
        assert max(float(a), float(b)) == pytest.approx(c, rel=1e-4)
    # the same four messages through this repository's own RunTrial (one batched launch per message): same costs
    # within the Monte Carlo error of two independent 20000-trial estimates
    rv = kfbo.ReadParaVehicle(nsimruns_=20000, cpu_core_number_=10)
    rt = kfbo.RunTrial(rv)
    for (pn, on), c in zip(msgs, costs):
        own = rt.ProcessNoiseCallback(kfbo.Float64MultiArray.noise(pn, on)).data
        assert own == pytest.approx(c, abs=0.012)


@pytest.mark.gpu
def test_reference_node_answers_an_impossible_candidate_with_nan(dropin):
    # pnoise <= 0 / NaN: the reference's own arithmetic publishes NaN for such a message (the optimiser's wait loop ends on
    # it, robot1d_bayesopt.cpp:107); through the drop-in the unmodified node does the same, and goes on to the next message
    set_params(dropin, Nsimruns=2000)
    costs, _ = node_run(dropin, [([-1.0], [0.1]), ([1.0], [0.0]), ([1.0], [0.1])])
    assert len(costs) == 3 and np.isnan(costs[0]) and np.isnan(costs[1]) and 0 <= costs[2] < 0.05


@pytest.mark.gpu
def test_reference_node_cost_choices_and_thread_counts(dropin):
    for choice, lo, hi in [("JNIS", 0.0, 0.01), ("CNEES", 0.0, 0.1), ("CNIS", 0.0, 0.05)]:
        for cores in (1, 4, 10):
            set_params(dropin, Nsimruns=20000, CPUcoreNumber=cores, costChoice=choice)
            costs, _ = node_run(dropin, [([1.0], [0.1])])
            assert lo <= costs[0] < hi, (choice, cores, costs)


@pytest.mark.gpu
def test_reference_test_trial_unmodified_runs_on_the_gpu(dropin):
    # src/test_trial.cpp: one sample time, pnoise_est / onoise_est from the parameter server (launch file: 1.0 / 0.1)
    set_params(dropin, Nsimruns=20000, pnoise_est=0.1, onoise_est=0.1)
    buf = C.create_string_buffer(1 << 14)
    dropin.kfbo_dropin_test_trial_run(buf, len(buf))
    log = buf.value.decode()
    assert "estimator process noise 0.1" in log and "simulator process noise 1" in log
    j = float(re.search(r"cost JNEES ([0-9.e+-]+)", log).group(1))
    assert j == pytest.approx(1.613, abs=0.01)
