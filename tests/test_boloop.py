"""BASELINE.json configs[4] -- "full Bayesian-optimisation loop (bayesopt node unchanged ...) driving the GPU cost evaluator over
ROS topics" -- with BOTH of the reference's nodes UNMODIFIED in one process (tests/native/boloop):
  * /root/reference/robot1d_bayesopt.cpp, the optimiser node: Robot1D::evaluateSample builds the "noise" message per
    optimizationChoice (robot1d_bayesopt.cpp:53-88), sleeps, publishes, waits for "cost" (:91-113);
  * /root/reference/trial_node.cpp, the evaluator node, on include/compat: its Trial::Run calls are launches on the B200;
one thread each, over the roscpp stand-in in bus mode.  What stands in for the absent bayesopt LIBRARY is a header with its
interface over this repository's small GP / EI optimiser (tests/native/boloop/shim/bayesopt/bayesopt.hpp): the optimiser's
algorithm is therefore not the reference's (parity unpinned) -- everything around it is the reference's own code.
The library is prebuilt where the reference tree is present (build()) and travels to the GPU box."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import yaml

from conftest import ROOT
from test_host_cpp import BAYES_YAML, VEHICLE_YAML

LIB = os.path.join(ROOT, "tests", "native", "boloop", "_boloop", "libkfbo_boloop.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="tests/native/boloop not built (needs /root/reference)")
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def boloop(kfbo):
    l = C.CDLL(LIB)
    l.kfbo_boloop_param_clear.restype = None
    l.kfbo_boloop_param_set_int.argtypes = [C.c_char_p, C.c_longlong]
    l.kfbo_boloop_param_set_double.argtypes = [C.c_char_p, C.c_double]
    l.kfbo_boloop_param_set_string.argtypes = [C.c_char_p, C.c_char_p]
    l.kfbo_boloop_param_set_vector.argtypes = [C.c_char_p, _dp, C.c_int]
    l.kfbo_boloop_run.restype = C.c_int
    l.kfbo_boloop_run.argtypes = [C.c_int, C.c_char_p, C.c_double, _dp, _dp, C.c_int, C.POINTER(C.c_int), _dp, _dp, C.c_char_p, C.c_int]
    return l


def set_params(l, **over):
    """The parameter server as `rosparam load` of the reference's two yaml files leaves it (launch/robot1d_kf.launch:3-4)."""
    params = dict(yaml.safe_load(VEHICLE_YAML), **yaml.safe_load(BAYES_YAML))
    params.update(over)
    l.kfbo_boloop_param_clear()
    for k, v in params.items():
        if isinstance(v, bool) or isinstance(v, int):
            l.kfbo_boloop_param_set_int(k.encode(), int(v))
        elif isinstance(v, float):
            l.kfbo_boloop_param_set_double(k.encode(), v)
        elif isinstance(v, str):
            l.kfbo_boloop_param_set_string(k.encode(), v.encode())
        else:
            a = np.ascontiguousarray(v, dtype=np.float64)
            l.kfbo_boloop_param_set_vector(k.encode(), a.ctypes.data_as(_dp), a.size)
    return params


def run(l, sleep_scale=0.0, cap=512, evaluator=0, evaluator_args=""):
    """evaluator 0: the reference's trial_node.cpp (on include/compat); 1: this repository's robot1d_kf.cpp (-DKFBO_WITH_ROS)."""
    noise, costs = np.zeros((cap, 4)), np.full(cap, np.nan)
    nc, wall, slept, log = C.c_int(0), C.c_double(0), C.c_double(0), C.create_string_buffer(1 << 15)
    n = l.kfbo_boloop_run(evaluator, evaluator_args.encode(), sleep_scale, noise.ctypes.data_as(_dp), costs.ctypes.data_as(_dp), cap, C.byref(nc), C.byref(wall),
                          C.byref(slept), log, len(log))
    return noise[:n].copy(), costs[:nc.value].copy(), wall.value, slept.value, log.value.decode(errors="replace")


def test_both_reference_nodes_compiled_unmodified(boloop):
    assert hasattr(boloop, "kfbo_boloop_run")          # robot1d_bayesopt.cpp + trial_node.cpp built where they lie and linked


@pytest.mark.gpu
def test_reference_optimiser_node_drives_the_reference_evaluator_node_on_the_gpu(boloop):
    p = set_params(boloop, n_iterations=30, random_seed=7, verbose_level=0)
    noise, costs, wall, slept, log = run(boloop)
    n = p["n_init_samples"] + p["n_iterations"]
    assert len(noise) == n and len(costs) == n                              # one "cost" per "noise", 20 + 30 of them
    assert np.all(noise[:, 0] == 1) and np.all(noise[:, 1] == 1)            # layout.dim sizes = |pnoise|, |onoise| (:58-61)
    lo, hi = np.array(p["lower_bound"]), np.array(p["upper_bound"])
    assert np.all(noise[:, 2:] >= lo - 1e-12) and np.all(noise[:, 2:] <= hi + 1e-12)   # optimizationChoice "all": data = the query
    assert np.all(np.isfinite(costs)) and np.all(costs >= 0)                # |log| costs from the GPU
    m = re.search(r"final optimization result ([0-9.e+-]+),([0-9.e+-]+),", log)          # robot1d_bayesopt.cpp:266-269
    assert m, log[-2000:]
    best = np.array([float(m.group(1)), float(m.group(2))])
    assert np.all(best >= lo) and np.all(best <= hi)
    assert costs.min() < 0.1                                                 # found a consistent filter (truth: q = 1, r = 0.1)
    assert slept == 0.0 and "get the cost, now change it to" in log and "publish the noise" in log
    # the evaluator node counted which sample time won after the initial samples (trial_node.cpp:121-126)
    assert f"it count {p['n_iterations']}," in log and "freq of index0 and index1" in log


@pytest.mark.gpu
@pytest.mark.parametrize("choice,lb,ub,col,fixed", [("processNoise", [0.1], [5.0], 3, 0.1), ("observationNoise", [0.01], [1.0], 2, 1.0)])
def test_optimization_choice_is_the_optimiser_nodes_business(boloop, choice, lb, ub, col, fixed):
    # processNoise: the query is the process noise, the observation noise is the yaml's (robot1d_bayesopt.cpp:68-72);
    # observationNoise: the other way round (:73-77).  The evaluator always receives both.
    set_params(boloop, optimizationChoice=choice, lower_bound=lb, upper_bound=ub, n_init_samples=6, n_iterations=6, random_seed=3,
               verbose_level=0)
    noise, costs, _, _, log = run(boloop)
    assert len(noise) == 12 == len(costs) and np.all(np.isfinite(costs))
    assert np.all(noise[:, col] == fixed)
    other = noise[:, 5 - col]
    assert np.all(other >= lb[0] - 1e-12) and np.all(other <= ub[0] + 1e-12) and len(set(other)) > 6


@pytest.mark.gpu
def test_the_reference_sleeps_half_a_second_per_iteration(boloop):
    # robot1d_bayesopt.cpp:91,96: usleep(300000) + usleep(200000) around every publish -- measured, not assumed
    set_params(boloop, n_init_samples=2, n_iterations=2, random_seed=1, verbose_level=0)
    noise, costs, wall, slept, _ = run(boloop, sleep_scale=1.0)
    assert len(costs) == 4 and slept == pytest.approx(4 * 0.5, abs=1e-9) and wall >= 2.0


@pytest.mark.gpu
def test_reference_optimiser_node_drives_this_repositorys_ros_node(boloop, tmp_path, kfbo):
    # the deployment configs[4] describes: the bayesopt node unchanged, the evaluator node replaced by kfbo_robot1d_kf (ROS glue,
    # one fused launch per message).  Every cost it publishes is what the Python RunTrial gives for the same message and seed.
    vy = tmp_path / "vehicleParams.yaml"
    vy.write_text(VEHICLE_YAML)
    seed = 99
    p = set_params(boloop, n_init_samples=8, n_iterations=12, random_seed=5, verbose_level=0, **{"~vehicle_yaml": str(vy)})
    noise, costs, wall, slept, log = run(boloop, evaluator=1, evaluator_args=f"--seed {seed} --quiet")
    assert len(noise) == 20 == len(costs) and np.all(np.isfinite(costs))
    rt = kfbo.RunTrial(kfbo.ReadParaVehicle.from_yaml(str(vy)), seed=seed)
    for (_, _, q, r), c in zip(noise, costs):
        assert rt.ProcessNoiseCallback(kfbo.Float64MultiArray.noise([q], [r])).data == c     # same streams: the same bits
    assert re.search(r"final optimization result ([0-9.e+-]+),([0-9.e+-]+),", log)
