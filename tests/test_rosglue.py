"""This repository's own ROS node -- kf_bayesopt_b200/host/robot1d_kf.cpp compiled with -DKFBO_WITH_ROS -- built and RUN
against the roscpp stand-in of oracle/ref_shim (tests/native/rosglue; ROS is not installed here).  The glue is what a
maintainer of the reference would launch instead of robot1d_kf: ros::init("robot_1d_kf"), the private ~vehicle_yaml
parameter, advertise("cost", 10), subscribe("noise", 10, boost::function), ros::spin() (trial_node.cpp:144-157), plus a
catch that answers a malformed message with a NaN cost instead of dying inside spin().  The stand-in restates roscpp's API
shape, so this proves the glue compiles against that shape and behaves; a build against real roscpp remains untested."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_host_cpp import yamls  # noqa: F401  (fixture)

DIR = os.path.join(ROOT, "tests", "native", "rosglue")
LIB = os.path.join(DIR, "_rosglue", "libkfbo_rosglue.so")
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def glue(kfbo):
    subprocess.check_call(["make", "-C", DIR, "-s"])
    l = C.CDLL(LIB)
    l.kfbo_rosglue_run.restype = C.c_int
    l.kfbo_rosglue_run.argtypes = [C.c_char_p, C.c_char_p, C.c_int, _ip, _ip, _ip, _dp, _dp, C.c_int, _ip]
    return l


def run(l, yaml, argv, messages):
    """messages: (npn, non, data) -- data shorter than npn + non makes a malformed message."""
    npn = np.array([m[0] for m in messages], dtype=np.int32)
    non = np.array([m[1] for m in messages], dtype=np.int32)
    nd = np.array([len(m[2]) for m in messages], dtype=np.int32)
    data = np.ascontiguousarray([x for m in messages for x in m[2]], dtype=np.float64)
    costs, rc = np.full(len(messages) + 4, -1.0), C.c_int(-999)
    n = l.kfbo_rosglue_run(yaml.encode(), argv.encode(), len(messages), npn.ctypes.data_as(_ip), non.ctypes.data_as(_ip),
                           nd.ctypes.data_as(_ip), data.ctypes.data_as(_dp), costs.ctypes.data_as(_dp), costs.size, C.byref(rc))
    return costs[:n].copy(), rc.value


def test_without_a_device_the_ros_node_exits_with_an_error_and_publishes_nothing(glue, yamls):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    costs, rc = run(glue, yamls[0], "--quiet", [(1, 1, [1.0, 0.1])])
    assert rc != 0 and len(costs) == 0           # no CPU fallback: the node refuses to start


@pytest.mark.gpu
def test_ros_node_answers_noise_with_cost_like_the_python_run_trial(glue, yamls, kfbo):
    seed = 20261020
    msgs = [(1, 1, [0.7, 0.2]), (1, 1, [1.0, 0.1]), (3, 1, [1.0, 0.1]), (1, 1, [3.0, 0.05])]     # the third is malformed
    costs, rc = run(glue, yamls[0], f"--seed {seed} --quiet", msgs)
    assert rc == 0 and len(costs) == len(msgs)          # one "cost" per "noise", the node survived the malformed one
    assert np.isnan(costs[2])                            # ... which it answered with NaN (robot1d_bayesopt.cpp:107 moves on)
    rv = kfbo.ReadParaVehicle.from_yaml(yamls[0])
    rt = kfbo.RunTrial(rv, seed=seed)
    for i in (0, 1, 3):
        want = rt.ProcessNoiseCallback(kfbo.Float64MultiArray.noise(msgs[i][2][:1], msgs[i][2][1:])).data
        assert costs[i] == want                          # same library, same Philox streams: the same bits
    # several GPUs from the one node process, where the box has them
    import torch
    if torch.cuda.device_count() >= 2:
        costs2, rc2 = run(glue, yamls[0], f"--seed {seed} --quiet --gpus 2", [msgs[0], msgs[1]])
        assert rc2 == 0 and np.allclose(costs2, costs[:2], rtol=1e-10, atol=1e-13)
