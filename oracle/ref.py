"""ctypes wrapper of oracle/_ref/libkfbo_ref.so: the REFERENCE'S OWN source files (trial.cpp, simulator.cpp,
kalman_filter.cpp, read_para_vehicle.cpp and include/*.hpp of arpg/kf_bayesopt) compiled from where they lie,
with oracle/ref_shim/ standing in for Eigen3 / ROS / matplotlibcpp and a counter for the wall clock.

TEST INFRASTRUCTURE ONLY (see oracle/ref_driver.cpp).  Built by oracle/Makefile when /root/reference is present;
`available()` is False elsewhere (e.g. on the GPU box) and the tests that need it skip."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libkfbo_ref.so")
_dp = C.POINTER(C.c_double)
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        l = C.CDLL(LIB_PATH)
        l.kfbo_ref_trial_run.restype = C.c_int
        l.kfbo_ref_trial_run.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_longlong, _dp, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
        l.kfbo_ref_models.restype = C.c_int
        l.kfbo_ref_models.argtypes = [C.c_double, C.c_double, C.c_double, _dp, _dp, _dp, C.c_int, _dp, _dp]
        l.kfbo_ref_continuous_to_discrete.restype = C.c_int
        l.kfbo_ref_continuous_to_discrete.argtypes = [_dp, _dp, C.c_double, _dp, _dp]
        l.kfbo_ref_kf_run.restype = C.c_int
        l.kfbo_ref_kf_run.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, _dp, _dp]
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def trial_run(dt, t_end, nsimruns, cores, q_truth, r_truth, q_est, r_est, clock0=1000):
    """Trial(rv_gt, rv).Run() -> (getters[4], clock reads, max_iterations_)."""
    out, calls, T = np.empty(4), C.c_longlong(), C.c_int()
    lib().kfbo_ref_trial_run(dt, t_end, nsimruns, cores, q_truth, r_truth, q_est, r_est, clock0, _p(out), C.byref(calls), C.byref(T))
    return out, calls.value, T.value


def models(pnoise, onoise, dt, n_u=2000):
    Fd, Qd, gu, H, R = np.empty(4), np.empty(4), np.empty((n_u, 2)), np.empty(2), np.empty(1)
    lib().kfbo_ref_models(pnoise, onoise, dt, _p(Fd), _p(Qd), _p(gu), n_u, _p(H), _p(R))
    return Fd.reshape(2, 2), Qd.reshape(2, 2), gu, H, float(R[0])


def continuous_to_discrete(Fc, Qc, dt):
    Fc, Qc = np.ascontiguousarray(Fc, dtype=np.float64), np.ascontiguousarray(Qc, dtype=np.float64)
    Fd, Qd = np.empty(4), np.empty(4)
    lib().kfbo_ref_continuous_to_discrete(_p(Fc), _p(Qc), dt, _p(Fd), _p(Qd))
    return Fd.reshape(2, 2), Qd.reshape(2, 2)


def kf_run(pnoise_est, onoise_est, dt, z):
    z = np.ascontiguousarray(z, dtype=np.float64)
    out = np.empty((z.size, 8))
    lib().kfbo_ref_kf_run(pnoise_est, onoise_est, dt, z.size, _p(z), _p(out))
    return out
