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


# ---- the reference's parameter readers and its NODE (trial_node.cpp) on the ROS stand-in of oracle/ref_shim/ros/ros.h

def _node_lib():
    l = lib()
    if not getattr(l, "_node_ready", False):
        _ip = C.POINTER(C.c_int)
        l.kfbo_ref_param_clear.restype = None
        l.kfbo_ref_param_set_int.argtypes = [C.c_char_p, C.c_longlong]
        l.kfbo_ref_param_set_double.argtypes = [C.c_char_p, C.c_double]
        l.kfbo_ref_param_set_string.argtypes = [C.c_char_p, C.c_char_p]
        l.kfbo_ref_param_set_vector.argtypes = [C.c_char_p, _dp, C.c_int]
        l.kfbo_ref_read_para_vehicle.argtypes = [C.c_int, _dp, C.c_char_p, C.c_int]
        l.kfbo_ref_read_para_bayes.argtypes = [_dp]
        l.kfbo_ref_node_run.restype = C.c_int
        l.kfbo_ref_node_run.argtypes = [C.c_int, _ip, _ip, _dp, C.c_longlong, C.c_int, _dp, C.c_int, C.POINTER(C.c_longlong)]
        l.kfbo_ref_eval_threads.restype = C.c_int
        l.kfbo_ref_eval_threads.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                            C.c_longlong, _dp, _ip]
        l._node_ready = True
    return l


def set_params(params: dict, clear: bool = True) -> None:
    """Fill the stand-in parameter server (what `rosparam load file.yaml` does): ints, floats, strings and lists of
    numbers, typed as yaml types them -- which is what roscpp's getParam overloads see."""
    l = _node_lib()
    if clear:
        l.kfbo_ref_param_clear()
    for k, v in params.items():
        name = k.encode()
        if isinstance(v, bool):
            l.kfbo_ref_param_set_int(name, int(v))
        elif isinstance(v, int):
            l.kfbo_ref_param_set_int(name, v)
        elif isinstance(v, float):
            l.kfbo_ref_param_set_double(name, v)
        elif isinstance(v, str):
            l.kfbo_ref_param_set_string(name, v.encode())
        elif isinstance(v, (list, tuple)):
            a = np.ascontiguousarray(v, dtype=np.float64)
            l.kfbo_ref_param_set_vector(name, _p(a), a.size)
        else:
            raise TypeError(f"parameter {k!r}: unsupported type {type(v).__name__}")


def load_yaml_params(*paths: str) -> dict:
    """yaml files -> one parameter dict (later files override, like consecutive <rosparam load>)."""
    import yaml
    out = {}
    for p in paths:
        with open(p) as f:
            out.update(yaml.safe_load(f) or {})
    return out


RPV_FIELDS = ("cpu_core_number_", "state_dof_", "observation_dof_", "t_start_", "dt_", "t_end_", "xi0_", "xidot0_", "nsimruns_",
              "max_iterations_", "dim_pn_", "dim_on_", "pnoise0", "onoise0")


def read_para_vehicle(show_read_info: bool = False) -> dict:
    """ReadParaVehicle(nh, showReadInfo) of the reference on the current parameter server -> its members."""
    nums, buf = np.empty(14), C.create_string_buffer(512)
    rc = _node_lib().kfbo_ref_read_para_vehicle(int(show_read_info), _p(nums), buf, 512)
    assert rc == 0
    out = dict(zip(RPV_FIELDS, nums.tolist()))
    out["P0_"], out["optimization_choice_"], out["cost_choice_"] = buf.value.decode().split("\n")
    return out


def read_para_bayes() -> dict:
    out = np.empty(7)
    _node_lib().kfbo_ref_read_para_bayes(_p(out))
    return {"n_init_samples_": int(out[0]), "n_iterations_": int(out[1]), "lower_bound_": out[2:4].tolist(),
            "upper_bound_": out[4:6].tolist(), "noise_": float(out[6])}


def node_run(messages, clock0: int = 1000, quiet: bool = True):
    """Queue `messages` = [(pn list, on list), ...] on the "noise" topic, run trial_node.cpp's main() (RunTrial ctor,
    subscribe, spin -> RunTrial::ProcessNoiseCallback per message) and return (the costs it published, clock reads).
    Vehicle / bayes parameters: the current parameter server (set_params)."""
    npn = np.array([len(pn) for pn, _ in messages], dtype=np.int32)
    non = np.array([len(on) for _, on in messages], dtype=np.int32)
    data = np.ascontiguousarray([x for pn, on in messages for x in list(pn) + list(on)], dtype=np.float64)
    costs, calls = np.full(len(messages) + 4, np.nan), C.c_longlong()
    _ip = C.POINTER(C.c_int)
    n = _node_lib().kfbo_ref_node_run(len(messages), npn.ctypes.data_as(_ip), non.ctypes.data_as(_ip), _p(data), clock0, int(quiet),
                                      _p(costs), costs.size, C.byref(calls))
    return costs[:n].copy(), calls.value


def eval_threads(dt, t_end, nsimruns, cores, q_truth, r_truth, q_est, r_est, clock0=1000):
    """One sample time of the callback (trial_node.cpp:54-75): `cores` Trial objects, one std::thread per Trial::Run,
    the four getters averaged over the threads -> (avg[4], max_iterations_).  cores > 1 races on the reference's
    process-wide static mt19937 exactly as the reference does: results are then not repeatable."""
    out, T = np.empty(4), C.c_int()
    _node_lib().kfbo_ref_eval_threads(dt, t_end, nsimruns, cores, q_truth, r_truth, q_est, r_est, clock0, _p(out), C.byref(T))
    return out, T.value
