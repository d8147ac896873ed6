"""ctypes loader for the CPU oracle (oracle/kfbo_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of kfbo_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
the kf_bayesopt_b200 package.  PARITY UNPINNED (the reference has no golden vectors).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkfbo_oracle.so")

COST_CHOICES = {"JNEES": 0, "CNEES": 1, "JNIS": 2, "CNIS": 3}


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "kfbo_oracle.c"), os.path.join(_HERE, "kfbo_oracle_aug.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libkfbo_oracle.so"])
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ko_discretize.argtypes = [C.c_double, C.c_double, _dp, _dp]
        L.ko_control_table.argtypes = [C.c_double, C.c_int, _dp]
        L.ko_eval_supplied.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_long, _dp, _dp, _dp, _dp]
        L.ko_eval_supplied.restype = C.c_int
        L.ko_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.ko_normal_pair.argtypes = [C.c_uint32, C.c_uint32, _dp, _dp]
        L.ko_noise_transform.argtypes = [_dp, C.c_double, _dp, _dp]
        L.ko_eval_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_double, C.c_double,
                                     C.c_double, C.c_double, C.c_uint32, C.c_long, _dp]
        L.ko_eval_philox.restype = C.c_int
        L.ko_thread_averages.argtypes = [_dp, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, _dp]
        L.ko_cost_from_threads.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.ko_max_cost.argtypes = [_dp, C.c_int, _dp, C.POINTER(C.c_int)]
        L.ko_eval_mt.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                 C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, _dp]
        L.ko_eval_mt.restype = C.c_int
        L.ko_riccati_trace.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, _dp]
        L.ko_eig_transform2.argtypes = [_dp, _dp]
        L.ko_mt19937_nth.argtypes = [C.c_uint32, C.c_int]
        L.ko_mt19937_nth.restype = C.c_uint32
        L.ko_trial_run_refclock.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                            C.c_int, C.c_longlong, _dp]
        L.ko_trial_run_refclock.restype = C.c_int
        L.ko_refclock_draws.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_longlong, C.c_int, _dp, _dp, _dp]
        L.ko_refclock_draws.restype = C.c_int
        L.ko_kf_run.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, _dp, _dp]
        L.ko_kf_run.restype = C.c_int
        L.ko_gu_table.argtypes = [C.c_double, C.c_int, _dp]
        # the state-augmented variants (kfbo_oracle_aug.c)
        L.koa_expm.argtypes = [C.c_int, _dp, _dp]
        L.koa_discretize.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp]
        L.koa_discretize.restype = C.c_int
        L.koa_gvector.argtypes = [C.c_int, C.c_double, _dp]
        L.koa_cholesky.argtypes = [C.c_int, _dp, _dp]
        L.koa_cholesky.restype = C.c_int
        L.koa_eval_supplied.argtypes = [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_long, _dp, _dp, _dp, _dp]
        L.koa_eval_supplied.restype = C.c_int
        L.koa_eval_philox.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                                      C.c_double, C.c_uint32, C.c_long, _dp]
        L.koa_eval_philox.restype = C.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def discretize(q: float, dt: float):
    Fd, Qd = np.empty(4), np.empty(4)
    lib().ko_discretize(q, dt, _p(Fd), _p(Qd))
    return Fd.reshape(2, 2), Qd.reshape(2, 2)


def control_table(dt: float, n: int = 2000) -> np.ndarray:
    u = np.empty(n)
    lib().ko_control_table(dt, n, _p(u))
    return u


def eval_supplied(dt, T, q_est, r_est, x0noise, w, v) -> np.ndarray:
    """per_trial[4][B] = {row_nees, row_nis, var_nees, var_nis}; x0noise[2][B], w[T][2][B], v[T][B]."""
    x0noise, w, v = _f64(x0noise), _f64(w), _f64(v)
    B = x0noise.shape[1]
    assert x0noise.shape == (2, B) and w.shape == (T, 2, B) and v.shape == (T, B)
    out = np.empty((4, B))
    rc = lib().ko_eval_supplied(dt, T, q_est, r_est, B, _p(x0noise), _p(w), _p(v), _p(out))
    if rc:
        raise ValueError(f"ko_eval_supplied rc={rc}")
    return out


def philox4x32_10(ctr, key) -> np.ndarray:
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (C.c_uint32 * 4)()
    lib().ko_philox4x32_10(c, k, o)
    return np.array(list(o), dtype=np.uint32)


def normal_pair(w0, w1):
    """64 random bits -> (rad cos theta, rad sin theta): 44-bit radius variate, 20-bit direction; see ko_normal_pair."""
    a, b = C.c_double(), C.c_double()
    lib().ko_normal_pair(int(w0) & 0xFFFFFFFF, int(w1) & 0xFFFFFFFF, C.byref(a), C.byref(b))
    return a.value, b.value


def noise_transform(Qd, R):
    L, sr = np.empty(3), C.c_double()
    lib().ko_noise_transform(_p(_f64(Qd).reshape(-1)), R, _p(L), C.byref(sr))
    return L, sr.value


def eval_philox(seed, stream, dt, T, q_truth, r_truth, q_est, r_est, trial0, ntrials) -> np.ndarray:
    out = np.empty((4, ntrials))
    rc = lib().ko_eval_philox(seed, stream, dt, T, q_truth, r_truth, q_est, r_est, trial0, ntrials, _p(out))
    if rc:
        raise ValueError(f"ko_eval_philox rc={rc}")
    return out


def thread_averages(per_trial, off, T, nsimruns, cores) -> np.ndarray:
    per_trial = _f64(per_trial)
    avg = np.empty(4)
    lib().ko_thread_averages(_p(per_trial), per_trial.shape[1], off, T, nsimruns, cores, _p(avg))
    return avg


def cost_from_threads(thread_avg, state_dof=2, obs_dof=1, cost_choice="JNEES") -> np.ndarray:
    thread_avg = _f64(thread_avg)
    out = np.empty(5)
    lib().ko_cost_from_threads(_p(thread_avg), thread_avg.shape[0], state_dof, obs_dof,
                               COST_CHOICES[cost_choice], _p(out))
    return out


def cost_from_per_trial(per_trial, T, nsimruns, cores, state_dof=2, obs_dof=1, cost_choice="JNEES"):
    """trial_node.cpp:54-102 for one sample time given per-trial outputs of the (nsimruns/cores)*cores trials
    actually run, thread th owning columns [th*n, (th+1)*n)."""
    n = nsimruns // cores
    ta = np.stack([thread_averages(per_trial, th * n, T, nsimruns, cores) for th in range(cores)])
    return cost_from_threads(ta, state_dof, obs_dof, cost_choice)


def max_cost(pic_max):
    pic_max = _f64(pic_max)
    m, i = C.c_double(), C.c_int()
    lib().ko_max_cost(_p(pic_max), len(pic_max), C.byref(m), C.byref(i))
    return m.value, i.value


def eval_mt(dt, T, q_truth, r_truth, q_est, r_est, nsimruns, cores, seed=1, state_dof=2, obs_dof=1,
            cost_choice="JNEES") -> np.ndarray:
    out = np.empty(5)
    rc = lib().ko_eval_mt(dt, T, q_truth, r_truth, q_est, r_est, nsimruns, cores, seed, state_dof, obs_dof,
                          COST_CHOICES[cost_choice], _p(out))
    if rc:
        raise ValueError(f"ko_eval_mt rc={rc}")
    return out


def riccati_trace(dt, q_est, r_est, T) -> np.ndarray:
    out = np.empty((T, 7))
    lib().ko_riccati_trace(dt, q_est, r_est, T, _p(out))
    return out


def eig_transform2(cov) -> np.ndarray:
    t = np.empty(4)
    lib().ko_eig_transform2(_p(_f64(cov).reshape(-1)), _p(t))
    return t.reshape(2, 2)


def mt19937_nth(seed: int, n: int) -> int:
    return int(lib().ko_mt19937_nth(seed, n))


def trial_run_refclock(dt, T, q_truth, r_truth, q_est, r_est, nsimruns, cores, clock0) -> np.ndarray:
    """The four getters of ONE Trial object replayed with the reference's own RNG discipline under the counter
    clock of oracle/ref_shim/fake_clock.h (checked against oracle/_ref in tests/test_oracle_vs_ref.py)."""
    out = np.empty(4)
    rc = lib().ko_trial_run_refclock(dt, T, q_truth, r_truth, q_est, r_est, nsimruns, cores, clock0, _p(out))
    if rc:
        raise ValueError(f"ko_trial_run_refclock rc={rc}")
    return out


def kf_run(q_est, r_est, dt, z) -> np.ndarray:
    z = _f64(z)
    out = np.empty((z.size, 8))
    rc = lib().ko_kf_run(q_est, r_est, dt, z.size, _p(z), _p(out))
    if rc:
        raise ValueError(f"ko_kf_run rc={rc}")
    return out


def gu_table(dt, n=2000) -> np.ndarray:
    out = np.empty((n, 2))
    lib().ko_gu_table(dt, n, _p(out))
    return out


def refclock_draws(dt, T, q_truth, r_truth, clock0, i=0):
    """(x0n[2], w[T][2], v[T]): the noise trial i of trial_run_refclock adds."""
    x0n, w, v = np.empty(2), np.empty((T, 2)), np.empty(T)
    rc = lib().ko_refclock_draws(dt, T, q_truth, r_truth, clock0, i, _p(x0n), _p(w), _p(v))
    if rc:
        raise ValueError(f"ko_refclock_draws rc={rc}")
    return x0n, w, v


# ---- the state-augmented variants, N = 3, 4 (oracle/kfbo_oracle_aug.c; N = 2 reproduces the functions above) ----
def aug_expm(A) -> np.ndarray:
    A = _f64(A)
    E = np.empty_like(A)
    lib().koa_expm(A.shape[0], _p(A), _p(E))
    return E


def aug_discretize(N: int, q: float, dt: float):
    Fd, Qd = np.empty((N, N)), np.empty((N, N))
    if lib().koa_discretize(N, q, dt, _p(Fd), _p(Qd)):
        raise ValueError("koa_discretize: N outside [2, 4]")
    return Fd, Qd


def aug_gvector(N: int, dt: float) -> np.ndarray:
    g = np.empty(N)
    lib().koa_gvector(N, dt, _p(g))
    return g


def aug_cholesky(A) -> np.ndarray:
    A = _f64(A)
    L = np.empty_like(A)
    if lib().koa_cholesky(A.shape[0], _p(A), _p(L)):
        raise ValueError("koa_cholesky: not positive definite")
    return L


def aug_eval_supplied(N, dt, T, q_est, r_est, x0noise, w, v) -> np.ndarray:
    """x0noise[N][B], w[T][N][B], v[T][B] -> per_trial[4][B]."""
    x0noise, w, v = _f64(x0noise), _f64(w), _f64(v)
    B = x0noise.shape[1]
    assert x0noise.shape == (N, B) and w.shape == (T, N, B) and v.shape == (T, B)
    out = np.empty((4, B))
    rc = lib().koa_eval_supplied(N, dt, T, q_est, r_est, B, _p(x0noise), _p(w), _p(v), _p(out))
    if rc:
        raise ValueError(f"koa_eval_supplied rc={rc}")
    return out


def aug_eval_philox(N, seed, stream, dt, T, q_truth, r_truth, q_est, r_est, trial0, ntrials) -> np.ndarray:
    out = np.empty((4, ntrials))
    rc = lib().koa_eval_philox(N, seed, stream, dt, T, q_truth, r_truth, q_est, r_est, trial0, ntrials, _p(out))
    if rc:
        raise ValueError(f"koa_eval_philox rc={rc}")
    return out
