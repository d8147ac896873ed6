"""Host-side mirror of the reference's cost-evaluation interface, on top of the C ABI.

Names follow the reference (all file:line relative to the reference tree):
  ReadParaVehicle     include/read_para_vehicle.h:9-49, read_para_vehicle.cpp:119-158
  Trial               include/trial.h:7-59, trial.cpp:12-131
  RunTrial            trial_node.cpp:13-142 (ProcessNoiseCallback: "noise" -> "cost")
  Float64MultiArray / Float64   the two std_msgs types on those topics (robot1d_bayesopt.cpp:53-92)

Everything numeric happens in libkfbo.so (CUDA, sm_100a); this module only marshals.
"""
from __future__ import annotations

import ctypes as C
import itertools
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import COST_CHOICES, KfboParams, KfboResult, MAX_SAMPLE_TIMES, SHARD_CANDIDATES, SHARD_TRIALS

_dp = C.POINTER(C.c_double)

# struct kfbo_result as a numpy record (same layout as _lib.KfboResult; checked in tests/test_host_abi.py)
RESULT_DTYPE = np.dtype([("average_nees", "f8", (MAX_SAMPLE_TIMES,)), ("average_nis", "f8", (MAX_SAMPLE_TIMES,)),
                         ("average_var_nees", "f8", (MAX_SAMPLE_TIMES,)), ("average_var_nis", "f8", (MAX_SAMPLE_TIMES,)),
                         ("cost", "f8", (MAX_SAMPLE_TIMES,)), ("max_cost", "f8"), ("index_max", "i4"), ("reserved", "i4")])

# trial_node.cpp:105-114: pass 0 uses the yaml (dt, t_end); pass 1 overwrites them with these.
CODE_SECOND_SAMPLE_TIME = (1.0, 201.0)
# the variant commented out at trial_node.cpp:109-112 and described in README.md:96 / BASELINE.json
README_SECOND_SAMPLE_TIME = (0.5, 105.5)


class KfboError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"kfbo error {code}: {msg}")
        self.code = code


@dataclass
class ReadParaVehicle:
    """The fields of ReadParaVehicle (include/read_para_vehicle.h:9-49) with the yaml defaults
    (config/vehicleParams.yaml)."""
    cpu_core_number_: int = 10
    state_dof_: int = 2
    observation_dof_: int = 1
    t_start_: float = 0.0
    dt_: float = 0.1
    t_end_: float = 20.1
    xi0_: float = 0.0          # read but never used by the hot path (trial.cpp:31 hard-codes x0 = 0)
    xidot0_: float = 0.0
    disturbance_: List[float] = field(default_factory=list)
    P0_: str = "identity"      # read but never used (trial.cpp:32 hard-codes P0 = I)
    nsimruns_: int = 200
    max_iterations_: int = 201
    pnoise_: List[float] = field(default_factory=lambda: [1.0])
    onoise_: List[float] = field(default_factory=lambda: [0.1])
    optimization_choice_: str = "all"
    cost_choice_: str = "JNEES"

    def __post_init__(self):
        self.update_max_iterations()

    def update_max_iterations(self) -> None:
        # read_para_vehicle.cpp:145 / trial_node.cpp:113: truncating double -> int
        self.max_iterations_ = int((self.t_end_ - self.t_start_) / self.dt_)

    @property
    def dim_pn_(self) -> int:
        return len(self.pnoise_)

    @property
    def dim_on_(self) -> int:
        return len(self.onoise_)

    @classmethod
    def from_yaml(cls, path: str) -> "ReadParaVehicle":
        """Read config/vehicleParams.yaml unchanged (keys of read_para_vehicle.cpp:121-157)."""
        import yaml
        with open(path) as f:
            y = yaml.safe_load(f)
        rv = cls(
            cpu_core_number_=int(y.get("CPUcoreNumber", 10)),
            state_dof_=int(y.get("stateDOF", 2)),
            observation_dof_=int(y.get("observationDOF", 1)),
            t_start_=float(y.get("t_start", 0.0)),
            dt_=float(y.get("dt", 0.1)),
            t_end_=float(y.get("t_end", 20.1)),
            xi0_=float(y.get("xi0", 0.0)),
            xidot0_=float(y.get("xidot0", 0.0)),
            disturbance_=list(y.get("disturbance", []) or []),
            P0_=str(y.get("P0", "identity")),
            nsimruns_=int(y.get("Nsimruns", 10)),
            pnoise_=[float(v) for v in y["pnoise"]],
            onoise_=[float(v) for v in y["onoise"]],
            optimization_choice_=str(y.get("optimizationChoice", "all")),
            cost_choice_=str(y.get("costChoice", "JNEES")),
        )
        return rv

    def copy(self) -> "ReadParaVehicle":
        import copy
        return copy.deepcopy(self)


class Result:
    """struct kfbo_result as plain Python.  The arrays are cut out of the C struct on first use: the 'noise' -> 'cost' call
    itself only needs max_cost (a 70 us evaluation should not spend 8 of them building arrays nobody reads)."""
    __slots__ = ("_c", "_D", "_rows", "max_cost", "index_max")

    def __init__(self, average_nees=None, average_nis=None, average_var_nees=None, average_var_nis=None, cost=None,
                 max_cost=float("nan"), index_max=0):
        self._c, self._D = None, 0
        self._rows = None if average_nees is None else [np.asarray(a, dtype=np.float64) for a in
                                                         (average_nees, average_nis, average_var_nees, average_var_nis, cost)]
        self.max_cost, self.index_max = float(max_cost), int(index_max)

    @classmethod
    def from_c(cls, r: KfboResult, D: int) -> "Result":
        self = cls.__new__(cls)
        self._c, self._D, self._rows = r, D, None
        self.max_cost, self.index_max = r.max_cost, r.index_max
        return self

    def _get(self, i):
        if self._rows is None:
            # one copy of the five [MAX_SAMPLE_TIMES] arrays at the head of the struct; the fields are its rows
            a = np.frombuffer(self._c, dtype=np.float64, count=5 * MAX_SAMPLE_TIMES).reshape(5, MAX_SAMPLE_TIMES)[:, :self._D].copy()
            self._rows = [a[0], a[1], a[2], a[3], a[4]]
        return self._rows[i]

    average_nees = property(lambda self: self._get(0))
    average_nis = property(lambda self: self._get(1))
    average_var_nees = property(lambda self: self._get(2))
    average_var_nis = property(lambda self: self._get(3))
    cost = property(lambda self: self._get(4))          # pic_max of trial_node.cpp:50

    def __repr__(self):
        return f"Result(cost={self.cost}, max_cost={self.max_cost}, index_max={self.index_max})"


def sample_times_from(rv: ReadParaVehicle, second: Optional[Tuple[float, float]] = CODE_SECOND_SAMPLE_TIME):
    """The (dt, max_iterations) list of RunTrial::ProcessNoiseCallback's two passes: pass 0 from the
    yaml, pass 1 from `second` = (dt, t_end) (trial_node.cpp:105-114).  second=None -> single pass
    (src/test_trial.cpp)."""
    out = [(rv.dt_, int((rv.t_end_ - rv.t_start_) / rv.dt_))]
    if second is not None:
        dt2, t_end2 = second
        out.append((dt2, int((t_end2 - rv.t_start_) / dt2)))
    return out


def make_params(rv_gt: ReadParaVehicle, sample_times: Sequence[Tuple[float, int]], seed: int = 0x4B46424F20260101,
                max_candidates: int = 1, device: int = -1, cost_choice: Optional[str] = None, state_dim: int = 0) -> KfboParams:
    if not 1 <= len(sample_times) <= MAX_SAMPLE_TIMES:
        raise ValueError(f"1..{MAX_SAMPLE_TIMES} sample times, got {len(sample_times)}")
    p = KfboParams()
    _lib.load().kfbo_default_params(C.byref(p))
    p.n_sample_times = len(sample_times)
    for k, (dt, T) in enumerate(sample_times):
        p.dt[k], p.max_iterations[k] = float(dt), int(T)
    p.nsimruns = rv_gt.nsimruns_
    p.cpu_core_number = rv_gt.cpu_core_number_
    p.state_dof, p.observation_dof = rv_gt.state_dof_, rv_gt.observation_dof_
    p.pnoise_truth, p.onoise_truth = float(rv_gt.pnoise_[0]), float(rv_gt.onoise_[0])
    p.cost_choice = COST_CHOICES[cost_choice or rv_gt.cost_choice_]
    p.device, p.seed, p.max_candidates = device, seed & 0xFFFFFFFFFFFFFFFF, max_candidates
    p.state_dim = state_dim        # 0 / 2: the reference's model; 3, 4: the state-augmented variants (kfbo.h)
    return p


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _small(a):
    """A short host sequence as a ctypes double array (0.6 us; numpy conversion + pointer cast are 5 us, which shows
    in the 'noise' -> 'cost' call of the default workload, 80 us in all)."""
    if not hasattr(a, "__len__") or getattr(a, "ndim", 1) == 0:     # a bare number / 0-d array, as the numpy path accepted
        a = (float(a),)
    return (C.c_double * len(a))(*a)


class CostEvaluator:
    """Owns one kfbo_handle: the device-side batch launcher that replaces the CPUcoreNumber thread
    pool of trial_node.cpp:54-67."""

    def __init__(self, rv_gt: Optional[ReadParaVehicle] = None, sample_times: Optional[Sequence[Tuple[float, int]]] = None,
                 seed: int = 0x4B46424F20260101, max_candidates: int = 1, device: int = -1,
                 cost_choice: Optional[str] = None, state_dim: int = 0):
        self._lib = _lib.load()
        self.rv_gt = rv_gt or ReadParaVehicle()
        self.sample_times = list(sample_times) if sample_times is not None else sample_times_from(self.rv_gt)
        self.params = make_params(self.rv_gt, self.sample_times, seed, max_candidates, device, cost_choice, state_dim)
        self.D = len(self.sample_times)
        self.N = state_dim if state_dim >= 2 else 2      # state dimension: rows of the supplied noise arrays
        self._h = C.c_void_p()
        rc = self._lib.kfbo_create(C.byref(self.params), C.byref(self._h))
        if rc:
            raise KfboError(rc, (self._lib.kfbo_last_error(None) or b"").decode())

    # -- lifetime
    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.kfbo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc:
            raise KfboError(rc, (self._lib.kfbo_last_error(self._h) or b"").decode())

    @property
    def effective_trials(self) -> int:
        return int(self._lib.kfbo_effective_trials(C.byref(self.params)))

    # -- the "noise" -> "cost" evaluation
    def eval(self, pn: Sequence[float], on: Sequence[float]) -> Result:
        r = KfboResult()
        cpn, con = _small(pn), _small(on)
        rc = self._lib.kfbo_eval(self._h, cpn, len(cpn), con, len(con), r)
        if rc:
            self._check(rc)
        return Result.from_c(r, self.D)

    def eval_batch(self, q_est: Sequence[float], r_est: Sequence[float]) -> List[Result]:
        q, r = _f64(q_est), _f64(r_est)
        if q.shape != r.shape or q.ndim != 1:
            raise ValueError("q_est and r_est must be 1-D arrays of equal length")
        res = (KfboResult * q.size)()
        self._check(self._lib.kfbo_eval_batch(self._h, q.size, _ptr(q), _ptr(r), res))
        return [Result.from_c(x, self.D) for x in res]

    def eval_batch_raw(self, q_est, r_est) -> np.ndarray:
        """The kfbo_result array of a candidate batch as ONE structured numpy array (fields of struct
        kfbo_result; no per-candidate Python objects -- a 4096-candidate sweep costs microseconds of host time)."""
        q, r = _f64(q_est), _f64(r_est)
        if q.shape != r.shape or q.ndim != 1:
            raise ValueError("q_est and r_est must be 1-D arrays of equal length")
        res = np.empty(q.size, dtype=RESULT_DTYPE)
        self._check(self._lib.kfbo_eval_batch(self._h, q.size, _ptr(q), _ptr(r), res.ctypes.data_as(C.POINTER(KfboResult))))
        return res

    def eval_batch_costs(self, q_est, r_est, with_index: bool = False):
        """(cost[C][D], max_cost[C]) of a candidate batch (with_index: also index_max[C]), assembled on the device
        (kfbo_eval_batch_costs): 16 bytes per candidate in, 8 (D + 1) + 4 out, no per-candidate host work."""
        q, r = _f64(q_est), _f64(r_est)
        if q.shape != r.shape or q.ndim != 1:
            raise ValueError("q_est and r_est must be 1-D arrays of equal length")
        cost, mx, idx = np.empty((q.size, self.D)), np.empty(q.size), np.empty(q.size, dtype=np.int32)
        self._check(self._lib.kfbo_eval_batch_costs(self._h, q.size, _ptr(q), _ptr(r), _ptr(cost), _ptr(mx),
                                                    idx.ctypes.data_as(C.POINTER(C.c_int32))))
        return (cost, mx, idx) if with_index else (cost, mx)

    def last_transfer_bytes(self) -> Tuple[int, int]:
        """(host -> device, device -> host) bytes of the last evaluation."""
        a, b = C.c_int64(), C.c_int64()
        self._check(self._lib.kfbo_last_transfer_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_sums(self, n_candidates: int = 1) -> np.ndarray:
        out = np.empty((n_candidates, self.D, 4))
        self._check(self._lib.kfbo_last_sums(self._h, _ptr(out), n_candidates))
        return out

    def last_kernel_ms(self) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_int32()
        self._check(self._lib.kfbo_last_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def kernel_ms_total(self, reset: bool = False) -> Tuple[float, int]:
        """(sum of the rollout-kernel device times, number of evaluations) since the last reset -- what a benchmark loop reads
        once at its end instead of last_kernel_ms() between evaluations.  The first call switches the handle to accumulating."""
        ms, n = C.c_double(), C.c_int64()
        self._check(self._lib.kfbo_kernel_ms_total(self._h, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value

    def last_host_us(self) -> dict:
        """Host wall time (us) of the phases of the last eval / eval_batch on this rank."""
        us = np.empty(5)
        self._check(self._lib.kfbo_last_host_us(self._h, _ptr(us)))
        return dict(zip(("build_cases", "enqueue", "device_wait", "cost_assembly", "exchange_device"), us.tolist()))

    # -- exact-parity entry points
    def eval_supplied(self, k: int, q_est, r_est, x0noise, w, v, per_trial: bool = True):
        """Returns (per_trial[C][4][B] or None, sums[C][4]); x0noise[N][B], w[T][N][B], v[T][B] (N = 2 unless state_dim says otherwise)."""
        q, r = np.atleast_1d(_f64(q_est)), np.atleast_1d(_f64(r_est))
        x0noise, w, v = _f64(x0noise), _f64(w), _f64(v)
        T = self.sample_times[k][1]
        B = x0noise.shape[-1]
        if x0noise.shape != (self.N, B) or w.shape != (T, self.N, B) or v.shape != (T, B):
            raise ValueError(f"expected x0noise[{self.N}][{B}], w[{T}][{self.N}][{B}], v[{T}][{B}]")
        pt = np.empty((q.size, 4, B)) if per_trial else None
        sums = np.empty((q.size, 4))
        self._check(self._lib.kfbo_eval_supplied(self._h, k, q.size, _ptr(q), _ptr(r), B, _ptr(x0noise), _ptr(w), _ptr(v),
                                                 _ptr(pt) if per_trial else None, _ptr(sums)))
        return pt, sums

    def eval_supplied_dev(self, k: int, q_est, r_est, B: int, d_x0noise: int, d_w: int, d_v: int, d_per_trial: int = 0):
        """Device-pointer variant (ints from tensor.data_ptr()); returns sums[C][4]."""
        q, r = np.atleast_1d(_f64(q_est)), np.atleast_1d(_f64(r_est))
        sums = np.empty((q.size, 4))
        self._check(self._lib.kfbo_eval_supplied_dev(self._h, k, q.size, _ptr(q), _ptr(r), B, d_x0noise, d_w, d_v,
                                                     d_per_trial or None, _ptr(sums)))
        return sums

    def eval_philox_trials(self, k: int, q_est: float, r_est: float, stream: int, trial0: int, ntrials: int) -> np.ndarray:
        out = np.empty((4, ntrials))
        self._check(self._lib.kfbo_eval_philox_trials(self._h, k, q_est, r_est, stream, trial0, ntrials, _ptr(out)))
        return out

    def trace_trial(self, k: int, q_est: float, r_est: float, stream: int, trial: int) -> np.ndarray:
        """The Nsimruns == 1 record of trial.cpp:69-78: [T][9] = time, e1, e2, lb1, ub1, lb2, ub2, nees, nis."""
        out = np.empty((self.sample_times[k][1], 9))
        self._check(self._lib.kfbo_trace_trial(self._h, k, q_est, r_est, stream, trial, _ptr(out)))
        return out

    # -- Philox stream control / CUDA stream / NCCL
    def set_stream_base(self, base: int) -> None:
        self._check(self._lib.kfbo_set_stream_base(self._h, base))

    def set_eval_counter(self, e: int) -> None:
        self._check(self._lib.kfbo_set_eval_counter(self._h, e))

    def set_option(self, option: int, value: int) -> None:
        self._check(self._lib.kfbo_set_option(self._h, option, value))

    def set_cuda_stream(self, stream_ptr: int) -> None:
        self._check(self._lib.kfbo_set_cuda_stream(self._h, stream_ptr or None))

    def comm_init(self, id_bytes: bytes, rank: int, nranks: int, shard_mode: int = SHARD_TRIALS) -> None:
        buf = C.create_string_buffer(bytes(id_bytes), _lib.NCCL_ID_BYTES)
        self._check(self._lib.kfbo_comm_init(self._h, buf, rank, nranks, shard_mode))

    # -- peer-memory exchange of the cost sums (kfbo_peer_export / kfbo_peer_attach in include/kfbo.h)
    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        self._check(self._lib.kfbo_peer_export(self._h, buf))
        return buf.raw

    def peer_attach(self, handles) -> None:
        """handles: the peer_export() bytes of ALL ranks, in rank order."""
        handles = [bytes(x) for x in handles]
        if any(len(x) != _lib.PEER_HANDLE_BYTES for x in handles):
            raise ValueError("peer handles are PEER_HANDLE_BYTES each")
        blob = b"".join(handles)
        buf = C.create_string_buffer(blob, len(blob))
        self._check(self._lib.kfbo_peer_attach(self._h, buf, len(handles)))   # the library checks the count against its ranks

    def last_reduce_path(self) -> str:
        return _lib.REDUCE_PATHS.get(self._lib.kfbo_last_reduce_path(self._h), "?")


class CostEvaluatorGroup:
    """Owns one kfbo_group: ONE process driving `n_gpus` devices of the box (kfbo_group_* in include/kfbo.h) -- the
    single-process counterpart of the one-process-per-GPU mode (dist.init_comm); same sharding, same Philox streams,
    same results."""

    def __init__(self, rv_gt: Optional[ReadParaVehicle] = None, sample_times: Optional[Sequence[Tuple[float, int]]] = None,
                 n_gpus: int = 1, devices: Optional[Sequence[int]] = None, shard_mode: int = SHARD_TRIALS,
                 seed: int = 0x4B46424F20260101, max_candidates: int = 1, cost_choice: Optional[str] = None):
        self._lib = _lib.load()
        self.rv_gt = rv_gt or ReadParaVehicle()
        self.sample_times = list(sample_times) if sample_times is not None else sample_times_from(self.rv_gt)
        self.params = make_params(self.rv_gt, self.sample_times, seed, max_candidates, -1, cost_choice)
        self.D, self.n_gpus = len(self.sample_times), n_gpus
        self._g = C.c_void_p()
        devs = (C.c_int32 * n_gpus)(*devices) if devices is not None else None
        rc = self._lib.kfbo_group_create(C.byref(self.params), n_gpus, devs, shard_mode, C.byref(self._g))
        if rc:
            raise KfboError(rc, (self._lib.kfbo_group_last_error(None) or b"").decode())

    def close(self) -> None:
        if getattr(self, "_g", None) and self._g.value:
            self._lib.kfbo_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc:
            raise KfboError(rc, (self._lib.kfbo_group_last_error(self._g) or b"").decode())

    def eval(self, pn: Sequence[float], on: Sequence[float]) -> Result:
        r = KfboResult()
        cpn, con = _small(pn), _small(on)
        rc = self._lib.kfbo_group_eval(self._g, cpn, len(cpn), con, len(con), r)
        if rc:
            self._check(rc)
        return Result.from_c(r, self.D)

    def eval_batch_raw(self, q_est, r_est) -> np.ndarray:
        q, r = _f64(q_est), _f64(r_est)
        if q.shape != r.shape or q.ndim != 1:
            raise ValueError("q_est and r_est must be 1-D arrays of equal length")
        res = np.empty(q.size, dtype=RESULT_DTYPE)
        self._check(self._lib.kfbo_group_eval_batch(self._g, q.size, _ptr(q), _ptr(r), res.ctypes.data_as(C.POINTER(KfboResult))))
        return res

    def eval_batch_costs(self, q_est, r_est, with_index: bool = False):
        """(cost[C][D], max_cost[C][, index_max[C]]), assembled on the collecting device (kfbo_group_eval_batch_costs)."""
        q, r = _f64(q_est), _f64(r_est)
        if q.shape != r.shape or q.ndim != 1:
            raise ValueError("q_est and r_est must be 1-D arrays of equal length")
        cost, mx, idx = np.empty((q.size, self.D)), np.empty(q.size), np.empty(q.size, dtype=np.int32)
        self._check(self._lib.kfbo_group_eval_batch_costs(self._g, q.size, _ptr(q), _ptr(r), _ptr(cost), _ptr(mx),
                                                          idx.ctypes.data_as(C.POINTER(C.c_int32))))
        return (cost, mx, idx) if with_index else (cost, mx)

    def last_sums(self, n_candidates: int = 1) -> np.ndarray:
        out = np.empty((n_candidates, self.D, 4))
        self._check(self._lib.kfbo_group_last_sums(self._g, _ptr(out), n_candidates))
        return out

    def last_kernel_ms(self) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_int32()
        self._check(self._lib.kfbo_group_last_kernel_ms(self._g, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_option(self, option: int, value: int) -> None:
        self._check(self._lib.kfbo_group_set_option(self._g, option, value))

    def set_device_option(self, i: int, option: int, value: int) -> None:
        """An option of ONE device's handle of the group (kfbo_group_handle + kfbo_set_option) -- the tests' fault injection."""
        h = self._lib.kfbo_group_handle(self._g, i)
        if not h:
            raise ValueError(f"the group has no device {i}")
        rc = self._lib.kfbo_set_option(h, option, value)
        if rc:
            raise KfboError(rc, (self._lib.kfbo_last_error(h) or b"").decode())

    def last_reduce_path(self) -> str:
        return _lib.REDUCE_PATHS.get(self._lib.kfbo_last_reduce_path(self._lib.kfbo_group_handle(self._g, 0)), "?")

    def set_stream_base(self, base: int) -> None:
        self._check(self._lib.kfbo_group_set_stream_base(self._g, base))

    def set_eval_counter(self, e: int) -> None:
        self._check(self._lib.kfbo_group_set_eval_counter(self._g, e))


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(_lib.NCCL_ID_BYTES)
    rc = _lib.load().kfbo_nccl_unique_id(buf)
    if rc:
        raise KfboError(rc, (_lib.load().kfbo_last_error(None) or b"").decode())
    return buf.raw


def finalize(params: KfboParams, sums) -> Result:
    """Host-only: raw sums[D][4] -> averages, costs, max (trial.cpp:110-113, trial_node.cpp:70-119)."""
    sums = _f64(sums)
    r = KfboResult()
    rc = _lib.load().kfbo_finalize(C.byref(params), _ptr(sums), C.byref(r))
    if rc:
        raise KfboError(rc, "kfbo_finalize: invalid params")
    return Result.from_c(r, params.n_sample_times)


def shard_range(n: int, rank: int, nranks: int) -> Tuple[int, int]:
    b, e = C.c_int64(), C.c_int64()
    rc = _lib.load().kfbo_shard_range(n, rank, nranks, C.byref(b), C.byref(e))
    if rc:
        raise KfboError(rc, "kfbo_shard_range: invalid argument")
    return b.value, e.value


def discretize(pnoise0: float, dt: float):
    Fd, Qd = np.empty(4), np.empty(4)
    _lib.load().kfbo_discretize(pnoise0, dt, _ptr(Fd), _ptr(Qd))
    return Fd.reshape(2, 2), Qd.reshape(2, 2)


def control_table(dt: float, n: int = 2000) -> np.ndarray:
    u = np.empty(n)
    _lib.load().kfbo_control_table(dt, n, _ptr(u))
    return u


def measure_fp64_peak(device: int = -1, millis: float = 200.0, both: bool = False):
    """Sustained DFMA throughput (TFLOP/s) of the register-multiplier chain x = fma(x, y, U) -- the FP64 roofline
    denominator; both=True: (that, the same with both other operands uniform: x = fma(x, U, U), ~10 % slower)."""
    t, u = C.c_double(), C.c_double()
    rc = _lib.load().kfbo_measure_fp64_peak(device, millis, C.byref(t), C.byref(u) if both else None)
    if rc:
        raise KfboError(rc, (_lib.load().kfbo_last_error(None) or b"").decode())
    return (t.value, u.value) if both else t.value


# -------------------------------------------------------------------------------------------
# Reference-shaped interface
# -------------------------------------------------------------------------------------------
_trial_streams = itertools.count(_lib.TRIAL_STREAM_BASE)   # the Trial adapters' own range of the 62-bit stream ids


class Trial:
    """Drop-in for the reference's Trial (include/trial.h:7-59): one Trial is ONE thread's share of
    the Monte Carlo batch -- nsimruns_/cpu_core_number_ trials (trial.cpp:28) -- run on the GPU."""

    def __init__(self, simulator_parameters: ReadParaVehicle, estimator_parameters: ReadParaVehicle, device: int = -1,
                 seed: int = 0x4B46424F20260101):
        self.param_vehicle_ = simulator_parameters.copy()
        self.estimator_parameters_ = estimator_parameters.copy()
        self._device, self._seed = device, seed
        self.average_nees_ = self.average_nis_ = self.average_var_nees_ = self.average_var_nis_ = float("nan")

    def Run(self) -> None:
        sim, est = self.param_vehicle_, self.estimator_parameters_
        n = sim.nsimruns_ // sim.cpu_core_number_                       # trial.cpp:28
        T = sim.max_iterations_
        k_average = float(T * sim.nsimruns_ // sim.cpu_core_number_)    # trial.cpp:19
        one = sim.copy()
        one.nsimruns_, one.cpu_core_number_ = n, 1
        with CostEvaluator(one, [(sim.dt_, T)], seed=self._seed, device=self._device) as ev:
            ev.set_stream_base(next(_trial_streams))                    # every Trial draws its own noise
            ev.eval(est.pnoise_, est.onoise_)
            s = ev.last_sums(1)[0, 0]
        self.average_var_nees_ = s[2] * sim.cpu_core_number_ / sim.nsimruns_   # trial.cpp:110
        self.average_var_nis_ = s[3] * sim.cpu_core_number_ / sim.nsimruns_    # trial.cpp:111
        self.average_nees_ = s[0] / k_average                                    # trial.cpp:112
        self.average_nis_ = s[1] / k_average                                     # trial.cpp:113

    def get_average_nees(self) -> float:
        return self.average_nees_

    def get_average_nis(self) -> float:
        return self.average_nis_

    def get_average_var_nees(self) -> float:
        return self.average_var_nees_

    def get_average_var_nis(self) -> float:
        return self.average_var_nis_


@dataclass
class MultiArrayDimension:
    label: str = ""
    size: int = 0
    stride: int = 0


@dataclass
class MultiArrayLayout:
    dim: List[MultiArrayDimension] = field(default_factory=list)
    data_offset: int = 0


@dataclass
class Float64MultiArray:
    """std_msgs/Float64MultiArray as published on "noise" (robot1d_bayesopt.cpp:53-88)."""
    layout: MultiArrayLayout = field(default_factory=MultiArrayLayout)
    data: List[float] = field(default_factory=list)

    @classmethod
    def noise(cls, pn: Sequence[float], on: Sequence[float]) -> "Float64MultiArray":
        m = cls()
        m.layout.dim = [MultiArrayDimension("processNoiseDimension", len(pn)),
                        MultiArrayDimension("measurmentNoiseDimension", len(on))]
        m.data = list(pn) + list(on)
        return m


@dataclass
class Float64:
    """std_msgs/Float64 as published on "cost" (trial_node.cpp:47,127-129)."""
    data: float = 0.0


class RunTrial:
    """ROS-free RunTrial (trial_node.cpp:13-142): ProcessNoiseCallback takes the "noise" message and
    returns the "cost" message; with `publish` set it is also called with it (pub_cost_.publish)."""

    def __init__(self, rv: Optional[ReadParaVehicle] = None, second_sample_time=CODE_SECOND_SAMPLE_TIME,
                 n_init_samples: int = 20, publish=None, device: int = -1, seed: int = 0x4B46424F20260101,
                 verbose: bool = False, evaluator: Optional[CostEvaluator] = None):
        self.rv_ = rv or ReadParaVehicle()
        self.n_init_samples_ = n_init_samples                     # rb_.n_init_samples_ (trial_node.cpp:121)
        self.it_count_ = self.dt0_count_ = self.dt1_count_ = 0    # trial_node.cpp:140
        self.publish_, self.verbose_ = publish, verbose
        self.last_result: Optional[Result] = None
        # the device-side batch launcher; built once, not per callback like the reference's vector<Trial>
        self._ev = evaluator or CostEvaluator(self.rv_, sample_times_from(self.rv_, second_sample_time), seed=seed,
                                              device=device)

    def ProcessNoiseCallback(self, msg: Float64MultiArray) -> Float64:
        npn = msg.layout.dim[0].size                               # trial_node.cpp:29
        pn, on = msg.data[:npn], msg.data[npn:]                    # :32-33
        res = self._ev.eval(pn, on)
        self.last_result = res
        self.it_count_ += 1                                        # :121-126
        if self.it_count_ > self.n_init_samples_:
            if res.index_max == 1:
                self.dt1_count_ += 1
            else:
                self.dt0_count_ += 1
        if self.verbose_:
            print("process noise", *pn)
            print("observation noise", on[0])
            print(",".join(str(c) for c in res.cost))
        cost = Float64(res.max_cost)                               # :127
        if self.publish_ is not None:
            self.publish_(cost)                                    # :129
        return cost
