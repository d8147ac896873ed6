"""ctypes binding of libkfbo.so (the C ABI declared in include/kfbo.h).

The library is the product: if it is missing this module raises -- there is no Python or CPU
fallback for the cost evaluation.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_SAMPLE_TIMES = 4
MAX_STEPS = 2000
NCCL_ID_BYTES = 128

JNEES, CNEES, JNIS, CNIS = 0, 1, 2, 3
COST_CHOICES = {"JNEES": JNEES, "CNEES": CNEES, "JNIS": JNIS, "CNIS": CNIS}
SHARD_TRIALS, SHARD_CANDIDATES = 0, 1
OPT_SUPPLIED_KERNEL = 1
OPT_BULK_SMEM_PAD = 2
OPT_ROLLOUT_FORM = 3
OPT_REDUCE = 4
OPT_PEER_FUSE = 5
OPT_PEER_TIMEOUT_MS = 6
OPT_RESERVE_SUPPLIED_TRIALS = 7
OPT_CUDA_GRAPH = 8
OPT_HOST_POLL = 9
OPT_SMALL_LAUNCH = 10
OPT_AUG_LAYOUT = 11
OPT_GROUP_THREADS = 12
OPT_INJECT_FAULT = 13
AUG_LAYOUT_LANES, AUG_LAYOUT_THREAD = 0, 1
STREAM_BITS = 62
TRIAL_STREAM_BASE = 1 << 61
REDUCE_AUTO, REDUCE_NCCL = 0, 1
PEER_FUSE_NEVER, PEER_FUSE_AUTO, PEER_FUSE_ALWAYS = 0, 1, 2
MAX_PEERS = 8
PEER_HANDLE_BYTES = 64
# kfbo_last_reduce_path
REDUCE_PATHS = {0: "single rank", 1: "ncclAllReduce", 2: "peer-memory exchange (kernel behind the rollout)",
                3: "peer-memory exchange fused into the rollout kernel"}
FORM_REFERENCE = 0
SUPPLIED_AUTO, SUPPLIED_LDG, SUPPLIED_BULK = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
# KFBO_LIB_PATH: load another build of the same library (kernel tuning experiments only)
LIB_PATH = os.environ.get("KFBO_LIB_PATH") or os.path.join(_HERE, "libkfbo.so")


class KfboParams(C.Structure):
    """struct kfbo_params (include/kfbo.h)."""
    _fields_ = [
        ("n_sample_times", C.c_int32),
        ("dt", C.c_double * MAX_SAMPLE_TIMES),
        ("max_iterations", C.c_int32 * MAX_SAMPLE_TIMES),
        ("nsimruns", C.c_int32),
        ("cpu_core_number", C.c_int32),
        ("state_dof", C.c_int32),
        ("observation_dof", C.c_int32),
        ("pnoise_truth", C.c_double),
        ("onoise_truth", C.c_double),
        ("cost_choice", C.c_int32),
        ("device", C.c_int32),
        ("seed", C.c_uint64),
        ("max_candidates", C.c_int32),
        ("state_dim", C.c_int32),
    ]


class KfboResult(C.Structure):
    """struct kfbo_result (include/kfbo.h)."""
    _fields_ = [
        ("average_nees", C.c_double * MAX_SAMPLE_TIMES),
        ("average_nis", C.c_double * MAX_SAMPLE_TIMES),
        ("average_var_nees", C.c_double * MAX_SAMPLE_TIMES),
        ("average_var_nis", C.c_double * MAX_SAMPLE_TIMES),
        ("cost", C.c_double * MAX_SAMPLE_TIMES),
        ("max_cost", C.c_double),
        ("index_max", C.c_int32),
        ("reserved", C.c_int32),
    ]


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); this table is also what tests check against include/kfbo.h
SIGNATURES = {
    "kfbo_default_params": (None, [C.POINTER(KfboParams)]),
    "kfbo_create": (C.c_int, [C.POINTER(KfboParams), C.POINTER(_vp)]),
    "kfbo_destroy": (None, [_vp]),
    "kfbo_last_error": (C.c_char_p, [_vp]),
    "kfbo_eval": (C.c_int, [_vp, _dp, C.c_int32, _dp, C.c_int32, C.POINTER(KfboResult)]),
    "kfbo_eval_batch": (C.c_int, [_vp, C.c_int32, _dp, _dp, C.POINTER(KfboResult)]),
    "kfbo_eval_batch_costs": (C.c_int, [_vp, C.c_int32, _dp, _dp, _dp, _dp, C.POINTER(C.c_int32)]),
    "kfbo_last_transfer_bytes": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "kfbo_eval_supplied": (C.c_int, [_vp, C.c_int32, C.c_int32, _dp, _dp, C.c_int64, _dp, _dp, _dp, _dp, _dp]),
    "kfbo_eval_supplied_dev": (C.c_int, [_vp, C.c_int32, C.c_int32, _dp, _dp, C.c_int64, _vp, _vp, _vp, _vp, _dp]),
    "kfbo_eval_philox_trials": (C.c_int, [_vp, C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_uint32, C.c_int64, _dp]),
    "kfbo_trace_trial": (C.c_int, [_vp, C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_uint32, _dp]),
    "kfbo_set_stream_base": (C.c_int, [_vp, C.c_uint64]),
    "kfbo_set_eval_counter": (C.c_int, [_vp, C.c_uint64]),
    "kfbo_last_sums": (C.c_int, [_vp, _dp, C.c_int32]),
    "kfbo_kernel_ms_total": (C.c_int, [_vp, _dp, C.POINTER(C.c_int64), C.c_int32]),
    "kfbo_last_kernel_ms": (C.c_int, [_vp, _dp, C.POINTER(C.c_int32)]),
    "kfbo_last_host_us": (C.c_int, [_vp, _dp]),
    "kfbo_set_cuda_stream": (C.c_int, [_vp, _vp]),
    "kfbo_set_option": (C.c_int, [_vp, C.c_int32, C.c_int64]),
    "kfbo_nccl_unique_id": (C.c_int, [_vp]),
    "kfbo_comm_init": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32]),
    "kfbo_peer_export": (C.c_int, [_vp, _vp]),
    "kfbo_peer_attach": (C.c_int, [_vp, _vp, C.c_int32]),
    "kfbo_last_reduce_path": (C.c_int, [_vp]),
    "kfbo_group_create": (C.c_int, [C.POINTER(KfboParams), C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(_vp)]),
    "kfbo_group_destroy": (None, [_vp]),
    "kfbo_group_last_error": (C.c_char_p, [_vp]),
    "kfbo_group_size": (C.c_int32, [_vp]),
    "kfbo_group_handle": (_vp, [_vp, C.c_int32]),
    "kfbo_group_eval": (C.c_int, [_vp, _dp, C.c_int32, _dp, C.c_int32, C.POINTER(KfboResult)]),
    "kfbo_group_eval_batch": (C.c_int, [_vp, C.c_int32, _dp, _dp, C.POINTER(KfboResult)]),
    "kfbo_group_eval_batch_costs": (C.c_int, [_vp, C.c_int32, _dp, _dp, _dp, _dp, C.POINTER(C.c_int32)]),
    "kfbo_group_last_sums": (C.c_int, [_vp, _dp, C.c_int32]),
    "kfbo_group_last_kernel_ms": (C.c_int, [_vp, _dp, C.POINTER(C.c_int32)]),
    "kfbo_group_set_option": (C.c_int, [_vp, C.c_int32, C.c_int64]),
    "kfbo_group_set_stream_base": (C.c_int, [_vp, C.c_uint64]),
    "kfbo_group_set_eval_counter": (C.c_int, [_vp, C.c_uint64]),
    "kfbo_shard_range": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "kfbo_discretize": (C.c_int, [C.c_double, C.c_double, _dp, _dp]),
    "kfbo_control_table": (C.c_int, [C.c_double, C.c_int32, _dp]),
    "kfbo_finalize": (C.c_int, [C.POINTER(KfboParams), _dp, C.POINTER(KfboResult)]),
    "kfbo_effective_trials": (C.c_int64, [C.POINTER(KfboParams)]),
    "kfbo_measure_fp64_peak": (C.c_int, [C.c_int32, C.c_double, _dp, _dp]),
    "kfbo_version": (C.c_char_p, []),
}

_lib = None


def load() -> C.CDLL:
    """Load libkfbo.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C kf_bayesopt_b200/csrc` "
                "(there is no CPU fallback for the cost evaluation)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
