"""kf_bayesopt_b200 -- B200-native (sm_100a) Monte Carlo cost evaluation for kf_bayesopt.

Only the hot path of the reference lives here: the truth/measurement rollout, the Kalman
predict/update recursion and the NEES/NIS -> J_NEES/J_NIS accumulation, as hand-written CUDA
behind the C ABI of include/kfbo.h, plus the host-side mirror of the reference's Trial /
"noise"->"cost" interface.  The CUDA library is mandatory: there is no CPU fallback.
"""
from ._lib import (CNEES, CNIS, COST_CHOICES, JNEES, JNIS, LIB_PATH, SHARD_CANDIDATES, SHARD_TRIALS, KfboParams,
                   KfboResult)
from .evaluator import (CODE_SECOND_SAMPLE_TIME, README_SECOND_SAMPLE_TIME, CostEvaluator, CostEvaluatorGroup, Float64, Float64MultiArray,
                        KfboError, ReadParaVehicle, Result, RunTrial, Trial, control_table, discretize, finalize,
                        make_params, measure_fp64_peak, nccl_unique_id, sample_times_from, shard_range)

__all__ = [
    "CostEvaluator", "CostEvaluatorGroup", "Trial", "RunTrial", "ReadParaVehicle", "Result", "Float64", "Float64MultiArray", "KfboError",
    "KfboParams", "KfboResult", "make_params", "sample_times_from", "finalize", "shard_range", "discretize",
    "control_table", "measure_fp64_peak", "nccl_unique_id", "CODE_SECOND_SAMPLE_TIME", "README_SECOND_SAMPLE_TIME",
    "COST_CHOICES", "JNEES", "CNEES", "JNIS", "CNIS", "SHARD_TRIALS", "SHARD_CANDIDATES", "LIB_PATH",
]
