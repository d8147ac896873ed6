"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only to hand the NCCL
unique id around; the data path is ONE ncclAllReduce of the [C][D][4] cost sums, issued inside
libkfbo.so on the evaluator's stream (kfbo_comm_init / kfbo_eval_batch).

Trials and candidates are independent (trial_node.cpp:54-67 already splits trials over threads), so
the path shards with no other exchange:
  SHARD_TRIALS      rank r runs trials [r*B/G, (r+1)*B/G); Philox counters use the GLOBAL trial index,
                    so the draws do not depend on G.
  SHARD_CANDIDATES  rank r runs a contiguous block of candidates; the other slots contribute zeros.
"""
from __future__ import annotations

import numpy as np

from .evaluator import CostEvaluator, nccl_unique_id
from ._lib import SHARD_CANDIDATES, SHARD_TRIALS  # noqa: F401


def init_comm(ev: CostEvaluator, shard_mode: int = SHARD_TRIALS) -> None:
    """Collective over the default torch.distributed group: rank 0 creates the NCCL id, everyone
    joins the evaluator's communicator."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ev.comm_init(box[0], rank, world, shard_mode)


def allreduce_sums_host(local_sums: np.ndarray) -> np.ndarray:
    """Sum of per-rank partial cost sums through torch.distributed on HOST tensors (gloo).  Not on
    the GPU data path (that is the in-library ncclAllReduce); used where the partial sums already
    live on the host, and by the world_size>1 CPU tests of the sharding logic."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local_sums, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
