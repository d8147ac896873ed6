"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only to hand the NCCL
unique id and the CUDA IPC handles of the exchange buffers around.  The data path is ONE exchange
of the [C][D][4] cost sums inside libkfbo.so: over peer memory (NVLink / NVSwitch stores from the
tail of the rollout kernel, kfbo_peer_export / kfbo_peer_attach) when the ranks can map each
other's buffers, else one ncclAllReduce on the evaluator's stream (kfbo_comm_init).

Trials and candidates are independent (trial_node.cpp:54-67 already splits trials over threads), so
the path shards with no other exchange:
  SHARD_TRIALS      rank r runs trials [r*B/G, (r+1)*B/G); Philox counters use the GLOBAL trial index,
                    so the draws do not depend on G.
  SHARD_CANDIDATES  rank r runs a contiguous block of candidates; the other slots contribute zeros.
"""
from __future__ import annotations

import numpy as np

from .evaluator import CostEvaluator, KfboError, nccl_unique_id
from ._lib import MAX_PEERS, OPT_REDUCE, REDUCE_NCCL, SHARD_CANDIDATES, SHARD_TRIALS  # noqa: F401


def init_comm(ev: CostEvaluator, shard_mode: int = SHARD_TRIALS, peer: bool = True) -> bool:
    """Collective over the default torch.distributed group: rank 0 creates the NCCL id, everyone
    joins the evaluator's communicator; then (peer=True, at most MAX_PEERS ranks) the ranks exchange
    the IPC handles of their exchange buffers and attach to each other.  Returns True when the
    peer-memory exchange is in place on EVERY rank; otherwise every rank stays on the NCCL call."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ev.comm_init(box[0], rank, world, shard_mode)
    if not peer or world < 2 or world > MAX_PEERS:
        return False
    return attach_peers(ev, world)


def attach_peers(ev, world: int, all_gather=None) -> bool:
    """Export / all-gather / attach, then agree: the exchange is used only if it came up on all ranks
    (a rank that cannot map a peer's buffer -- no P2P, IPC not permitted -- turns it off for everyone).
    all_gather(obj) -> list of every rank's obj; default: torch.distributed.all_gather_object."""
    if all_gather is None:
        import torch.distributed as dist

        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
    try:
        mine = ev.peer_export()
    except KfboError:
        mine = None
    handles = all_gather(mine)
    ok = all(h is not None for h in handles)
    if ok:
        try:
            ev.peer_attach(handles)
        except KfboError:
            ok = False
    ok = all(all_gather(ok))
    if not ok:
        ev.set_option(OPT_REDUCE, REDUCE_NCCL)
    return ok


def allreduce_sums_host(local_sums: np.ndarray) -> np.ndarray:
    """Sum of per-rank partial cost sums through torch.distributed on HOST tensors (gloo).  Not on
    the GPU data path (that is the in-library ncclAllReduce); used where the partial sums already
    live on the host, and by the world_size>1 CPU tests of the sharding logic."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local_sums, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
