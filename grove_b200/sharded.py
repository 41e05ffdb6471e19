"""Multi-GPU score pass: node-range shards + ONE all-reduce of per-shard feasibility (BASELINE.json north_star).

What outgrows a GPU in this path is the Q x N score matrix, not the node table (32 B a node): so every rank loads the whole node
table and the whole submission, builds K1 + K2 only for ITS node range (the topology-sorted table cut at top-level domain
boundaries), and contributes one int32[G + Q] vector -- per gang the candidate domains of its Required level inside the shard
that could hold it, per clique the pods that fit on the shard's nodes.  Their SUM over the ranks is the one collective on the
path (`dist.all_reduce`, NCCL over NVLink on the GPU box, gloo in the CPU tests); cluster-wide feasibility is read off the sum.
The admission itself (K3) is one sequential dependency chain and does not shard.

`summary_fn(rank) -> int32[G + Q] tensor on the collective's device` abstracts who computes the shard summary: the engine on a
GPU (engine_summary below), the numpy restatement in oracle/ for the CPU tests.
"""
from __future__ import annotations

import numpy as np

from . import tables as T


def infeasible_from_sum(total: np.ndarray, gangs: np.ndarray, cliques: np.ndarray) -> np.ndarray:
    """bool[G]: gangs that cannot be admitted from the snapshot, given the all-reduced int32[G + Q] vector: no domain of the
    Required level holds every clique's MinReplicas, or (no Required level) some clique's MinReplicas exceed what fits in the
    whole cluster.  A necessary condition only: the others may still lose to gangs that rank before them."""
    G = len(gangs)
    feas, cap = total[:G], total[G:]
    out = np.zeros(G, dtype=bool)
    short = cap < cliques["min_replicas"].astype(np.int64)          # per clique: less capacity cluster-wide than it needs
    any_short = np.zeros(G, dtype=bool)
    for g in range(G):
        a = int(gangs["clique_off"][g])
        any_short[g] = short[a: a + int(gangs["n_cliques"][g])].any()
    has_level = gangs["level"] != T.LEVEL_NONE
    out[has_level] = feas[has_level] == 0
    out |= any_short
    return out


def sharded_score_pass(dist, world: int, summary_fn, rank: int):
    """-> (total int32[G + Q] numpy, seconds spent in the collective).  One all-reduce(SUM)."""
    import time
    import torch

    t = summary_fn(rank)
    t0 = time.perf_counter()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if t.is_cuda:
            torch.cuda.synchronize(t.device)
    return t.cpu().numpy(), time.perf_counter() - t0


def engine_summary(engine, device):
    """summary_fn for a PlacementEngine that has just run its score pass: the shard summary lands in a torch tensor on
    `device` (the buffer NCCL reduces in place)"""
    import torch

    buf = {}

    def fn(_rank):
        n = engine.G + engine.Q
        if buf.get("n") != n:   # one buffer per submission size: the pass runs every cycle
            buf["t"], buf["n"] = torch.empty(n, dtype=torch.int32, device=device), n
        engine.shard_summary_into(buf["t"].data_ptr(), n)   # every word of the vector is written
        return buf["t"]
    return fn
