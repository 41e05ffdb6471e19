"""Multi-GPU cycle: gang rows sharded over ranks, node table and gang state replicated.

The only data exchanged are three small int32 buffers per cycle phase:
claims (all-reduce MIN, N words per round), commit deltas + new gang states (all-reduce SUM,
4N + G words per round) and, once per cycle, the admitted gangs' placement entries (all-reduce SUM).
Every rank applies the same reduced data, so all ranks end with identical node tables, statuses and
placements -- identical to the single-GPU cycle and to the oracle.

`run_sharded_cycle(stepper, dist)` drives any object with the stepping interface of
`grove_b200.engine.PlacementEngine` (cycle_begin / round_eval / round_commit / round_apply /
cycle_gather / cycle_end) whose step methods return buffers `as_tensor` can view: the CUDA engine
returns device pointers (reduced with NCCL), the CPU test stepper returns numpy arrays (gloo).
"""
from __future__ import annotations

import numpy as np


class _CudaWords:
    """zero-copy view of an engine-owned int32 device buffer for torch (CUDA array interface)"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}


def as_tensor(buf, n: int):
    import torch
    if isinstance(buf, np.ndarray):
        return torch.from_numpy(buf[:n])
    if n == 0:
        return torch.empty(0, dtype=torch.int32, device="cuda")
    return torch.as_tensor(_CudaWords(int(buf), n), device="cuda")


def run_sharded_cycle(stepper, dist) -> dict:
    """One scheduling cycle over all ranks of `dist` (torch.distributed, already initialised)."""
    import torch

    cuda = None

    def reduce(buf, n, op):
        nonlocal cuda
        if n == 0:
            return
        t = as_tensor(buf, n)
        dist.all_reduce(t, op=op)
        if t.is_cuda:  # the engine reads the buffer from its own stream next
            cuda = True
            torch.cuda.current_stream().synchronize()

    stepper.cycle_begin()
    while True:
        claim, n, go = stepper.round_eval()
        if not go:
            break
        reduce(claim, n, dist.ReduceOp.MIN)
        delta, m = stepper.round_commit()
        reduce(delta, m, dist.ReduceOp.SUM)
        stepper.round_apply()
    fin, k = stepper.cycle_gather()
    reduce(fin, k, dist.ReduceOp.SUM)
    return stepper.cycle_end()
