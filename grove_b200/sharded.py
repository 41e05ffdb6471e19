"""Multi-GPU cycle: gang rows sharded over ranks, node table and gang state replicated.

One collective per round: every rank evaluates its own gangs into an exchange
buffer (up to K alternative placements per gang; zero for gangs it does not own), the buffer is
all-reduced with SUM, and every rank then resolves the conflicts and commits identically.  All ranks
end with identical node tables, statuses and placements -- identical to the single-GPU cycle and to
the oracle.

`run_sharded_cycle(stepper, dist)` drives any object with the stepping interface of
`grove_b200.engine.PlacementEngine` (cycle_begin / round_eval / round_resolve / cycle_end) whose
round_eval returns a buffer `as_tensor` can view: the CUDA engine returns a device pointer (reduced
with NCCL), the CPU test stepper a numpy array (gloo).
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
    """One scheduling cycle over all ranks of `dist` (torch.distributed, already initialised).

    With the CUDA engine the all-reduce is enqueued on the engine's own stream (wrapped as a
    torch.cuda.ExternalStream), so eval -> all-reduce -> resolve are ordered by that stream and the host
    never waits inside a round except for the few counters the engine reads to size its grids."""
    import torch

    ext = None
    if hasattr(stepper, "stream") and hasattr(stepper, "set_stream_ordered"):
        ext = torch.cuda.ExternalStream(stepper.stream())
        stepper.set_stream_ordered(True)
    view, view_key = None, None
    try:
        stepper.cycle_begin()
        while True:
            buf, n, go = stepper.round_eval()
            if not go:
                break
            if n:
                if view is None or view_key != (buf if not isinstance(buf, np.ndarray) else id(buf), n):
                    view, view_key = as_tensor(buf, n), (buf if not isinstance(buf, np.ndarray) else id(buf), n)
                if ext is not None:
                    with torch.cuda.stream(ext):
                        dist.all_reduce(view, op=dist.ReduceOp.SUM)
                else:
                    dist.all_reduce(view, op=dist.ReduceOp.SUM)
            stepper.round_resolve()
        return stepper.cycle_end()
    finally:
        if ext is not None:
            stepper.set_stream_ordered(False)
