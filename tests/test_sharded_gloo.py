"""world_size-2 (and 3) gloo runs of the multi-rank cycle protocol on CPU: grove_b200.sharded drives
the CPU stepper of oracle/ through the same begin/eval/commit/apply/gather/end sequence and the same
reductions the CUDA engine uses with NCCL; every rank must end bit-identical to the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from grove_b200 import synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _config(cfg_name, cfg_kw):
    if cfg_name == "random":   # a seeded random snapshot of tests/test_random_parity_gpu.py (Preferred levels, base chains, ragged labels)
        from test_random_parity_gpu import random_case
        nodes, L, tabs = random_case(**cfg_kw)
        return dict(nodes=nodes, n_levels=L, tables=tabs)
    return getattr(synth, cfg_name)(**cfg_kw)


def _worker(rank, world, port, cfg_name, cfg_kw, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from grove_b200.sharded import run_sharded_cycle
        from oracle import oracle_py as O
        cfg = _config(cfg_name, cfg_kw)
        g, c, s = cfg["tables"]
        st = O.OracleStepper(cfg["nodes"], cfg["n_levels"], g, c, s, rank, world)
        stats = run_sharded_cycle(st, dist)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), rounds=stats["rounds"], **st.result)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cfg_name,cfg_kw", [
    (2, "config_c3", dict(n=756, g=120)),
    (2, "config_c4", dict(n=2520, g=300)),
    (3, "config_c2", dict(n=300, g=60)),
    (2, "random", dict(seed=3005, pref=True)),
    (3, "random", dict(seed=4001, big=True, pref=True)),
])
def test_sharded_protocol_matches_unsharded(oracle, tmp_path, world, cfg_name, cfg_kw):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, cfg_name, cfg_kw, str(tmp_path)), nprocs=world, join=True)
    cfg = _config(cfg_name, cfg_kw)
    g, c, s = cfg["tables"]
    ref = oracle.run_cycle(cfg["nodes"], cfg["n_levels"], g, c, s)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert int(z["rounds"]) == ref["stats"]["rounds"]
        assert np.array_equal(z["placements"], ref["placements"])
        assert np.array_equal(z["status"], ref["status"])
        assert np.array_equal(z["nodes_after"], ref["nodes_after"])
