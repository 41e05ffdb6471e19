"""world_size-2/3 run of the multi-GPU orchestration (replicas only: grove_b200/replicas.py) over gloo on CPU, the oracle
standing in for the engine: every rank schedules its own cluster, nothing but the bookkeeping reduction crosses ranks, and
the aggregate is the sum of what each cluster's own sequential pass admits."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from grove_b200 import synth
    from grove_b200.replicas import run_replicas
    from oracle import oracle_py as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def make_cycle(r):
        cfg = synth.config_c4(n=2520, g=300, seed=synth.SEED_BASE + 4 + 1000 * r)
        g, c, s = cfg["tables"]

        def cycle():
            res = O.run_cycle(cfg["nodes"], cfg["n_levels"], g, c, s)
            return res["stats"]["gangs_admitted"], res["stats"]["gangs_rejected"]
        return cycle

    out = run_replicas(dist, rank, world, make_cycle, steps=2, warmup=1)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([out["admitted"], out["rejected"], out["seconds"]] + [x for p in out["per_rank"] for x in p]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_replicas_aggregate_over_gloo(oracle, tmp_path, world):
    import torch.multiprocessing as mp
    from grove_b200 import synth
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    for r in rows[1:]:
        assert np.array_equal(r[:2], rows[0][:2]) and np.array_equal(r[3:], rows[0][3:])   # every rank holds the same aggregate
    expect_adm = expect_rej = 0
    per = rows[0][3:].reshape(world, 3)
    for r in range(world):
        cfg = synth.config_c4(n=2520, g=300, seed=synth.SEED_BASE + 4 + 1000 * r)
        res = oracle.run_cycle(cfg["nodes"], cfg["n_levels"], *cfg["tables"])
        assert per[r][0] == res["stats"]["gangs_admitted"] and per[r][1] == res["stats"]["gangs_rejected"]
        expect_adm += res["stats"]["gangs_admitted"]; expect_rej += res["stats"]["gangs_rejected"]
    assert rows[0][0] == expect_adm and rows[0][1] == expect_rej and expect_adm > 0
    assert abs(rows[0][2] - per[:, 2].max()) < 1e-9   # the job's time is the slowest rank's
    assert len({tuple(p[:2]) for p in per.tolist()}) > 1   # different clusters: the ranks did different work
