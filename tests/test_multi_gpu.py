"""-m gpu, needs >= 2 GPUs (skipped otherwise): the multi-GPU path as bench.py --gpus N drives it -- one process per GPU
over NCCL, every rank scheduling its own cluster through the C ABI (replicas only, grove_b200/replicas.py) -- with every
rank's outputs compared against the sequential oracle."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from grove_b200 import synth
    from grove_b200.engine import PlacementEngine
    from grove_b200.replicas import run_replicas
    from oracle import oracle_py as O
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = synth.config_c4(n=5040, g=1000, seed=synth.SEED_BASE + 4 + 1000 * rank)
    g, c, s = cfg["tables"]
    eng = PlacementEngine(cfg["n_levels"], device=rank)
    eng.submit_gangs(g, c, s)

    def make_cycle(_r):
        def cycle():
            eng.load_nodes(cfg["nodes"])
            st = eng.run_cycle()
            return st["gangs_admitted"], st["gangs_rejected"]
        return cycle

    out = run_replicas(dist, rank, world, make_cycle, steps=2, warmup=1, device=torch.device("cuda", rank))
    ref = O.run_cycle(cfg["nodes"], cfg["n_levels"], g, c, s, threads=4)
    ok = (np.array_equal(eng.placements(), ref["placements"]) and np.array_equal(eng.gang_status(), ref["status"]) and
          np.array_equal(eng.scope_domains(), ref["scope_status"]) and np.array_equal(eng.nodes(), ref["nodes_after"]))
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([int(ok), out["admitted"], ref["stats"]["gangs_admitted"]]))
    eng.close()
    dist.destroy_process_group()


def test_replicas_over_nccl(built_lib, oracle, tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    assert all(r[0] == 1 for r in rows)                       # every rank bit-identical to the oracle on its own cluster
    assert rows[0][1] == sum(r[2] for r in rows) > 0          # the aggregate is the sum over the clusters
