"""The multi-GPU score pass (grove_b200/sharded.py; BASELINE.json north_star: node-range shards + one all-reduce of per-shard
feasibility) at world size 2 and 3 over gloo on CPU, the numpy restatement of the shard summary (oracle/oracle_py.py) standing in
for the engine: the all-reduced vector equals the summary of the whole table, and what it calls infeasible the sequential pass
never admits."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case():
    from grove_b200 import synth, tables as T
    cfg = synth.config_c4(n=7560, g=300, max_used_pct=97)   # three zones   # a nearly full cluster: some gangs fit nowhere
    g, c, s = (a.copy() for a in cfg["tables"])
    g["level"][::7] = T.LEVEL_NONE                          # some gangs without a Required level: judged by the capacity counts
    return cfg["nodes"], cfg["n_levels"], g, c, s


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from grove_b200.sharded import sharded_score_pass
    from oracle import oracle_py as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nodes, L, g, c, s = _case()
    perm, dom, _, _ = O.topology(nodes, L)
    lo, hi = O.shard_cut(len(nodes), dom[:, 0], rank, world), O.shard_cut(len(nodes), dom[:, 0], rank + 1, world)

    def summary(_r):
        return torch.from_numpy(O.shard_summary(nodes, L, g, c, s, lo, hi))

    total, _ = sharded_score_pass(dist, world, summary, rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[lo, hi], total]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_all_reduced_shard_summaries(oracle, tmp_path, world):
    import torch.multiprocessing as mp
    from grove_b200 import tables as T
    from grove_b200.sharded import infeasible_from_sum
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    nodes, L, g, c, s = _case()
    n = len(nodes)
    # the shards tile the table, cut at top-level domain boundaries
    assert rows[0][0] == 0 and rows[-1][1] == n and all(rows[r][1] == rows[r + 1][0] for r in range(world - 1))
    perm, dom, _, _ = oracle.topology(nodes, L)
    for r in range(1, world):
        cut = int(rows[r][0])
        assert 0 < cut < n and dom[cut, 0] != dom[cut - 1, 0]
    whole = oracle.shard_summary(nodes, L, g, c, s, 0, n)
    for r in rows:
        assert np.array_equal(r[2:], whole)            # every rank holds the sum, and it is the whole table's summary
    bad = infeasible_from_sum(whole, g, c)
    ref = oracle.run_cycle(nodes, L, g, c, s)
    assert bad.any() and not bad.all()
    assert (ref["status"]["state"][bad] != T.GANG_ADMITTED).all()   # a necessary condition: what it rules out is never admitted


def test_infeasible_from_sum_reads_the_reduced_vector():
    """gang 0: Required level, no feasible domain anywhere -> infeasible; gang 1: Required level, one feasible domain -> not;
    gang 2: no Required level, a clique whose MinReplicas exceed the cluster's capacity -> infeasible; gang 3: neither"""
    from grove_b200 import tables as T
    from grove_b200.sharded import infeasible_from_sum
    b = T.GangTableBuilder()
    for lvl in (1, 1, None, None):
        b.add_gang([(None, [dict(cpu=1000, mem=1, gpu=0, min=4), dict(cpu=1000, mem=1, gpu=0, min=0)])], level=lvl)
    g, c, s = b.build()
    total = np.array([0, 1, 0, 0,   9, 0,  9, 0,  3, 0,  4, 0], dtype=np.int64)   # [G feasible-domain counts | Q capacity counts]
    assert infeasible_from_sum(total, g, c).tolist() == [True, False, True, False]


def test_shard_cut_tiles_the_table_at_top_level_boundaries(oracle):
    dom0 = np.repeat(np.arange(5, dtype=np.uint32), [7, 3, 10, 1, 9])          # five zones of uneven size, 30 nodes
    dom0 = np.concatenate([dom0, np.full(2, 0xFFFFFFFF, dtype=np.uint32)])     # two nodes without the label, sorted last
    n = len(dom0)
    for world in (1, 2, 3, 4, 8):
        cuts = [oracle.shard_cut(n, dom0, r, world) for r in range(world + 1)]
        assert cuts[0] == 0 and cuts[-1] == n and cuts == sorted(cuts)
        for cpos in cuts[1:-1]:
            assert cpos == n or (dom0[cpos] != dom0[cpos - 1] and dom0[cpos] != 0xFFFFFFFF)
