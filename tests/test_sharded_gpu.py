"""The multi-GPU score pass through the C ABI (grove_run_score_pass / grove_shard_summary_device, world > 1 handles):
  * one GPU: the shards of a world-2 and a world-3 job, one handle after the other -- each shard's K2 columns equal the unsharded
    score matrix's, each summary equals the numpy restatement over the same node range, and the summaries add up to the whole
    table's;
  * >= 2 GPUs (skipped otherwise): one process per GPU, the ONE all-reduce over NCCL, the sum equal to the whole table's summary
    on every rank."""
import os
import socket

import numpy as np
import pytest

from grove_b200 import synth, tables as T

pytestmark = pytest.mark.gpu


def _case():
    cfg = synth.config_c4(n=7560, g=300, max_used_pct=97)
    g, c, s = (a.copy() for a in cfg["tables"])
    g["level"][::7] = T.LEVEL_NONE
    return cfg["nodes"], cfg["n_levels"], g, c, s


@pytest.mark.parametrize("world", [1, 2, 3])
def test_shards_on_one_gpu(built_lib, oracle, world):
    import torch
    from grove_b200.engine import PlacementEngine
    from grove_b200.sharded import engine_summary, infeasible_from_sum
    nodes, L, g, c, s = _case()
    n = len(nodes)
    perm, dom, _, _ = oracle.topology(nodes, L)
    with PlacementEngine(L) as full:
        full.load_nodes(nodes); full.submit_gangs(g, c, s); full.run_cycle()
        rows = {q: full.debug_score_row(q) for q in (0, 5, len(c) // 2, len(c) - 1)}   # caller node order
        state = full.gang_status()["state"].copy()
    total = np.zeros(len(g) + len(c), dtype=np.int64)
    covered = np.zeros(n, dtype=np.int32)
    for r in range(world):
        with PlacementEngine(L, rank=r, world=world) as e:
            e.load_nodes(nodes); e.submit_gangs(g, c, s)
            ms = e.run_score_pass()
            lo, hi = e.shard_range()
            assert (lo, hi) == (oracle.shard_cut(n, dom[:, 0], r, world), oracle.shard_cut(n, dom[:, 0], r + 1, world)) and ms > 0
            covered[lo:hi] += 1
            t = engine_summary(e, torch.device("cuda", 0))(r)
            mine = t.cpu().numpy()
            assert np.array_equal(mine, oracle.shard_summary(nodes, L, g, c, s, lo, hi))
            total += mine
            for q, want in rows.items():   # this shard's columns of the score row, zeros elsewhere
                got = e.debug_score_row(q)[perm]             # -> topology-sorted order
                w = want[perm]
                assert np.array_equal(got[lo:hi], w[lo:hi])
                a, b = (lo // 16) * 16, min(n, -(-hi // 16) * 16)   # the shard is built in whole 16-node chunks
                assert not got[:a].any() and not got[b:].any()
            # the handle is ready for an ordinary cycle afterwards
            e.run_cycle()
            assert np.array_equal(e.gang_status()["state"], state)
    assert (covered == 1).all()
    assert np.array_equal(total, oracle.shard_summary(nodes, L, g, c, s, 0, n))
    bad = infeasible_from_sum(total, g, c)
    assert bad.any() and (state[bad] != T.GANG_ADMITTED).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from grove_b200.engine import PlacementEngine
    from grove_b200.sharded import engine_summary, sharded_score_pass
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    nodes, L, g, c, s = _case()
    with PlacementEngine(L, device=rank, rank=rank, world=world) as e:
        e.load_nodes(nodes); e.submit_gangs(g, c, s)
        e.run_score_pass()
        total, _ = sharded_score_pass(dist, world, engine_summary(e, dev), rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), total)
    dist.destroy_process_group()


def test_all_reduce_over_nccl(built_lib, oracle, tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 3)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    nodes, L, g, c, s = _case()
    whole = oracle.shard_summary(nodes, L, g, c, s, 0, len(nodes))
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), whole)
