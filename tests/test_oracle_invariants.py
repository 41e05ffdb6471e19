"""Invariants every cycle result must satisfy whatever the snapshot, checked on the oracle over the seeded random
cases of tests/test_random_parity_gpu.py (ragged / non-tree topologies, cordons, selector classes, zero requests,
MinReplicas 0, surplus, nested Required and Preferred levels, priorities, base-gang chains).  The CUDA path is held
bit-identical to the oracle on the same cases (-m gpu), so these carry over to it."""
import numpy as np
import pytest

from grove_b200 import tables as T
from test_random_parity_gpu import random_case


def check(nodes, L, tabs, r):
    g, c, s = tabs
    st, pl, after = r["status"], r["placements"], r["nodes_after"]
    # ---- bookkeeping: placement ranges tile the output, counts agree
    off = 0
    for gi in range(len(g)):
        if st["state"][gi] == T.GANG_ADMITTED:
            assert st["placement_off"][gi] == off
            off += int(st["n_pods"][gi])
        else:
            assert st["n_pods"][gi] == 0
    assert off == len(pl) == r["stats"]["pods_bound"]
    # ---- resources: what left the node table is exactly what the bound pods request, and nothing goes negative
    used = {k: np.zeros(len(nodes), dtype=np.int64) for k in ("cpu", "mem", "gpu", "pods")}
    q = c[pl["clique"]]
    np.add.at(used["cpu"], pl["node"], q["req_cpu_milli"]); np.add.at(used["mem"], pl["node"], q["req_mem_mib"])
    np.add.at(used["gpu"], pl["node"], q["req_gpu"]); np.add.at(used["pods"], pl["node"], 1)
    for k, f in (("cpu", "free_cpu_milli"), ("mem", "free_mem_mib"), ("gpu", "free_gpu"), ("pods", "free_pods")):
        assert np.array_equal(nodes[f].astype(np.int64) - used[k], after[f].astype(np.int64)), k
        assert (nodes[f].astype(np.int64) >= used[k]).all(), k
    assert np.array_equal(nodes["flags"], after["flags"]) and np.array_equal(nodes["dom"], after["dom"])
    # ---- per pod: schedulable node, accepted selector class
    cls = (nodes["flags"][pl["node"]] >> T.NODE_CLASS_SHIFT) & 0xF
    assert (nodes["flags"][pl["node"]] & T.NODE_SCHEDULABLE).all()
    assert ((q["class_mask"] >> cls) & 1).all()
    # ---- per gang
    rank = np.empty(len(g), dtype=np.int64)
    rank[np.lexsort((np.arange(len(g)), -g["priority"].astype(np.int64)))] = np.arange(len(g))
    for gi in range(len(g)):
        state = st["state"][gi]
        gg = g[gi]
        if gg["flags"] & T.GANG_GATED:
            assert state == T.GANG_GATED_SKIP
            continue
        assert state in (T.GANG_ADMITTED, T.GANG_REJECTED, T.GANG_BASE_REJECTED)     # a full cycle decides every gang
        if gg["base_gang"] != T.NONE_U32:
            bstate = st["state"][gg["base_gang"]]
            if state == T.GANG_ADMITTED:
                assert bstate == T.GANG_ADMITTED and rank[gg["base_gang"]] < rank[gi]   # the base gang had its turn first
            if bstate != T.GANG_ADMITTED:
                assert state == T.GANG_BASE_REJECTED
        if state != T.GANG_ADMITTED:
            continue
        mine = pl[st["placement_off"][gi]: st["placement_off"][gi] + st["n_pods"][gi]]
        rel = mine["clique"].astype(np.int64) - int(gg["clique_off"])
        assert ((rel >= 0) & (rel < gg["n_cliques"])).all()
        assert 0 < st["score_den"][gi] and st["score_num"][gi] <= st["score_den"][gi]  # PlacementScore in [0, 1]
        asked = [(gg["level"], gg["preferred"])] + [(x["level"], x["preferred1"] - 1 if x["preferred1"] else T.LEVEL_NONE) for x in s[gg["scope_off"]: gg["scope_off"] + gg["n_scopes"]]]
        asked += [(x["level"], (x["scope"] >> 5) - 1 if x["scope"] >> 5 else T.LEVEL_NONE) for x in c[gg["clique_off"]: gg["clique_off"] + gg["n_cliques"]]]
        if all(p == T.LEVEL_NONE for _, p in asked):     # nothing Preferred: every Required level held => 1.0
            assert st["score_num"][gi] == st["score_den"][gi]
        dom = lambda sel, lvl: set(nodes["dom"][mine["node"][sel], lvl].tolist())
        for ci in range(gg["n_cliques"]):
            cq = c[gg["clique_off"] + ci]
            n = int((rel == ci).sum())
            assert cq["min_replicas"] <= n <= cq["replicas"]                          # all-or-nothing minimum, bounded surplus
            if cq["level"] != T.LEVEL_NONE and n:
                d = dom(rel == ci, cq["level"])
                assert len(d) == 1 and T.DOM_ABSENT not in d
        for si in range(gg["n_scopes"]):
            sc = s[gg["scope_off"] + si]
            sel = (rel >= sc["first_clique"]) & (rel < sc["first_clique"] + sc["n_cliques"])
            if sc["level"] != T.LEVEL_NONE and sel.any():
                d = dom(sel, sc["level"])
                assert len(d) == 1 and T.DOM_ABSENT not in d
        if gg["level"] != T.LEVEL_NONE and len(mine):
            d = dom(slice(None), gg["level"])
            assert len(d) == 1 and T.DOM_ABSENT not in d


@pytest.mark.parametrize("block", range(6))
def test_cycle_invariants_on_random_snapshots(oracle, block):
    for seed in range(block * 30, block * 30 + 30):
        for pref in (False, True):
            nodes, L, tabs = random_case(seed + (3000 if pref else 0), pref=pref)
            r = oracle.run_cycle(nodes, L, *tabs, threads=2)
            check(nodes, L, tabs, r)


def test_cycle_invariants_on_bigger_cases(oracle):
    for seed in (2000, 2001, 4000, 4001):
        nodes, L, tabs = random_case(seed, big=True, pref=seed >= 4000)
        r = oracle.run_cycle(nodes, L, *tabs, threads=8)
        check(nodes, L, tabs, r)


def test_results_do_not_depend_on_threads(oracle):
    for seed in range(40, 60):
        nodes, L, tabs = random_case(seed, pref=seed % 2 == 0)
        a = oracle.run_cycle(nodes, L, *tabs, threads=1)
        b = oracle.run_cycle(nodes, L, *tabs, threads=5)
        assert np.array_equal(a["placements"], b["placements"]) and np.array_equal(a["status"], b["status"])
