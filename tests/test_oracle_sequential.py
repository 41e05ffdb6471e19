"""The cycle is the SEQUENTIAL priority-ordered pass (PodGangSpec.PriorityClassName, podgang.go:62-64; SURVEY.md
section 7 step 1): gangs one at a time in (priority desc, submission index asc) order, each evaluated against what the
earlier ones left.  These tests pin that contract on the oracle with scenarios that an optimistic "everybody proposes,
winners commit" resolution gets wrong (round-1 VERDICT "What's weak" #1), and with an independent
re-derivation: a pure-Python driver that feeds the oracle ONE gang per call in rank order must reproduce the cycle."""
import numpy as np
import pytest

from grove_b200 import synth, tables as T
from test_random_parity_gpu import random_case


def one_pod_nodes(n):
    """n nodes that hold exactly one of our pods each, flat topology"""
    nodes = synth.kwok_nodes(n, [1], cpu_milli=1000, mem_mib=1024, gpu=0, pods=110)
    return nodes


def pod(n):
    return dict(cpu=1000, min=n)


def states(r):
    return r["status"]["state"].tolist()


def test_three_gang_inversion_of_the_optimistic_protocol(oracle):
    """6 one-pod nodes; H1 (prio 10, 3 pods, anchor 0), H2 (prio 9, 3 pods, anchor 0), L (prio 0, 3 pods, anchor 3).
    An optimistic resolution admits H1 and L in the same sub-round and rejects H2; priority order admits H1 and H2."""
    b = T.GangTableBuilder()
    b.add_gang([(None, [pod(3)])], priority=10, anchor=0)
    b.add_gang([(None, [pod(3)])], priority=9, anchor=0)
    b.add_gang([(None, [pod(3)])], priority=0, anchor=3)
    r = oracle.run_cycle(one_pod_nodes(6), 1, *b.build())
    assert states(r) == [T.GANG_ADMITTED, T.GANG_ADMITTED, T.GANG_REJECTED]
    pl = r["placements"]
    assert sorted(pl["node"][:3].tolist()) == [0, 1, 2] and sorted(pl["node"][3:6].tolist()) == [3, 4, 5]


def test_two_node_inversion(oracle):
    """2 nodes, three one-pod gangs that each fill a node: M (prio 100, anchor n0), H (prio 10, anchor n0),
    L (prio 0, anchor n1).  H must get n1; L is the one left out."""
    b = T.GangTableBuilder()
    b.add_gang([(None, [pod(1)])], priority=100, anchor=0)
    b.add_gang([(None, [pod(1)])], priority=10, anchor=0)
    b.add_gang([(None, [pod(1)])], priority=0, anchor=1)
    r = oracle.run_cycle(one_pod_nodes(2), 1, *b.build())
    assert states(r) == [T.GANG_ADMITTED, T.GANG_ADMITTED, T.GANG_REJECTED]
    assert r["placements"]["node"].tolist() == [0, 1]


def test_submission_order_breaks_priority_ties(oracle):
    b = T.GangTableBuilder()
    for _ in range(3):
        b.add_gang([(None, [pod(2)])], priority=5, anchor=0)
    r = oracle.run_cycle(one_pod_nodes(4), 1, *b.build())
    assert states(r) == [T.GANG_ADMITTED, T.GANG_ADMITTED, T.GANG_REJECTED]


def test_scaled_gang_waits_for_a_base_gang_that_ranks_later(oracle):
    """a scaled gang is considered at its own turn; a base gang that has not been admitted by then (here: lower
    priority, so it comes later) leaves it BASE_REJECTED for this cycle -- its pods keep their scheduling gate until
    the base gang is scheduled (pod/syncflow.go:319-358)"""
    b = T.GangTableBuilder()
    base = b.add_gang([(None, [pod(1)])], priority=0)
    b.add_gang([(None, [pod(1)])], priority=5, base=base)
    late = b.add_gang([(None, [pod(1)])], priority=0, base=base)
    r = oracle.run_cycle(one_pod_nodes(4), 1, *b.build())
    assert states(r) == [T.GANG_ADMITTED, T.GANG_BASE_REJECTED, T.GANG_ADMITTED]
    assert r["status"]["n_pods"][late] == 1


def _rank_order(g):
    return np.lexsort((np.arange(len(g)), -g["priority"].astype(np.int64)))


def _one_by_one(oracle, nodes, L, tabs):
    """the cycle re-derived outside the oracle's own loop: one single-gang cycle per gang, in rank order, each on the
    node table the previous call returned"""
    g, c, s = tabs
    state = np.zeros(len(g), dtype=np.uint8)
    placed = {}
    cur = nodes.copy()
    perm = oracle.topology(nodes, L)[0]
    inv = np.empty(len(nodes), dtype=np.int64); inv[perm] = np.arange(len(nodes))
    for gi in _rank_order(g):
        gg = g[gi]
        if gg["flags"] & T.GANG_GATED:
            state[gi] = T.GANG_GATED_SKIP
            continue
        if gg["base_gang"] != T.NONE_U32 and state[gg["base_gang"]] != T.GANG_ADMITTED:
            state[gi] = T.GANG_BASE_REJECTED
            continue
        one = g[gi: gi + 1].copy()
        cs = c[gg["clique_off"]: gg["clique_off"] + gg["n_cliques"]].copy()
        ss = s[gg["scope_off"]: gg["scope_off"] + gg["n_scopes"]].copy()
        one["clique_off"] = 0; one["scope_off"] = 0; one["base_gang"] = T.NONE_U32
        # the default anchor is derived from the gang's index in the submission: pin it to what the full table gets
        if one["anchor_node"][0] == T.NONE_U32:
            one["anchor_node"] = perm[_fmix32(int(gi)) % len(nodes)]
        r = oracle.run_cycle(cur, L, one, cs, ss)
        state[gi] = r["status"]["state"][0]
        if state[gi] == T.GANG_ADMITTED:
            pl = r["placements"].copy(); pl["clique"] += gg["clique_off"]
            placed[int(gi)] = pl
            cur = r["nodes_after"]
    return state, placed, cur


def _fmix32(x):
    m = 0xFFFFFFFF
    x = (x * 0x9E3779B1 + 0x7F4A7C15) & m
    x ^= x >> 16; x = (x * 0x85EBCA6B) & m; x ^= x >> 13; x = (x * 0xC2B2AE35) & m; x ^= x >> 16
    return x


@pytest.mark.parametrize("block", range(4))
def test_cycle_equals_one_gang_at_a_time(oracle, block):
    for seed in range(block * 25, block * 25 + 25):
        for pref in (False, True):
            nodes, L, tabs = random_case(seed + (3000 if pref else 0), pref=pref)
            r = oracle.run_cycle(nodes, L, *tabs)
            state, placed, after = _one_by_one(oracle, nodes, L, tabs)
            assert np.array_equal(r["status"]["state"], state), seed
            assert np.array_equal(r["nodes_after"], after), seed
            for gi, pl in placed.items():
                o, k = int(r["status"]["placement_off"][gi]), int(r["status"]["n_pods"][gi])
                assert np.array_equal(r["placements"][o: o + k], pl), (seed, gi)


def test_no_priority_inversion_on_random_snapshots(oracle):
    """ADVICE #1's invariant: no REJECTED gang would have been feasible on the state it met at its turn -- checked
    as: re-running the cycle without any one lower-ranked ADMITTED gang never changes the fate of a higher-ranked gang."""
    for seed in range(200, 230):
        nodes, L, (g, c, s) = random_case(seed)
        r = oracle.run_cycle(nodes, L, g, c, s)
        order = _rank_order(g)
        rank = np.empty(len(g), dtype=np.int64); rank[order] = np.arange(len(g))
        adm = [gi for gi in order if r["status"]["state"][gi] == T.GANG_ADMITTED]
        for victim in adm[-3:]:   # dropping a gang must leave everything that ranks before it untouched
            g2 = g.copy(); g2["flags"][victim] |= T.GANG_GATED
            r2 = oracle.run_cycle(nodes, L, g2, c, s)
            before = rank < rank[victim]
            assert np.array_equal(r2["status"]["state"][before], r["status"]["state"][before]), (seed, victim)
            for gi in np.nonzero(before)[0]:
                if r["status"]["state"][gi] == T.GANG_ADMITTED:
                    o, k = int(r["status"]["placement_off"][gi]), int(r["status"]["n_pods"][gi])
                    o2 = int(r2["status"]["placement_off"][gi])
                    assert np.array_equal(r["placements"][o: o + k], r2["placements"][o2: o2 + k]), (seed, victim, gi)


def test_c4_shape_priorities_follow_the_podcliqueset(oracle):
    """every PodGang of a PodCliqueSet carries the set's PriorityClassName (podgang/podgang.go:158): in the C4
    generator a scaled gang has its base gang's priority, so the base gang always has its turn first"""
    cfg = synth.config_c4(n=2520, g=400)
    g, c, s = cfg["tables"]
    scaled = g["base_gang"] != T.NONE_U32
    assert scaled.sum() == 300
    assert np.array_equal(g["priority"][scaled], g["priority"][g["base_gang"][scaled]])
    r = oracle.run_cycle(cfg["nodes"], cfg["n_levels"], g, c, s)
    st = r["status"]["state"]
    # nothing is BASE_REJECTED merely because of the order: only behind a base gang that was itself not admitted
    br = np.nonzero(st == T.GANG_BASE_REJECTED)[0]
    assert (st[g["base_gang"][br]] != T.GANG_ADMITTED).all()
