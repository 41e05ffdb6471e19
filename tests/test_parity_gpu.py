"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import numpy as np
import pytest

from grove_b200 import synth, tables as T

pytestmark = pytest.mark.gpu


def _run_engine(nodes, L, tabs, matrices=False, window=0):
    from grove_b200.engine import PlacementEngine
    g, c, s = tabs
    with PlacementEngine(L, window=window) as e:
        e.load_nodes(nodes)
        e.submit_gangs(g, c, s)
        stats = e.run_cycle()
        out = dict(stats=stats, placements=e.placements(), status=e.gang_status(), scope_status=e.scope_domains(),
                   nodes_after=e.nodes(), perm=e.debug_perm())
        if matrices:
            out["fit"] = np.stack([e.debug_fit_row(q) for q in range(len(c))]) if len(c) else None
            out["score"] = np.stack([e.debug_score_row(q) for q in range(len(c))]) if len(c) else None
    return out


def assert_same(gpu, ref):
    """every output of the cycle: the engine's relaxation must land on the sequential oracle's answer bit for bit"""
    assert np.array_equal(gpu["perm"], ref["perm"])
    for f in ("state", "level", "n_pods", "placement_off", "score_num", "score_den", "domain_node"):
        assert np.array_equal(gpu["status"][f], ref["status"][f]), f
    assert np.array_equal(gpu["status"], ref["status"])
    assert np.array_equal(gpu["scope_status"], ref["scope_status"])
    assert np.array_equal(gpu["placements"], ref["placements"])
    assert np.array_equal(gpu["nodes_after"], ref["nodes_after"])
    assert gpu["stats"]["gangs_admitted"] == ref["stats"]["gangs_admitted"] and gpu["stats"]["pods_bound"] == ref["stats"]["pods_bound"]


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3"])
def test_config_parity(built_lib, oracle, cfg):
    c = synth.CONFIGS[cfg]()
    g, cl, sc = c["tables"]
    ref = oracle.run_cycle(c["nodes"], c["n_levels"], g, cl, sc, threads=8)
    gpu = _run_engine(c["nodes"], c["n_levels"], c["tables"])
    assert_same(gpu, ref)


@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_fit_and_score_matrices(built_lib, oracle, cfg):
    """K1 fit bitmap and K2 score matrix over the cycle-start snapshot, every (clique, node) pair."""
    c = synth.CONFIGS[cfg]() if cfg == "C2" else synth.config_c3(n=2520, g=250)
    g, cl, sc = c["tables"]
    ref = oracle.run_cycle(c["nodes"], c["n_levels"], g, cl, sc, threads=8, want_matrices=True)
    gpu = _run_engine(c["nodes"], c["n_levels"], c["tables"], matrices=True)
    assert np.array_equal(gpu["fit"], ref["fit"])
    assert np.array_equal(gpu["score"], ref["score"])
    assert_same(gpu, ref)


def test_c4_reduced(built_lib, oracle):
    c = synth.config_c4(n=5040, g=1000)
    g, cl, sc = c["tables"]
    ref = oracle.run_cycle(c["nodes"], c["n_levels"], g, cl, sc, threads=8)
    gpu = _run_engine(c["nodes"], c["n_levels"], c["tables"])
    assert_same(gpu, ref)


@pytest.mark.parametrize("window", [1, 7, 64, 1000])
def test_window_size_never_changes_a_result(built_lib, oracle, window):
    """the relaxation window is a tuning knob: one gang at a time (window 1 = the sequential pass itself on the GPU),
    a few, or all of them at once must give the same answer"""
    c = synth.config_c4(n=5040, g=1000)
    g, cl, sc = c["tables"]
    ref = oracle.run_cycle(c["nodes"], c["n_levels"], g, cl, sc, threads=8)
    assert_same(_run_engine(c["nodes"], c["n_levels"], c["tables"], window=window), ref)


def test_update_nodes_and_second_cycle(built_lib, oracle):
    """churn: after a cycle, free some nodes through grove_update_nodes and run the next cycle on the
    committed state; the oracle follows with the same inputs."""
    from grove_b200.engine import PlacementEngine
    c = synth.config_c3(n=756, g=150)
    g, cl, sc = c["tables"]
    with PlacementEngine(c["n_levels"]) as e:
        e.load_nodes(c["nodes"]); e.submit_gangs(g, cl, sc)
        e.run_cycle()
        after = e.nodes()
        ref1 = oracle.run_cycle(c["nodes"], c["n_levels"], g, cl, sc)
        assert np.array_equal(after, ref1["nodes_after"])
        idx = np.arange(0, 756, 7, dtype=np.uint32)
        recs = c["nodes"][idx].copy()  # these nodes drained back to their initial free state
        e.update_nodes(idx, recs)
        state2 = after.copy(); state2[idx] = recs
        assert np.array_equal(e.nodes(), state2)
        e.submit_gangs(g, cl, sc)
        e.run_cycle()
        ref2 = oracle.run_cycle(state2, c["n_levels"], g, cl, sc)
        assert np.array_equal(e.placements(), ref2["placements"])
        assert np.array_equal(e.gang_status(), ref2["status"])
        assert np.array_equal(e.nodes(), ref2["nodes_after"])


def test_c4_full_size_bit_exact(built_lib, oracle):
    """BASELINE.json's metric configuration (50 000 nodes / 10 000 PodGangs / 27 500 PodCliques) against the sequential
    oracle: every status field, every scope domain, every placement, the committed node table"""
    c = synth.config_c4()
    g, cl, sc = c["tables"]
    ref = oracle.run_cycle(c["nodes"], c["n_levels"], g, cl, sc, threads=16)
    gpu = _run_engine(c["nodes"], c["n_levels"], c["tables"])
    assert_same(gpu, ref)
    st = ref["status"]["state"]
    assert (st == T.GANG_ADMITTED).sum() > 3000 and (st == T.GANG_REJECTED).sum() > 500


def test_full_size_properties(built_lib):
    """BASELINE.json's full C4 size, through size-independent properties (no oracle needed):
    no node is over-committed, every admitted gang has all its MinReplicas, every scope's pods share its
    Required domain, rejected gangs bound nothing, and resources are conserved."""
    from grove_b200.engine import PlacementEngine
    c = synth.config_c4()
    g, cl, sc = c["tables"]
    nodes = c["nodes"]
    with PlacementEngine(c["n_levels"]) as e:
        e.load_nodes(nodes); e.submit_gangs(g, cl, sc)
        st = e.run_cycle()
        pl, gs, after = e.placements(), e.gang_status(), e.nodes()
    assert st["gangs_admitted"] + st["gangs_rejected"] == len(g)
    # conservation: free_before - free_after == sum of requests bound per node
    for f, req in (("free_cpu_milli", "req_cpu_milli"), ("free_mem_mib", "req_mem_mib"), ("free_gpu", "req_gpu")):
        used = np.bincount(pl["node"], weights=cl[req][pl["clique"]].astype(np.float64), minlength=len(nodes))
        assert np.array_equal(nodes[f].astype(np.int64) - after[f].astype(np.int64), used.astype(np.int64)), f
    pods = np.bincount(pl["node"], minlength=len(nodes))
    assert np.array_equal(nodes["free_pods"].astype(np.int64) - after["free_pods"].astype(np.int64), pods)
    assert (after["free_cpu_milli"] <= nodes["free_cpu_milli"]).all() and (after["free_gpu"] <= nodes["free_gpu"]).all()
    assert (nodes["flags"][pl["node"]] & T.NODE_SCHEDULABLE).all()
    cls = (nodes["flags"][pl["node"]] >> T.NODE_CLASS_SHIFT) & 0xF
    assert ((cl["class_mask"][pl["clique"]] >> cls) & 1).all()
    per_clique = np.bincount(pl["clique"], minlength=len(cl))
    clique_gang = np.repeat(np.arange(len(g)), g["n_cliques"])
    adm = gs["state"][clique_gang] == T.GANG_ADMITTED
    assert (per_clique[adm] >= cl["min_replicas"][adm]).all() and (per_clique[adm] <= cl["replicas"][adm]).all()
    assert (per_clique[~adm] == 0).all()
    # Required domains: gang -> block (level 1), scaling-group scope -> rack (2), leader -> host (3)
    dom = nodes["dom"][pl["node"]]
    gang_of = clique_gang[pl["clique"]]
    for lvl_field, lvl_src, key in (("gang", g["level"][gang_of], gang_of),):
        lv = lvl_src.astype(np.int64)
        sel = lv != T.LEVEL_NONE
        val = dom[np.arange(len(pl)), np.where(sel, lv, 0)]
        first = {}
        for k, v, ok in zip(key.tolist(), val.tolist(), sel.tolist()):
            if ok:
                assert first.setdefault(k, v) == v
    scope_id = np.repeat(np.arange(len(sc)), sc["n_cliques"])  # scopes tile the clique table in order
    lv = sc["level"][scope_id[pl["clique"]]].astype(np.int64)
    first = {}
    for k, l, row in zip(scope_id[pl["clique"]].tolist(), lv.tolist(), dom.tolist()):
        if l != T.LEVEL_NONE:
            assert first.setdefault(k, row[l]) == row[l]
    lv = cl["level"][pl["clique"]].astype(np.int64)
    first = {}
    for k, l, row in zip(pl["clique"].tolist(), lv.tolist(), dom.tolist()):
        if l != T.LEVEL_NONE:
            assert first.setdefault(k, row[l]) == row[l]


def test_c5_churn_matches_oracle_tick_by_tick(built_lib, oracle):
    """BASELINE.json config 5 (steady-state churn): arrivals every tick, pending gangs carried over,
    finished gangs release their nodes through grove_update_nodes; the engine keeps its node table
    across ticks, the oracle is re-fed the same table; every tick must match bit for bit."""
    from grove_b200.engine import PlacementEngine
    ch = synth.ChurnC5(n=50000, arrivals=100, release_pct=5)
    with PlacementEngine(ch.n_levels) as e:
        e.load_nodes(ch.nodes)
        for tick in range(12):
            specs, tabs = ch.begin_tick()
            g, c, s = tabs
            before = ch.nodes
            e.submit_gangs(g, c, s)
            e.run_cycle()
            pl, st, after = e.placements(), e.gang_status(), e.nodes()
            ref = oracle.run_cycle(before, ch.n_levels, g, c, s, threads=8)
            assert np.array_equal(pl, ref["placements"]), tick
            assert np.array_equal(st, ref["status"]), tick
            assert np.array_equal(after, ref["nodes_after"]), tick
            idx, recs = ch.end_tick(specs, tabs, st, pl, after)
            if len(idx):
                e.update_nodes(idx, recs)
            assert np.array_equal(e.nodes(), ch.nodes), tick
        assert ch.tick == 12 and len(ch.running) > 0


@pytest.mark.parametrize("env", [{"GROVE_TUNE_WINDOW": "97"}, {"GROVE_TUNE_REFRESH": "1"}, {"GROVE_TUNE_REFRESH": "1000000"},
                                 {"GROVE_TUNE_EVAL_CTAS": "1", "GROVE_TUNE_SCORE": "1"}, {"GROVE_TUNE_MAX_ATT": "1", "GROVE_TUNE_HEAVY_ATT": "1"},
                                 {"GROVE_TUNE_MAX_ATT": "0", "GROVE_TUNE_HEAVY_ATT": "200"}])
def test_tuning_knobs_never_change_a_result(built_lib, oracle, env):
    """window size, how often the capacity tables are rebuilt (every round / never: the evaluator then leans on the
    per-node slow path), grid sizes, the score matrix built beside the relaxation, every gang heavy after one attempt / no gang ever heavy:
    all must give the oracle's answer."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent('''
        import sys, numpy as np
        sys.path.insert(0, %r)
        from grove_b200 import synth
        from grove_b200.engine import PlacementEngine
        from oracle import oracle_py as O
        for cfg in (synth.config_c3(n=2016, g=300), synth.config_c4(n=5040, g=800), synth.config_c2(n=500, g=80)):
            g, c, s = cfg["tables"]
            ref = O.run_cycle(cfg["nodes"], cfg["n_levels"], g, c, s, threads=4)
            with PlacementEngine(cfg["n_levels"]) as e:
                e.load_nodes(cfg["nodes"]); e.submit_gangs(g, c, s); e.run_cycle()
                assert np.array_equal(e.placements(), ref["placements"])
                assert np.array_equal(e.gang_status(), ref["status"])
                assert np.array_equal(e.nodes(), ref["nodes_after"])
        sys.path.insert(0, %r)
        from test_random_parity_gpu import random_case
        for seed in list(range(3000, 3030)) + [4000]:   # Preferred levels through the same paths
            nodes, L, (g, c, s) = random_case(seed, big=seed >= 4000, pref=True)
            ref = O.run_cycle(nodes, L, g, c, s, threads=4)
            with PlacementEngine(L) as e:
                e.load_nodes(nodes); e.submit_gangs(g, c, s); e.run_cycle()
                assert np.array_equal(e.placements(), ref["placements"]), seed
                assert np.array_equal(e.gang_status(), ref["status"]), seed
        print("ok")
    ''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env={**os.environ, **env})
    assert r.returncode == 0 and "ok" in r.stdout, (env, r.stderr[-1500:])


def test_edge_cases(built_lib, oracle):
    """empty submission, single node, everything gated, MinReplicas = 0 everywhere, API misuse"""
    from grove_b200.engine import GroveError, PlacementEngine
    nodes = synth.e2e_cluster(5)
    with PlacementEngine(4) as e:
        with pytest.raises(GroveError):
            e.run_cycle()                      # nothing loaded yet
        e.load_nodes(nodes)
        with pytest.raises(GroveError):
            e.run_cycle()                      # no gangs submitted
        g, c, s = T.GangTableBuilder().build()
        e.submit_gangs(g, c, s)                # empty submission is a valid (no-op) cycle
        st = e.run_cycle()
        assert st["rounds"] == 0 and st["gangs_admitted"] == 0 and len(e.placements()) == 0 and len(e.scope_domains()) == 0
        b = T.GangTableBuilder()
        b.add_gang([(None, [dict(mem=80, min=1, class_mask=synth.AGENT)])], gated=True)
        b.add_gang([(None, [dict(mem=80, min=0, replicas=0, class_mask=synth.AGENT)])])
        b.add_gang([(None, [dict(mem=80, min=0, replicas=3, class_mask=synth.AGENT)])], level=2)
        g, c, s = b.build()
        e.submit_gangs(g, c, s)
        e.run_cycle()
        ref = oracle.run_cycle(nodes, 4, g, c, s)
        assert np.array_equal(e.gang_status(), ref["status"]) and np.array_equal(e.placements(), ref["placements"])
        assert ref["status"]["state"].tolist() == [T.GANG_GATED_SKIP, T.GANG_ADMITTED, T.GANG_ADMITTED]
        bad = g.copy(); bad["level"][2] = 9
        with pytest.raises(GroveError):
            e.submit_gangs(bad, c, s)
        with pytest.raises(GroveError):
            e.update_nodes(np.array([99], dtype=np.uint32), nodes[:1])
    one = synth.e2e_cluster(1)
    b = T.GangTableBuilder(); b.add_gang([(None, [dict(mem=80, min=1, class_mask=synth.AGENT)])], level=3)
    g, c, s = b.build()
    with PlacementEngine(4) as e:
        e.load_nodes(one); e.submit_gangs(g, c, s); e.run_cycle()
        assert e.gang_status()["state"][0] == T.GANG_ADMITTED and e.placements()["node"][0] == 0


def test_preferred_levels_match_the_oracle(built_lib, oracle):
    """the scenarios of tests/test_oracle_preferred.py (fits at the Preferred level / widens one level at a time /
    bounded by Required / whole-cluster fallback / clique- and scope-level), many gangs competing"""
    from grove_b200.engine import GroveError, PlacementEngine
    A = synth.AGENT
    clq = lambda n, **kw: dict(mem=40, min=n, class_mask=A, **kw)
    nodes = synth.e2e_cluster(112)
    b = T.GangTableBuilder()
    for i in range(12):
        b.add_gang([(None, [clq(10), clq(8)])], preferred=2, anchor=(i * 9) % 112)
        b.add_gang([(None, [clq(15), clq(15)])], preferred=2, level=1 if i % 2 else None)
        b.add_gang([(None, [clq(25), clq(25)])], preferred=3)
        b.add_gang([(None, [clq(3, preferred=3), clq(4, preferred=3)]), (None, [clq(6), clq(6)], 2), (1, [clq(12), clq(12)], 2)],
                   priority=i % 3)
    g, c, s = b.build()
    ref = oracle.run_cycle(nodes, synth.E2E_LEVELS, g, c, s, threads=4)
    assert (ref["status"]["state"] == T.GANG_ADMITTED).sum() > 5 and (ref["status"]["state"] == T.GANG_REJECTED).sum() > 5
    with PlacementEngine(synth.E2E_LEVELS) as e:
        e.load_nodes(nodes); e.submit_gangs(g, c, s)
        e.run_cycle()
        assert np.array_equal(e.gang_status(), ref["status"])
        assert np.array_equal(e.placements(), ref["placements"])
        assert np.array_equal(e.nodes(), ref["nodes_after"])
        for kw in (dict(level=2, preferred=2), dict(level=2, preferred=1), dict(preferred=7)):
            bad = T.GangTableBuilder(); bad.add_gang([(None, [clq(1)])], **kw)
            with pytest.raises(GroveError):
                e.submit_gangs(*bad.build())
        bad = T.GangTableBuilder(); bad.add_gang([(2, [clq(1)], 1)])
        with pytest.raises(GroveError):
            e.submit_gangs(*bad.build())
        bad = T.GangTableBuilder(); bad.add_gang([(None, [clq(1, level=2, preferred=2)])])
        with pytest.raises(GroveError):
            e.submit_gangs(*bad.build())


def test_e2e_sequences_match_the_oracle_step_by_step(built_lib, oracle):
    """the reference's multi-step suites (tests/test_oracle_e2e_sequences.py: GS2-GS12) through the engine: every
    scheduling pass of every scenario must equal the oracle's, and the pod counts must be the suites'"""
    from grove_b200.engine import PlacementEngine
    from test_oracle_e2e_sequences import GS, replay
    with PlacementEngine(synth.E2E_LEVELS) as e:
        def place(nodes, g, c, s):
            ref = oracle.run_cycle(nodes, synth.E2E_LEVELS, g, c, s)
            e.load_nodes(nodes); e.submit_gangs(g, c, s); e.run_cycle()
            out = dict(status=e.gang_status(), placements=e.placements(), nodes_after=e.nodes())
            for k in out:
                assert np.array_equal(out[k], ref[k]), k
            return out
        for name in GS:
            for kw in (dict(), dict(split_surplus=True, same_pass_unlock=False)):   # surplus with the gang / minimums first
                sim = replay(place, GS[name], **kw)
                assert sim.running() == sim.pods(), name


def test_more_priorities_than_priority_classes(built_lib, oracle):
    """order ranks come from a counting sort over the distinct priorities (few PriorityClasses) and fall back to a
    stable sort beyond 64 of them; both must order conflicts like the oracle"""
    from grove_b200.engine import PlacementEngine
    rng = np.random.default_rng(11)
    for n_prio, G in ((3, 3000), (64, 2500), (65, 300), (500, 2600)):
        nodes = synth.kwok_nodes(1200, [240, 48, 8, 1])
        b = T.GangTableBuilder()
        for _ in range(G):
            b.add_gang([(2, [dict(cpu=16000, mem=65536, gpu=4, min=2)])], level=1, priority=int(rng.integers(0, n_prio)) - n_prio // 2,
                       anchor=int(rng.integers(0, 1200)))
        g, c, s = b.build()
        ref = oracle.run_cycle(nodes, 4, g, c, s, threads=8)
        with PlacementEngine(4) as e:
            e.load_nodes(nodes); e.submit_gangs(g, c, s); e.run_cycle()
            assert np.array_equal(e.gang_status(), ref["status"]), n_prio
            assert np.array_equal(e.placements(), ref["placements"]), n_prio
