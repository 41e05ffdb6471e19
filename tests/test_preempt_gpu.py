"""Reclaim pass through the C ABI (grove_run_cycle_preempt / grove_get_victims) against the oracle's gang-by-gang statement of
it.  The engine re-submits the rejected gangs a whole priority class at a time; the oracle reconsiders them one at a time:
identical outputs say that the batching is exact."""
import numpy as np
import pytest

from grove_b200 import synth, tables as T
from preempt_cases import churned_cluster, holdings_of

pytestmark = pytest.mark.gpu


def _same(e, ref):
    assert np.array_equal(e.gang_status(), ref["status"])
    assert np.array_equal(e.placements(), ref["placements"])
    assert np.array_equal(e.scope_domains(), ref["scope_status"])
    assert np.array_equal(e.nodes(), ref["nodes_after"])
    assert np.array_equal(e.victims(), ref["victims"])


def test_gs_style_eviction(built_lib, oracle):
    from grove_b200.engine import PlacementEngine
    nodes = synth.e2e_cluster(6)
    b = T.GangTableBuilder()
    b.add_gang([(None, [dict(mem=80, min=6, class_mask=synth.AGENT)])], priority=0)
    g1, c1, s1 = b.build()
    b = T.GangTableBuilder()
    b.add_gang([(None, [dict(mem=80, min=2, class_mask=synth.AGENT)])], priority=5)
    b.add_gang([(None, [dict(mem=80, min=1, class_mask=synth.AGENT)])], priority=0)
    g2, c2, s2 = b.build()
    with PlacementEngine(4) as e:
        e.load_nodes(nodes); e.submit_gangs(g1, c1, s1); e.run_cycle()
        running, holdings = holdings_of(e.placements(), e.gang_status(), g1, c1)
        assert len(running) == 1 and len(e.victims()) == 0
        now = e.nodes()
        e.submit_gangs(g2, c2, s2)
        st = e.run_cycle_preempt(running, holdings)
        ref = oracle.run_cycle_preempt(now, 4, g2, c2, s2, running, holdings)
        _same(e, ref)
        assert st["gangs_admitted"] == 1 and st["gangs_rejected"] == 1
        assert e.gang_status()["reserved0"][0] == T.STATUS_PREEMPTOR
        assert [(int(v["running"]), int(v["preemptor"])) for v in e.victims()] == [(0, 0)]
        # the handle is back on the caller's submission and the really free node table: an ordinary cycle evicts nobody
        e.run_cycle()
        assert len(e.victims()) == 0


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_reclaim_matches_oracle(built_lib, oracle, seed):
    from grove_b200.engine import PlacementEngine
    nodes, L, (g, c, s), running, holdings = churned_cluster(oracle, seed)
    ref = oracle.run_cycle_preempt(nodes, L, g, c, s, running, holdings)
    assert ((ref["status"]["reserved0"] & T.STATUS_PREEMPTOR) != 0).any()
    with PlacementEngine(L) as e:
        e.load_nodes(nodes); e.submit_gangs(g, c, s)
        e.run_cycle_preempt(running, holdings)
        _same(e, ref)


def test_without_running_gangs_it_is_the_ordinary_cycle(built_lib, oracle):
    from grove_b200.engine import PlacementEngine
    cfg = synth.config_c4(n=1260, g=200)
    g, c, s = cfg["tables"]
    ref = oracle.run_cycle(cfg["nodes"], cfg["n_levels"], g, c, s)
    with PlacementEngine(cfg["n_levels"]) as e:
        e.load_nodes(cfg["nodes"]); e.submit_gangs(g, c, s)
        e.run_cycle_preempt(np.zeros(0, dtype=T.running_dt), np.zeros(0, dtype=T.holding_dt))
        assert np.array_equal(e.gang_status(), ref["status"]) and np.array_equal(e.placements(), ref["placements"])
        assert np.array_equal(e.nodes(), ref["nodes_after"]) and len(e.victims()) == 0
