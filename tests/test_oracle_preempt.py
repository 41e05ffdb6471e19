"""Reclaim pass (include/grove_place.h "preemption / reclaim"): properties of the definition, on the oracle.

The reference reserves the OUTCOME (PodGangConditionTypeDisruptionTarget, scheduler/api/core/v1alpha1/podgang.go:166-170: "PodGang
is preempted by a higher priority PodGang"; PriorityClassName podgang.go:62-64) and holds no arithmetic or test for it, so
these are properties any such pass must have, not golden vectors."""
import numpy as np
import pytest

from grove_b200 import synth, tables as T
from preempt_cases import churned_cluster, holdings_of


def _usage(placements, cliques, n):
    u = np.zeros((n, 4), dtype=np.int64)
    for e in placements:
        q = cliques[int(e["clique"])]
        u[int(e["node"])] += (int(q["req_cpu_milli"]), int(q["req_mem_mib"]), int(q["req_gpu"]), 1)
    return u


def _free(nodes):
    return np.stack([nodes["free_cpu_milli"].astype(np.int64), nodes["free_mem_mib"].astype(np.int64),
                     nodes["free_gpu"].astype(np.int64), nodes["free_pods"].astype(np.int64)], axis=1)


def test_high_priority_gang_evicts_a_low_priority_one(oracle):
    """GS-style (the e2e cluster of hack/e2e.yaml: 150 MiB nodes, 80 MiB pods -> one pod a node): six nodes full of a priority-0
    running gang; a priority-5 gang of two pods is rejected by the ordinary pass, admitted by the reclaim pass, and the running
    gang is evicted WHOLE.  A priority-0 gang may not evict its equal and stays rejected (it is only reconsidered while a
    running gang of a lower priority still stands)."""
    nodes = synth.e2e_cluster(6)
    b = T.GangTableBuilder()
    b.add_gang([(None, [dict(mem=80, min=6, class_mask=synth.AGENT)])], priority=0)
    g1, c1, s1 = b.build()
    first = oracle.run_cycle(nodes, 4, g1, c1, s1)
    assert first["status"]["state"][0] == T.GANG_ADMITTED
    running, holdings = holdings_of(first["placements"], first["status"], g1, c1)
    b = T.GangTableBuilder()
    b.add_gang([(None, [dict(mem=80, min=2, class_mask=synth.AGENT)])], priority=5)
    b.add_gang([(None, [dict(mem=80, min=1, class_mask=synth.AGENT)])], priority=0)
    g2, c2, s2 = b.build()
    plain = oracle.run_cycle(first["nodes_after"], 4, g2, c2, s2)
    assert list(plain["status"]["state"]) == [T.GANG_REJECTED, T.GANG_REJECTED]
    out = oracle.run_cycle_preempt(first["nodes_after"], 4, g2, c2, s2, running, holdings)
    assert out["status"]["state"][0] == T.GANG_ADMITTED and out["status"]["reserved0"][0] == T.STATUS_PREEMPTOR
    assert out["status"]["state"][1] == T.GANG_REJECTED
    assert [(int(v["running"]), int(v["preemptor"])) for v in out["victims"]] == [(0, 0)]
    # everything the victim held is free again, minus what the preemptor took
    want = _free(first["nodes_after"]) + _usage(first["placements"], c1, len(nodes)) - _usage(out["placements"], c2, len(nodes))
    assert np.array_equal(_free(out["nodes_after"]), want)


def test_no_running_gangs_is_the_ordinary_pass(oracle):
    cfg = synth.config_c4(n=1260, g=200)
    g, c, s = cfg["tables"]
    a = oracle.run_cycle(cfg["nodes"], cfg["n_levels"], g, c, s)
    b = oracle.run_cycle_preempt(cfg["nodes"], cfg["n_levels"], g, c, s, np.zeros(0, dtype=T.running_dt), np.zeros(0, dtype=T.holding_dt))
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["placements"], b["placements"])
    assert np.array_equal(a["nodes_after"], b["nodes_after"]) and len(b["victims"]) == 0


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_reclaim_invariants(oracle, seed):
    nodes, L, (g, c, s), running, holdings = churned_cluster(oracle, seed)
    plain = oracle.run_cycle(nodes, L, g, c, s)
    out = oracle.run_cycle_preempt(nodes, L, g, c, s, running, holdings)
    st, pst = out["status"], plain["status"]
    pre = (st["reserved0"] & T.STATUS_PREEMPTOR) != 0
    assert pre.any(), "scenario too easy: nobody had to preempt"
    # the ordinary pass's decisions stand; only its REJECTED gangs can change, and only to ADMITTED
    same = ~pre
    assert np.array_equal(st["state"][same], pst["state"][same])
    assert (pst["state"][pre] == T.GANG_REJECTED).all() and (st["state"][pre] == T.GANG_ADMITTED).all()
    for gi in np.nonzero(same & (st["state"] == T.GANG_ADMITTED))[0]:
        a = out["placements"][st["placement_off"][gi]: st["placement_off"][gi] + st["n_pods"][gi]]
        b = plain["placements"][pst["placement_off"][gi]: pst["placement_off"][gi] + pst["n_pods"][gi]]
        assert np.array_equal(a, b)
    # victims: each running gang at most once, always of a strictly lower priority than its preemptor, which is a preemptor
    vr = out["victims"]["running"]
    assert len(set(vr.tolist())) == len(vr)
    for v in out["victims"]:
        assert int(running["priority"][v["running"]]) < int(g["priority"][v["preemptor"]])
        assert pre[v["preemptor"]]
    # resources: free after = free before + what the victims held - what was placed; never negative, never above allocatable
    back = np.zeros((len(nodes), 4), dtype=np.int64)
    for r in vr:
        for h in holdings[running["holding_off"][r]: running["holding_off"][r] + running["n_holdings"][r]]:
            back[int(h["node"])] += (int(h["cpu_milli"]), int(h["mem_mib"]), int(h["gpu"]), int(h["pods"]))
    want = _free(nodes) + back - _usage(out["placements"], c, len(nodes))
    assert (want >= 0).all() and np.array_equal(_free(out["nodes_after"]), want)
    # a preemptor that evicted nobody would have fitted in the ordinary pass: every preemptor has a victim OR lives off
    # what earlier evictions of this pass freed
    assert len(out["victims"]) >= 1
    # gang semantics hold for preemptors too: at least MinReplicas of every clique
    for gi in np.nonzero(pre)[0]:
        pl = out["placements"][st["placement_off"][gi]: st["placement_off"][gi] + st["n_pods"][gi]]
        for cr in range(int(g["n_cliques"][gi])):
            q = int(g["clique_off"][gi]) + cr
            assert (pl["clique"] == q).sum() >= int(c["min_replicas"][q])
