"""Scenarios for the reclaim pass: a cluster on which an earlier cycle admitted gangs (now RUNNING, holding resources), and a new
submission whose gangs may outrank them.  Used by the oracle property tests and the GPU parity tests."""
import numpy as np

from grove_b200 import synth, tables as T


def holdings_of(placements, status, gangs, cliques):
    """(running_dt, holding_dt) of the gangs a cycle admitted: one holding per (gang, node)"""
    run, hold = [], []
    for g in range(len(gangs)):
        if status["state"][g] != T.GANG_ADMITTED:
            continue
        pl = placements[status["placement_off"][g]: status["placement_off"][g] + status["n_pods"][g]]
        per = {}
        for e in pl:
            q = cliques[int(e["clique"])]
            u = per.setdefault(int(e["node"]), [0, 0, 0, 0])
            u[0] += int(q["req_cpu_milli"]); u[1] += int(q["req_mem_mib"]); u[2] += int(q["req_gpu"]); u[3] += 1
        run.append((int(gangs["priority"][g]), len(hold), len(per), 0))
        hold += [(n, u[0], u[1], u[2], u[3]) for n, u in per.items()]
    return (np.array(run, dtype=T.running_dt) if run else np.zeros(0, dtype=T.running_dt),
            np.array(hold, dtype=T.holding_dt) if hold else np.zeros(0, dtype=T.holding_dt))


def churned_cluster(oracle, seed, n=1260, g_running=400, g_pending=200, boost=2, used_pct=80):
    """-> nodes (free now), n_levels, pending tables, running, holdings.  The running gangs come out of an ordinary oracle cycle
    over a C4-shaped submission; the pending submission is another one whose gangs get `boost` added to every other priority,
    so that part of it outranks part of what is running."""
    first = synth.config_c4(n=n, g=g_running, seed=seed, max_used_pct=used_pct)
    g1, c1, s1 = first["tables"]
    ref = oracle.run_cycle(first["nodes"], first["n_levels"], g1, c1, s1, threads=2)
    running, holdings = holdings_of(ref["placements"], ref["status"], g1, c1)
    second = synth.config_c4(n=n, g=g_pending, seed=seed + 7919, max_used_pct=used_pct)
    g2, c2, s2 = (a.copy() for a in second["tables"])
    # a scaled gang keeps its base gang's priority (podgang/podgang.go:158)
    for i in range(len(g2)):
        b = int(g2["base_gang"][i])
        root = i if b == T.NONE_U32 else b
        if root % 2 == 0:
            g2["priority"][i] += boost
    return ref["nodes_after"], first["n_levels"], (g2, c2, s2), running, holdings
