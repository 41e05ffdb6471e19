"""ctypes loader for the CPU oracle (oracle/libgrove_oracle.so).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).
Nothing under grove_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleStats(C.Structure):
    _fields_ = [
        ("gangs_admitted", C.c_uint32), ("gangs_rejected", C.c_uint32), ("pods_bound", C.c_uint32),
        ("threads", C.c_uint32), ("pairs_evaluated", C.c_uint64), ("seconds_total", C.c_double),
        ("non_tree_labels", C.c_uint32), ("reserved", C.c_uint32),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libgrove_oracle.so")
    src = os.path.join(_HERE, "grove_oracle_seq.c")
    hdr = os.path.join(_HERE, "..", "include", "grove_place.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libgrove_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_run_cycle.restype = C.c_int32
        _LIB.oracle_topology.restype = C.c_int32
        _LIB.oracle_abi_version.restype = C.c_uint32
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run_cycle(nodes, n_levels, gangs, cliques, scopes, threads=1, want_matrices=False):
    """The sequential priority-ordered pass.  Returns dict(placements, status, scope_status, nodes_after, perm,
    stats[, fit, score]); fit / score are the K1 / K2 rows over the cycle-start snapshot (sorted node order)."""
    from grove_b200 import tables as T

    n, G, Q, S = len(nodes), len(gangs), len(cliques), len(scopes)
    nodes = np.ascontiguousarray(nodes, dtype=T.node_dt)
    gangs = np.ascontiguousarray(gangs, dtype=T.gang_dt)
    cliques = np.ascontiguousarray(cliques, dtype=T.clique_dt)
    scopes = np.ascontiguousarray(scopes, dtype=T.scope_dt)
    cap = int(cliques["replicas"].astype(np.int64).sum()) if Q else 0
    pl = np.zeros(max(cap, 1), dtype=T.placement_dt)
    st = np.zeros(max(G, 1), dtype=T.status_dt)
    ss = np.zeros(max(S, 1), dtype=T.scope_status_dt)
    nodes_after = np.zeros(max(n, 1), dtype=T.node_dt)
    perm = np.zeros(max(n, 1), dtype=np.uint32)
    words = (n + 31) // 32
    fit = np.zeros((Q, words), dtype=np.uint32) if want_matrices else None
    score = np.zeros((Q, n), dtype=np.uint8) if want_matrices else None
    n_pl = C.c_uint32(0)
    stats = OracleStats()
    rc = lib().oracle_run_cycle(
        _p(nodes), C.c_uint32(n), C.c_uint32(n_levels), _p(gangs), C.c_uint32(G), _p(cliques), C.c_uint32(Q),
        _p(scopes), C.c_uint32(S), C.c_int32(threads), _p(pl), C.c_uint32(len(pl)),
        C.byref(n_pl), _p(st), _p(ss), _p(nodes_after), _p(perm), _p(fit), _p(score), C.byref(stats))
    if rc != 0:
        raise RuntimeError(f"oracle_run_cycle failed: {rc}")
    out = dict(placements=pl[: n_pl.value].copy(), status=st[:G].copy(), scope_status=ss[:S].copy(),
               nodes_after=nodes_after[:n].copy(), perm=perm[:n].copy(),
               stats={k: getattr(stats, k) for k, _ in OracleStats._fields_})
    if want_matrices:
        out["fit"], out["score"] = fit, score
    return out


def topology(nodes, n_levels):
    from grove_b200 import tables as T

    nodes = np.ascontiguousarray(nodes, dtype=T.node_dt)
    n = len(nodes)
    perm = np.zeros(n, dtype=np.uint32)
    dom = np.zeros((n, T.MAX_LEVELS), dtype=np.uint32)
    ndom = np.zeros(T.MAX_LEVELS, dtype=np.uint32)
    nt = C.c_uint32(0)
    rc = lib().oracle_topology(_p(nodes), C.c_uint32(n), C.c_uint32(n_levels), _p(perm), _p(dom), _p(ndom), C.byref(nt))
    if rc != 0:
        raise RuntimeError(f"oracle_topology failed: {rc}")
    return perm, dom, ndom[:n_levels].copy(), nt.value


def _fmix32(x: int) -> int:
    x = (x * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF
    x ^= x >> 16; x = (x * 0x85EBCA6B) & 0xFFFFFFFF; x ^= x >> 13; x = (x * 0xC2B2AE35) & 0xFFFFFFFF; x ^= x >> 16
    return x


def run_cycle_preempt(nodes, n_levels, gangs, cliques, scopes, running, holdings, threads=1):
    """The reclaim pass as its definition reads (include/grove_place.h "preemption / reclaim"; the API reserves the outcome,
    PodGangConditionTypeDisruptionTarget podgang.go:166-170, the arithmetic is ours): the ordinary sequential pass, then every
    REJECTED gang ONE AT A TIME in order rank against free + everything running gangs of a lower priority still hold; a gang that
    fits evicts whole running gangs, node by node in the order of its placement, lowest priority first then highest running
    index, until each node's free resources cover what it puts there.  Small cases only (a full oracle cycle per rejected
    gang).  Returns the ordinary outputs merged with the preemptors', plus `victims` (running, preemptor)."""
    from grove_b200 import tables as T

    base = run_cycle(nodes, n_levels, gangs, cliques, scopes, threads=threads)
    G, N = len(gangs), len(nodes)
    status = base["status"].copy()
    sstat = base["scope_status"].copy()
    per_gang = {g: base["placements"][status["placement_off"][g]: status["placement_off"][g] + status["n_pods"][g]].copy()
                for g in range(G) if status["state"][g] == T.GANG_ADMITTED}
    real = base["nodes_after"].copy()
    perm = base["perm"]
    evicted = np.zeros(len(running), dtype=bool)
    victims = []
    on_node = {}
    for r in range(len(running)):
        for h in holdings[running["holding_off"][r]: running["holding_off"][r] + running["n_holdings"][r]]:
            on_node.setdefault(int(h["node"]), []).append(r)

    def give(tab, h):
        n = int(h["node"])
        tab["free_cpu_milli"][n] += h["cpu_milli"]; tab["free_mem_mib"][n] += h["mem_mib"]
        tab["free_gpu"][n] += h["gpu"]; tab["free_pods"][n] += h["pods"]

    for g in sorted(range(G), key=lambda i: (-int(gangs["priority"][i]), i)):
        if status["state"][g] != T.GANG_REJECTED:
            continue
        p = int(gangs["priority"][g])
        elig = [r for r in range(len(running)) if not evicted[r] and int(running["priority"][r]) < p]
        if not elig:
            continue
        view = real.copy()
        for r in elig:
            for h in holdings[running["holding_off"][r]: running["holding_off"][r] + running["n_holdings"][r]]:
                give(view, h)
        sg = gangs[g: g + 1].copy()
        co, so = int(sg["clique_off"][0]), int(sg["scope_off"][0])
        sq = cliques[co: co + int(sg["n_cliques"][0])].copy()
        ss = scopes[so: so + int(sg["n_scopes"][0])].copy()
        if sg["anchor_node"][0] == T.NONE_U32:
            sg["anchor_node"][0] = perm[_fmix32(g) % N]   # the anchor the ordinary pass derived for this gang
        sg["base_gang"][0] = T.NONE_U32; sg["clique_off"][0] = 0; sg["scope_off"][0] = 0
        out = run_cycle(view, n_levels, sg, sq, ss, threads=threads)
        if out["status"]["state"][0] != T.GANG_ADMITTED:
            continue
        pl = out["placements"].copy()
        use, order = {}, []
        for e in pl:
            n, q = int(e["node"]), sq[int(e["clique"])]
            if n not in use:
                use[n] = [0, 0, 0, 0]; order.append(n)
            use[n][0] += int(q["req_cpu_milli"]); use[n][1] += int(q["req_mem_mib"]); use[n][2] += int(q["req_gpu"]); use[n][3] += 1
        for n in order:
            u = use[n]
            while not (real["free_cpu_milli"][n] >= u[0] and real["free_mem_mib"][n] >= u[1] and real["free_gpu"][n] >= u[2] and real["free_pods"][n] >= u[3]):
                cands = [r for r in on_node.get(n, []) if not evicted[r] and int(running["priority"][r]) < p]
                v = min(cands, key=lambda r: (int(running["priority"][r]), -r))
                evicted[v] = True
                for h in holdings[running["holding_off"][v]: running["holding_off"][v] + running["n_holdings"][v]]:
                    give(real, h)
                victims.append((v, g))
            real["free_cpu_milli"][n] -= u[0]; real["free_mem_mib"][n] -= u[1]; real["free_gpu"][n] -= u[2]; real["free_pods"][n] -= u[3]
        pl["clique"] += co
        per_gang[g] = pl
        st = out["status"][0].copy()
        st["reserved0"] = T.STATUS_PREEMPTOR
        status[g] = st
        sstat[so: so + len(ss)] = out["scope_status"]
    merged = []
    off = 0
    for g in range(G):
        if status["state"][g] == T.GANG_ADMITTED:
            status["placement_off"][g] = off
            merged.append(per_gang[g]); off += len(per_gang[g])
    placements = np.concatenate(merged) if merged else np.zeros(0, dtype=T.placement_dt)
    return dict(placements=placements, status=status, scope_status=sstat, nodes_after=real,
                victims=np.array(victims, dtype=T.victim_dt) if victims else np.zeros(0, dtype=T.victim_dt), perm=perm)


def shard_cut(n, dom0_sorted, rank, world):
    """first node (topology-sorted index) of rank's shard: the table cut at top-level domain boundaries, about n / world nodes
    each (include/grove_place.h "multi-GPU score pass"); nodes without the top-level label stay with the last rank"""
    if rank <= 0:
        return 0
    if rank >= world:
        return n
    target = n * rank // world
    present = dom0_sorted != 0xFFFFFFFF
    starts = np.nonzero(present & np.concatenate([[True], dom0_sorted[1:] != dom0_sorted[:-1]]))[0]
    later = starts[starts >= target]
    return int(later[0]) if len(later) else n


def shard_summary(nodes, n_levels, gangs, cliques, scopes, lo, hi):
    """numpy restatement of the shard summary of the multi-GPU score pass: int32[G + Q] over the topology-sorted node range
    [lo, hi) -- [g]: domains of gang g's Required level starting in the range in which every clique of the gang finds MinReplicas
    worth of capacity on its own; [G + q]: pods of clique q that fit on the range's nodes (a node counts at most 255)."""
    from grove_b200 import tables as T

    perm, dom, _, _ = topology(nodes, n_levels)
    sn = np.ascontiguousarray(nodes, dtype=T.node_dt)[perm]
    n, G, Q = len(sn), len(gangs), len(cliques)
    absent = dom[:, :n_levels] == 0xFFFFFFFF
    vdepth = np.where(absent.any(axis=1), absent.argmax(axis=1), n_levels)
    sched = (sn["flags"] & T.NODE_SCHEDULABLE) != 0
    ncls = (sn["flags"].astype(np.int64) >> T.NODE_CLASS_SHIFT) & 0xF
    out = np.zeros(G + Q, dtype=np.int32)
    caps = {}
    for g in range(G):
        gg = gangs[g]
        co, so = int(gg["clique_off"]), int(gg["scope_off"])
        rows = []
        for c in range(int(gg["n_cliques"])):
            q = cliques[co + c]
            sc = scopes[so + (int(q["scope"]) & 0x1F)]
            need = 0
            for lv in (int(gg["level"]), int(sc["level"]), int(q["level"])):
                if lv != T.LEVEL_NONE:
                    need = max(need, lv + 1)
            ok = sched & (((int(q["class_mask"]) >> ncls) & 1) != 0) & (vdepth >= need)
            cap = sn["free_pods"].astype(np.int64)
            for free, req in ((sn["free_cpu_milli"], int(q["req_cpu_milli"])), (sn["free_mem_mib"], int(q["req_mem_mib"])), (sn["free_gpu"], int(q["req_gpu"]))):
                if req:
                    cap = np.minimum(cap, free.astype(np.int64) // req)
            cap = np.where(ok, np.minimum(cap, 255), 0)
            rows.append(cap)
            out[G + co + c] = int(cap[lo:hi].sum())
        lv = int(gg["level"])
        if lv != T.LEVEL_NONE:
            ids = dom[:, lv].astype(np.int64)
            has = ids != 0xFFFFFFFF
            nd = int(ids[has].max()) + 1 if has.any() else 0
            first = np.full(nd, n, dtype=np.int64)
            np.minimum.at(first, ids[has], np.nonzero(has)[0])
            good = (first >= lo) & (first < hi)
            for c, cap in enumerate(rows):
                m = int(cliques[co + c]["min_replicas"])
                if m:
                    good &= np.bincount(ids[has], weights=cap[has], minlength=nd) >= m
            out[g] = int(good.sum())
    return out
