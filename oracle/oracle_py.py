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
