"""Packed table formats of the placement engine (numpy mirrors of include/grove_place.h).

The record layouts are the wire format between the host side (a scheduler backend's cycle loop)
and libgrove_place.so.  Field meanings follow the reference's PodGang schema
(/root/reference scheduler/api/core/v1alpha1/podgang.go:51-131); see include/grove_place.h.
"""
from __future__ import annotations

import numpy as np

MAX_LEVELS = 4
LEVEL_NONE = 0xFF
DOM_ABSENT = 0xFFFFFFFF
NONE_U32 = 0xFFFFFFFF
MAX_GANG_PODS = 128
MAX_GANG_CLIQUES = 32
MAX_GANG_SCOPES = 32

NODE_SCHEDULABLE = 0x1
NODE_CLASS_SHIFT = 8

GANG_GATED = 0x1

GANG_PENDING, GANG_ADMITTED, GANG_REJECTED, GANG_GATED_SKIP, GANG_BASE_REJECTED = 0, 1, 2, 3, 4

node_dt = np.dtype([
    ("free_cpu_milli", "<u4"), ("free_mem_mib", "<u4"), ("free_gpu", "<u2"), ("free_pods", "<u2"),
    ("flags", "<u4"), ("dom", "<u4", (MAX_LEVELS,)),
])
clique_dt = np.dtype([
    ("req_cpu_milli", "<u4"), ("req_mem_mib", "<u4"), ("req_gpu", "<u2"), ("min_replicas", "u1"),
    ("replicas", "u1"), ("class_mask", "<u2"), ("level", "u1"), ("scope", "u1"),
])
scope_dt = np.dtype([
    ("first_clique", "<u2"), ("n_cliques", "<u2"), ("level", "u1"), ("preferred1", "u1"), ("reserved", "u1", (2,)),
])
gang_dt = np.dtype([
    ("clique_off", "<u4"), ("scope_off", "<u4"), ("n_cliques", "<u2"), ("n_scopes", "<u2"),
    ("priority", "<i4"), ("anchor_node", "<u4"), ("base_gang", "<u4"), ("level", "u1"),
    ("preferred", "u1"), ("flags", "<u2"), ("reserved", "<u4"),
])
placement_dt = np.dtype([("clique", "<u4"), ("node", "<u4")])

# preemption / reclaim inputs and outputs (include/grove_place.h)
holding_dt = np.dtype([("node", np.uint32), ("cpu_milli", np.uint32), ("mem_mib", np.uint32), ("gpu", np.uint16), ("pods", np.uint16)])
running_dt = np.dtype([("priority", np.int32), ("holding_off", np.uint32), ("n_holdings", np.uint32), ("reserved", np.uint32)])
victim_dt = np.dtype([("running", np.uint32), ("preemptor", np.uint32)])
STATUS_PREEMPTOR = 0x1
status_dt = np.dtype([
    ("state", "u1"), ("level", "u1"), ("reserved0", "<u2"), ("score_num", "<u2"), ("score_den", "<u2"),
    ("n_pods", "<u4"), ("placement_off", "<u4"), ("domain_node", "<u4"), ("reserved1", "<u4", (3,)),
])
scope_status_dt = np.dtype([("level", "u1"), ("reserved", "u1", (3,)), ("domain_node", "<u4")])
config_dt = np.dtype([
    ("abi_version", "<u4"), ("device", "<i4"), ("n_levels", "<u4"), ("window", "<u4"),
    ("rank", "<u4"), ("world", "<u4"), ("reserved", "<u4", (2,)),
])
stats_dt = np.dtype([
    ("rounds", "<u4"), ("gangs_admitted", "<u4"), ("gangs_rejected", "<u4"), ("pods_bound", "<u4"),
    ("pairs_evaluated", "<u8"), ("kernel_launches", "<u8"), ("evaluations", "<u8"), ("ms_fit", "<f4"), ("ms_score", "<f4"),
    ("ms_admit", "<f4"), ("ms_commit", "<f4"), ("ms_total", "<f4"), ("reserved", "<f4"),
])

assert node_dt.itemsize == 32 and clique_dt.itemsize == 16 and scope_dt.itemsize == 8
assert gang_dt.itemsize == 32 and placement_dt.itemsize == 8 and status_dt.itemsize == 32 and scope_status_dt.itemsize == 8
assert config_dt.itemsize == 32 and stats_dt.itemsize == 64


def make_nodes(n: int) -> np.ndarray:
    nodes = np.zeros(n, dtype=node_dt)
    nodes["dom"][:] = DOM_ABSENT
    return nodes


class GangTableBuilder:
    """Accumulates PodGangs into the three packed tables.

    One `add_gang` call is one PodGang: `scopes` is a list of (level, [clique dict, ...]) or
    (level, [clique dict, ...], preferred); a scope with level None is the implicit scope of PodGroups
    that are in no TopologyConstraintGroupConfig.  Clique dict keys: cpu, mem, gpu (per-pod request),
    min, replicas (default = min), level, preferred, class_mask (default 0xFFFF).  `preferred` is the
    PackConstraint.Preferred level (podgang.go:110-117), deeper than the Required one.
    """

    def __init__(self) -> None:
        self.gangs: list[tuple] = []
        self.cliques: list[tuple] = []
        self.scopes: list[tuple] = []

    @staticmethod
    def _lvl(level) -> int:
        return LEVEL_NONE if level is None else int(level)

    @staticmethod
    def _pref1(level) -> int:
        return 0 if level is None else int(level) + 1

    def add_gang(self, scopes, level=None, priority=0, anchor=None, base=None, gated=False, preferred=None) -> int:
        clique_off, scope_off = len(self.cliques), len(self.scopes)
        rel = 0
        for si, scope in enumerate(scopes):
            slevel, cliques = scope[0], scope[1]
            spref = scope[2] if len(scope) > 2 else None
            self.scopes.append((rel, len(cliques), self._lvl(slevel), self._pref1(spref)))
            for c in cliques:
                mn = int(c.get("min", 1))
                self.cliques.append((
                    int(c.get("cpu", 0)), int(c.get("mem", 0)), int(c.get("gpu", 0)), mn,
                    int(c.get("replicas", mn)), int(c.get("class_mask", 0xFFFF)), self._lvl(c.get("level")),
                    si | (self._pref1(c.get("preferred")) << 5),
                ))
                rel += 1
        self.gangs.append((
            clique_off, scope_off, rel, len(scopes), int(priority),
            NONE_U32 if anchor is None else int(anchor), NONE_U32 if base is None else int(base),
            self._lvl(level), GANG_GATED if gated else 0, self._lvl(preferred),
        ))
        return len(self.gangs) - 1

    def build(self):
        gangs = np.zeros(len(self.gangs), dtype=gang_dt)
        cliques = np.zeros(len(self.cliques), dtype=clique_dt)
        scopes = np.zeros(len(self.scopes), dtype=scope_dt)
        for i, g in enumerate(self.gangs):
            (gangs["clique_off"][i], gangs["scope_off"][i], gangs["n_cliques"][i], gangs["n_scopes"][i],
             gangs["priority"][i], gangs["anchor_node"][i], gangs["base_gang"][i], gangs["level"][i],
             gangs["flags"][i], gangs["preferred"][i]) = g
        for i, c in enumerate(self.cliques):
            (cliques["req_cpu_milli"][i], cliques["req_mem_mib"][i], cliques["req_gpu"][i],
             cliques["min_replicas"][i], cliques["replicas"][i], cliques["class_mask"][i],
             cliques["level"][i], cliques["scope"][i]) = c
        for i, s in enumerate(self.scopes):
            scopes["first_clique"][i], scopes["n_cliques"][i], scopes["level"][i], scopes["preferred1"][i] = s
        return gangs, cliques, scopes
