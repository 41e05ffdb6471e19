"""ctypes binding of libgrove_place.so -- the call a scheduler backend makes.

Mirrors the C ABI of include/grove_place.h one to one (the cgo stub in INTEGRATION.md binds the same
symbols).  Nothing here computes placements and nothing here falls back to a CPU path: if the
library is missing or no CUDA device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import tables as T

ABI_VERSION = 2
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgrove_place.so")

ERRORS = {0: "OK", -1: "INVALID_ARG", -2: "NO_DEVICE", -3: "CUDA", -4: "LIMIT", -5: "STATE", -6: "OOM"}

# every symbol include/grove_place.h declares
SYMBOLS = [
    "grove_abi_version", "grove_engine_create", "grove_engine_destroy", "grove_last_error",
    "grove_load_nodes", "grove_update_nodes", "grove_get_nodes", "grove_submit_gangs", "grove_run_cycle",
    "grove_get_placements", "grove_get_gang_status", "grove_get_scope_domains", "grove_load_nodes_device",
    "grove_build_score_matrix", "grove_run_cycle_preempt", "grove_get_victims",
    "grove_shard_range", "grove_run_score_pass", "grove_shard_summary_device",
    "grove_debug_get_perm", "grove_debug_get_fit_row", "grove_debug_get_score_row",
]


class GroveError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen the engine.  Raises (never falls back) when the shared library has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get("GROVE_PLACE_LIB", path)  # experiments: another build of the same sources
        if not os.path.exists(path):
            raise GroveError(-2, f"{path} not built; run `python -m grove_b200.build` (needs nvcc)")
        lib = C.CDLL(path)
        for s in SYMBOLS:
            getattr(lib, s).restype = C.c_int32
        lib.grove_abi_version.restype = C.c_uint32
        lib.grove_last_error.restype = C.c_char_p
        lib.grove_engine_destroy.restype = None
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PlacementEngine:
    """One scheduler session on one GPU (not thread-safe; one cycle in flight)."""

    def __init__(self, n_levels: int, device: int = 0, window: int = 0, rank: int = 0, world: int = 1):
        self.lib = load_library()
        cfg = np.zeros(1, dtype=T.config_dt)
        cfg["abi_version"] = ABI_VERSION
        cfg["device"], cfg["n_levels"], cfg["window"] = device, n_levels, window
        cfg["rank"], cfg["world"] = rank, world
        self.h = C.c_void_p()
        rc = self.lib.grove_engine_create(_p(cfg), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise GroveError(rc, "grove_engine_create failed (no CPU fallback exists)")
        self.n_levels = n_levels
        self.n = self.G = self.Q = self.S = 0
        self._cap_pods = 0
        self._pl_buf = self._st_buf = None

    def _check(self, rc: int):
        if rc != 0:
            raise GroveError(rc, (self.lib.grove_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.grove_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- inputs ----
    def load_nodes(self, nodes: np.ndarray):
        nodes = np.ascontiguousarray(nodes, dtype=T.node_dt)
        self._check(self.lib.grove_load_nodes(self.h, _p(nodes), C.c_uint32(len(nodes))))
        self.n = len(nodes)

    def load_nodes_device(self, dev_ptr: int, n: int):
        self._check(self.lib.grove_load_nodes_device(self.h, C.c_void_p(dev_ptr), C.c_uint32(n)))

    def update_nodes(self, idx: np.ndarray, recs: np.ndarray):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        recs = np.ascontiguousarray(recs, dtype=T.node_dt)
        self._check(self.lib.grove_update_nodes(self.h, _p(idx), _p(recs), C.c_uint32(len(idx))))

    def submit_gangs(self, gangs: np.ndarray, cliques: np.ndarray, scopes: np.ndarray):
        gangs = np.ascontiguousarray(gangs, dtype=T.gang_dt)
        cliques = np.ascontiguousarray(cliques, dtype=T.clique_dt)
        scopes = np.ascontiguousarray(scopes, dtype=T.scope_dt)
        self._check(self.lib.grove_submit_gangs(self.h, _p(gangs), C.c_uint32(len(gangs)), _p(cliques),
                                                C.c_uint32(len(cliques)), _p(scopes), C.c_uint32(len(scopes))))
        self.G, self.Q, self.S = len(gangs), len(cliques), len(scopes)
        self._cap_pods = int(cliques["replicas"].astype(np.int64).sum()) if len(cliques) else 0

    # ---- the cycle ----
    def run_cycle(self) -> dict:
        st = np.zeros(1, dtype=T.stats_dt)
        self._check(self.lib.grove_run_cycle(self.h, _p(st)))
        return {k: st[k][0].item() for k in T.stats_dt.names}

    def run_cycle_preempt(self, running: np.ndarray, holdings: np.ndarray) -> dict:
        """ordinary pass + reclaim pass: `running` (T.running_dt) are the gangs admitted by earlier cycles, `holdings`
        (T.holding_dt) what each holds per node; rejected gangs may evict running gangs of a lower priority"""
        running = np.ascontiguousarray(running, dtype=T.running_dt)
        holdings = np.ascontiguousarray(holdings, dtype=T.holding_dt)
        st = np.zeros(1, dtype=T.stats_dt)
        self._check(self.lib.grove_run_cycle_preempt(self.h, _p(running), C.c_uint32(len(running)), _p(holdings),
                                                     C.c_uint32(len(holdings)), _p(st)))
        return {k: st[k][0].item() for k in T.stats_dt.names}

    def victims(self) -> np.ndarray:
        """(running gang, preemptor gang) pairs of the last run_cycle_preempt: the caller sets DisruptionTarget on them"""
        n = C.c_uint32(0)
        self._check(self.lib.grove_get_victims(self.h, None, C.c_uint32(0), C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=T.victim_dt)
        self._check(self.lib.grove_get_victims(self.h, _p(out), C.c_uint32(len(out)), C.byref(n)))
        return out[: n.value]

    # ---- outputs ----
    def placements(self, copy: bool = True) -> np.ndarray:
        """copy=False returns a view into a buffer the wrapper reuses (valid until the next call)."""
        cap = max(self._cap_pods, 1)
        if self._pl_buf is None or len(self._pl_buf) < cap:
            self._pl_buf = np.empty(cap, dtype=T.placement_dt)
        n = C.c_uint32(0)
        self._check(self.lib.grove_get_placements(self.h, _p(self._pl_buf), C.c_uint32(len(self._pl_buf)), C.byref(n)))
        out = self._pl_buf[: n.value]
        return out.copy() if copy else out

    def gang_status(self, copy: bool = True) -> np.ndarray:
        cap = max(self.G, 1)
        if self._st_buf is None or len(self._st_buf) < cap:
            self._st_buf = np.empty(cap, dtype=T.status_dt)
        self._check(self.lib.grove_get_gang_status(self.h, _p(self._st_buf), C.c_uint32(len(self._st_buf))))
        out = self._st_buf[: self.G]
        return out.copy() if copy else out

    def scope_domains(self) -> np.ndarray:
        """chosen topology domain of every scope (TopologyConstraintGroupConfig) of the submission"""
        out = np.zeros(max(self.S, 1), dtype=T.scope_status_dt)
        self._check(self.lib.grove_get_scope_domains(self.h, _p(out), C.c_uint32(len(out))))
        return out[: self.S]

    def nodes(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=T.node_dt)
        self._check(self.lib.grove_get_nodes(self.h, _p(out), C.c_uint32(self.n)))
        return out

    def build_score_matrix(self) -> float:
        """K2 over the last cycle's start snapshot; returns the kernel's device time in ms"""
        ms = C.c_float(0)
        self._check(self.lib.grove_build_score_matrix(self.h, C.byref(ms)))
        return ms.value

    # ---- multi-GPU score pass (node-range shards; grove_b200/sharded.py drives it) ----
    def shard_range(self) -> tuple[int, int]:
        """topology-sorted node range [lo, hi) this handle builds K1 / K2 for (the whole table when world <= 1)"""
        lo, hi = C.c_uint32(0), C.c_uint32(0)
        self._check(self.lib.grove_shard_range(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def run_score_pass(self) -> float:
        """K1 + K2 over the loaded snapshot for this handle's node range, no admission; returns device ms"""
        ms = C.c_float(0)
        self._check(self.lib.grove_run_score_pass(self.h, C.byref(ms)))
        return ms.value

    def shard_summary_into(self, dev_ptr: int, cap_words: int):
        """int32[G + Q] feasibility / capacity counts of this shard into a DEVICE buffer (all-reduced by the caller)"""
        self._check(self.lib.grove_shard_summary_device(self.h, C.c_void_p(dev_ptr), C.c_uint32(cap_words)))

    # ---- introspection (parity tests) ----
    def debug_perm(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint32)
        self._check(self.lib.grove_debug_get_perm(self.h, _p(out), C.c_uint32(self.n)))
        return out

    def debug_fit_row(self, clique: int) -> np.ndarray:
        out = np.zeros((self.n + 31) // 32, dtype=np.uint32)
        self._check(self.lib.grove_debug_get_fit_row(self.h, C.c_uint32(clique), _p(out), C.c_uint32(len(out))))
        return out

    def debug_score_row(self, clique: int) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint8)
        self._check(self.lib.grove_debug_get_score_row(self.h, C.c_uint32(clique), _p(out), C.c_uint32(self.n)))
        return out
