#!/usr/bin/env python
"""bench.py -- PodGang placements/sec on the 50k-node / 10k-gang synthetic snapshot (BASELINE.json).

A "step" is one scheduling cycle: every pending PodGang of the snapshot goes through
fit -> score -> admit -> commit (optimistic rounds) until it is admitted or rejected.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4] [--impl reference]

`value`       whole-job admitted gangs / second, node table already resident in HBM (per step:
              device->device reset of the node table, then the cycle).
`e2e`         the same metric through the C ABI with HOST buffers every step: grove_load_nodes +
              grove_submit_gangs + grove_run_cycle + grove_get_placements + grove_get_gang_status.
`roofline`    the dominant kernel (K2 score matrix): algorithmic bytes / CUDA-event time, against the
              measured HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline`/--impl reference   the CPU oracle (oracle/, a C restatement: the reference tree holds no
              scheduler and there is no Go toolchain) on this box's host cores, same snapshot.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from grove_b200 import synth, tables as T  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def ncu_traffic():
    """DRAM bytes (read + write) of the captured k_score launch (round 1: 12 500 rows), from the committed
    `ncu --set full` capture; None when the capture is absent."""
    import csv
    p = os.path.join(ROOT, "profiles", "r1_ncu_k_score_raw.csv")
    try:
        rows = list(csv.reader(open(p)))
        h, units, vals = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = h.index(name)
            tot += float(vals[i].replace(",", "")) * scale[units[i]]
        return tot
    except (OSError, ValueError, KeyError, IndexError):
        return None


def make_workload(name: str):
    cfg = synth.CONFIGS[name]()
    g, c, s = cfg["tables"]
    return cfg["nodes"], cfg["n_levels"], g, c, s


def workload_desc(name, nodes, g, c, dev_note):
    return {"workload": f"{name}: {len(nodes)} nodes / {len(g)} PodGangs / {len(c)} PodCliques, "
                        + {"C4": "4-level tree zone/block/rack/host, hierarchical PCSG gangs (base + scaled)",
                           "C3": "3-level tree, prefill+decode cliques", "C2": "flat", "C1": "simple1.yaml"}[name],
            "nodes": int(len(nodes)), "gangs": int(len(g)), "cliques": int(len(c)),
            "pairs_per_full_pass": int(len(nodes)) * int(len(c)),
            "l2": dev_note}


def run_reference(args):
    """--impl reference: the CPU oracle on all host cores.  A step is one cycle over the full workload; if K such
    steps would not fit a few minutes, every step runs a bounded sample instead (the first G' PodGangs of the same
    snapshot that are self-contained, against all nodes), and the line says so."""
    from oracle import oracle_py as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nodes, L, g, c, s = make_workload(args.config)
    cores = os.cpu_count() or 1
    O.build()
    t0 = time.perf_counter()
    O.run_cycle(nodes, L, g, c, s, threads=cores)  # probe (also warms the page cache)
    t_full = time.perf_counter() - t0
    budget = 150.0
    sample = f"full {args.config} cycle per step"
    total_steps = max(1, args.steps + args.warmup)
    if t_full * total_steps > budget and len(g) > 1000:
        keep = max(500, int(len(g) * budget / (t_full * total_steps)))
        # base gangs first: a prefix of the table keeps every base_gang reference inside the sample
        while keep < len(g) and g["base_gang"][:keep].max(initial=0) != T.NONE_U32 and \
                (g["base_gang"][:keep][g["base_gang"][:keep] != T.NONE_U32] >= keep).any():
            keep += 1
        gs = g[:keep].copy()
        nc = int(gs["clique_off"][-1] + gs["n_cliques"][-1]); ns = int(gs["scope_off"][-1] + gs["n_scopes"][-1])
        g, c, s = gs, c[:nc], s[:ns]
        sample = f"bounded sample: first {keep} PodGangs of {args.config} against all {len(nodes)} nodes per step"
    for _ in range(args.warmup):
        O.run_cycle(nodes, L, g, c, s, threads=cores)
    t0 = time.perf_counter()
    adm = 0
    for _ in range(args.steps):
        r = O.run_cycle(nodes, L, g, c, s, threads=cores)
        adm = r["stats"]["gangs_admitted"]
    dt = (time.perf_counter() - t0) / args.steps
    val = adm / dt
    line = {
        "impl": "reference", "metric": "podgang_placements_per_sec", "value": val, "unit": "gangs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_desc(args.config, nodes, g, c, "n/a (CPU)"),
        "cpu_baseline": {"value": val, "unit": "gangs/s", "cores": cores, "kind": "port",
                         "sample": sample + "; OpenMP over gangs; C restatement (no scheduler in the reference tree, "
                                            "Go toolchain absent)"},
        "e2e": {"value": val, "unit": "gangs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_gpu(args):
    import torch
    from grove_b200 import build
    from grove_b200.engine import PlacementEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU oracle)")
    build.build()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    nodes, L, g, c, s = make_workload(args.config)
    eng = PlacementEngine(L, device=local, rank=rank, world=world)
    eng.load_nodes(nodes)
    eng.submit_gangs(g, c, s)
    d_nodes = torch.from_numpy(nodes.view(np.uint8).reshape(-1)).cuda()  # pristine snapshot, resident in HBM
    h2d = nodes.nbytes + g.nbytes + c.nbytes + s.nbytes

    if world > 1:
        from grove_b200.sharded import run_sharded_cycle
        def step_dev():
            eng.load_nodes_device(d_nodes.data_ptr(), len(nodes))
            return run_sharded_cycle(eng, dist)
        def step_e2e():
            eng.load_nodes(nodes); eng.submit_gangs(g, c, s)
            st = run_sharded_cycle(eng, dist)
            return st, eng.placements(copy=False), eng.gang_status(copy=False)
    else:
        def step_dev():
            eng.load_nodes_device(d_nodes.data_ptr(), len(nodes))
            return eng.run_cycle()
        def step_e2e():
            eng.load_nodes(nodes); eng.submit_gangs(g, c, s)
            st = eng.run_cycle()
            return st, eng.placements(copy=False), eng.gang_status(copy=False)

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # sampled from the warm-up on: a cycle is milliseconds, nvidia-smi samples every 20 ms
    for _ in range(args.warmup):
        step_dev()
    sync()
    t0 = time.perf_counter()
    acc = {k: 0.0 for k in ("ms_fit", "ms_score", "ms_admit", "ms_commit", "ms_total")}
    launches = 0
    for _ in range(args.steps):
        st = step_dev()
        for k in acc:
            acc[k] += st[k]
        launches += st["kernel_launches"] + 1  # + k_gather of the node-table reset
    sync()
    dt = time.perf_counter() - t0
    # e2e through the C ABI with host buffers
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    sync()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        st_e, pl, gs = step_e2e()
    sync()
    dt_e = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None  # covers warm-up, the timed steps and the e2e steps
    if dist is not None:
        tt = torch.tensor([dt, dt_e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e = tt.tolist()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    adm = st["gangs_admitted"]
    resolved = st["gangs_admitted"] + st["gangs_rejected"]
    ms_step = dt / args.steps * 1e3
    ms_e2e = dt_e / args.steps * 1e3
    peak, peak_src = peaks()
    # K2 (score matrix): per (clique,node) pair it reads 1 fit bit and writes 1 score byte.  In the product
    # path K2 runs on a second stream BESIDE the admission kernel, so its CUDA-event duration there includes
    # the SMs it yields to K3; its stand-alone duration is timed in a second pass of the same K steps on an
    # engine created with the overlap switched off (GROVE_TUNE_OVERLAP=0), same kernels, same inputs.
    pairs = st["pairs_evaluated"]
    k2_bytes = pairs * (1.0 + 1.0 / 8.0)
    k2_ms_overlapped = acc["ms_score"] / args.steps
    k2_ms = k2_ms_overlapped
    if world == 1 and rank == 0:
        os.environ["GROVE_TUNE_OVERLAP"] = "0"
        try:
            eng2 = PlacementEngine(L, device=local)
            eng2.load_nodes(nodes); eng2.submit_gangs(g, c, s)
            tot = 0.0
            for i in range(args.warmup + args.steps):
                eng2.load_nodes_device(d_nodes.data_ptr(), len(nodes))
                s2 = eng2.run_cycle()
                if i >= args.warmup:
                    tot += s2["ms_score"]
            k2_ms = tot / args.steps
            eng2.close()
        finally:
            del os.environ["GROVE_TUNE_OVERLAP"]
    achieved = k2_bytes / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else 0.0
    d2h = pl.nbytes + gs.nbytes
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle_py as O
        cores = os.cpu_count() or 1
        O.build()
        tc = time.perf_counter()
        r = O.run_cycle(nodes, L, g, c, s, threads=cores)
        tcpu = time.perf_counter() - tc
        same = bool(np.array_equal(r["placements"], pl) and np.array_equal(r["status"]["state"], gs["state"]))
        cpu = {"value": r["stats"]["gangs_admitted"] / tcpu, "unit": "gangs/s", "cores": cores, "kind": "port",
               "sample": f"one full {args.config} cycle ({r['stats']['pairs_evaluated']} pairs, {r['stats']['rounds']} rounds, {tcpu:.2f} s); "
                         "C restatement timed on this box (reference tree has no scheduler; Go toolchain absent)",
               "placements_identical_to_gpu": same}
    line = {
        "metric": "podgang_placements_per_sec", "value": adm / (ms_step * 1e-3), "unit": "gangs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {**workload_desc(args.config, nodes, g, c, "fit+score matrices (%.2f GB) exceed the 126 MB L2; no flush needed"
                                   % ((len(c) * eng_npad(len(nodes)) * 1.125) / 1e9)),
                   "gangs_resolved_per_sec": resolved / (ms_step * 1e-3), "rounds": st["rounds"],
                   "admitted": adm, "rejected": st["gangs_rejected"], "pods_bound": st["pods_bound"],
                   "parallelism": "gang rows sharded over %d GPU(s)" % world},
        "clocks": clocks,
        "e2e": {"value": adm / (ms_e2e * 1e-3), "unit": "gangs/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_score (K2 topology-distance score matrix)", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic() if args.config == "C4" else None,
                     "traffic_note": "DRAM read+write of the round-1 k_score launch (12 500 rows x 50 176 B; algorithmic 703 MB) from "
                                     "profiles/r1_ncu_k_score_raw.csv; write-only ceiling on this box = 3.93 TB/s (torch memset), "
                                     "i.e. 0.60 of the copy peak used as denominator",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_step": k2_bytes, "ms_per_step": k2_ms,
                     "ms_per_step_overlapped_with_admit": k2_ms_overlapped,
                     "how": "CUDA events around every k_score launch on its launching stream; stand-alone pass with "
                            "GROVE_TUNE_OVERLAP=0 (product path overlaps k_score with k_admit on two streams)"},
        "kernel_ms_per_step": {k: v / args.steps for k, v in acc.items()},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_churn(args):
    """--config C5 (BASELINE.json config 5, steady-state churn): the C4 cluster, 100 PodGang arrivals per 100 ms tick
    (1 000 /s) joining the gangs still pending, ~1 % of the running gangs finishing per tick (their resources come back
    through grove_update_nodes).  A step is one tick: update_nodes + submit_gangs + run_cycle + results to host buffers,
    all through the C ABI with host arrays (e2e); `value` counts the cycle alone (tables already resident)."""
    import torch
    from grove_b200 import build
    from grove_b200.engine import PlacementEngine
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("bench.py --config C5 is a single-GPU line (the churn host loop is sequential)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    build.build()
    ch = synth.ChurnC5(n=50000, arrivals=100, release_pct=1)
    sampler = ClockSampler(0)
    sampler.start()
    t_tick, t_cycle, adm, launches, pend, h2d, d2h = [], [], 0, 0, [], 0, 0
    cpu_t, cpu_adm, same, kept = 0.0, 0, True, []
    idx = recs = None
    with PlacementEngine(ch.n_levels) as e:
        e.load_nodes(ch.nodes)
        for tick in range(args.warmup + args.steps):
            specs, tabs = ch.begin_tick()
            g, c, s = tabs
            before = ch.nodes
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if idx is not None and len(idx):
                e.update_nodes(idx, recs)
            e.submit_gangs(g, c, s)
            st = e.run_cycle()
            pl, gs = e.placements(), e.gang_status()
            dt = time.perf_counter() - t0
            if tick >= args.warmup:
                t_tick.append(dt * 1e3); t_cycle.append(st["ms_total"]); adm += st["gangs_admitted"]
                launches += st["kernel_launches"]; pend.append(len(g))
                h2d += g.nbytes + c.nbytes + s.nbytes + (0 if idx is None else idx.nbytes + recs.nbytes); d2h += pl.nbytes + gs.nbytes
            if not args.no_cpu_baseline and args.warmup <= tick < args.warmup + 5:   # bounded CPU sample: five ticks, replayed below
                kept.append((before.copy(), g, c, s, pl.copy(), gs.copy()))
            idx, recs = ch.end_tick(specs, tabs, gs, pl, e.nodes())
    # the CPU leg runs AFTER the timed ticks: the oracle's OpenMP team (all cores) and the engine's (8 threads) share one
    # libgomp, and alternating team sizes makes it re-create its threads -- tens of milliseconds that belong to neither
    for before, g, c, s, pl, gs in kept:
        from oracle import oracle_py as O
        tc = time.perf_counter()
        r = O.run_cycle(before, ch.n_levels, g, c, s, threads=os.cpu_count() or 1)
        cpu_t += time.perf_counter() - tc; cpu_adm += r["stats"]["gangs_admitted"]
        same &= bool(np.array_equal(r["placements"], pl) and np.array_equal(r["status"], gs))
    clocks = sampler.stop()
    ms_tick, ms_cyc = float(np.mean(t_tick)), float(np.mean(t_cycle))
    line = {
        "metric": "podgang_placements_per_sec", "value": adm / (sum(t_cycle) * 1e-3), "unit": "gangs/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_cyc, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "C5: steady-state churn on the C4 cluster (50000 nodes, 4-level tree): 100 PodGang arrivals per 100 ms tick "
                               "+ carried-over pending gangs, ~1 % of running gangs finish per tick (grove_update_nodes)",
                   "nodes": 50000, "arrivals_per_tick": 100, "tick_budget_ms": 100.0, "pending_per_tick_mean": float(np.mean(pend)),
                   "pending_per_tick_last": int(pend[-1]), "admitted": int(adm), "tick_ms_mean": ms_tick, "tick_ms_max": float(np.max(t_tick)),
                   "tick_budget_used": ms_tick / 100.0, "arrivals_per_sec_sustained_at_this_latency": 100.0 / (ms_tick * 1e-3),
                   "l2": "per tick the score matrix is cliques x 50176 B (> 126 MB L2 from ~2500 pending cliques on); inputs change every tick"},
        "clocks": clocks,
        "e2e": {"value": adm / (sum(t_tick) * 1e-3), "unit": "gangs/s", "ms_per_step": ms_tick,
                "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": int(d2h / args.steps)},
        "gpu_launches": int(launches),
        "roofline": None,
        "cpu_baseline": None if args.no_cpu_baseline else {
            "value": cpu_adm / cpu_t if cpu_t else 0.0, "unit": "gangs/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "the first five timed ticks (same inputs, C restatement on this box's host cores)", "placements_identical_to_gpu": same},
    }
    print(json.dumps(line))


def eng_npad(n):
    return (n + 1023) // 1024 * 1024


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C4", choices=sorted(synth.CONFIGS) + ["C5"])
    ap.add_argument("--impl", default="grove_b200", choices=["grove_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        if args.config == "C5":
            raise SystemExit("bench.py: the C5 line times the CPU oracle itself (cpu_baseline, five ticks); --impl reference covers C1-C4")
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        if args.config == "C5":
            run_churn(args)
        else:
            run_gpu(args)


if __name__ == "__main__":
    main()
