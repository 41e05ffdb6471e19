#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export PYTHONPATH="$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
cat > /tmp/c4run.py <<'PY'
import numpy as np, time, os, sys
from grove_b200 import synth
from grove_b200.engine import PlacementEngine
from oracle import oracle_py as O
cfg = synth.config_c4()
g, c, s = cfg["tables"]
ref = None
if os.environ.get("CHECK"):
    ref = O.run_cycle(cfg["nodes"], cfg["n_levels"], g, c, s, threads=16)
with PlacementEngine(cfg["n_levels"]) as e:
    e.load_nodes(cfg["nodes"]); e.submit_gangs(g, c, s)
    best = None
    for i in range(6):
        e.load_nodes(cfg["nodes"])
        st = e.run_cycle()
        if best is None or st["ms_total"] < best["ms_total"]: best = st
    st = best
    print(os.environ.get("TAG", ""), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if k in ("rounds", "evaluations", "ms_admit", "ms_total")}, flush=True)
    if ref is not None:
        print("identical:", np.array_equal(e.placements(), ref["placements"]), np.array_equal(e.gang_status(), ref["status"]), np.array_equal(e.scope_domains(), ref["scope_status"]), np.array_equal(e.nodes(), ref["nodes_after"]), flush=True)
PY
(
CHECK=1 TAG="noinline default" timeout 300 python /tmp/c4run.py
for w4 in 800 1200 1600; do for w16 in 300 600; do TAG="noinline w4=$w4 w8=$w16" GROVE_TUNE_WARP4=$w4 GROVE_TUNE_WARP16=$w16 timeout 120 python /tmp/c4run.py; done; done
export GROVE_PLACE_LIB="$GRAFT_REPO_ROOT/grove_b200/libgrove_place_inl.so"
CHECK=1 TAG="inline default" timeout 300 python /tmp/c4run.py
for w4 in 800 1200 1600; do for w16 in 300 600; do TAG="inline w4=$w4 w8=$w16" GROVE_TUNE_WARP4=$w4 GROVE_TUNE_WARP16=$w16 timeout 120 python /tmp/c4run.py; done; done
TAG="inline refresh=1024" GROVE_TUNE_REFRESH=1024 timeout 120 python /tmp/c4run.py
TAG="inline refresh=2048" GROVE_TUNE_REFRESH=2048 timeout 120 python /tmp/c4run.py
TAG="inline noscore" GROVE_TUNE_SCORE=0 timeout 120 python /tmp/c4run.py
) 2>&1 | tee gpurun_out/dev_sweep4.log
