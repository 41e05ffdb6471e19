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
    for i in range(4):
        e.load_nodes(cfg["nodes"])
        st = e.run_cycle()
    print(os.environ.get("TAG", ""), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if k in ("rounds", "evaluations", "gangs_admitted", "ms_fit", "ms_score", "ms_admit", "ms_total")}, flush=True)
    if ref is not None:
        print("identical:", np.array_equal(e.placements(), ref["placements"]), np.array_equal(e.gang_status(), ref["status"]), np.array_equal(e.scope_domains(), ref["scope_status"]), np.array_equal(e.nodes(), ref["nodes_after"]), flush=True)
PY
(
CHECK=1 TAG=default timeout 300 python /tmp/c4run.py
TAG="w4=0 (1 warp always)" GROVE_TUNE_WARP4=0 GROVE_TUNE_WARP16=0 timeout 120 python /tmp/c4run.py
TAG="w4=100000 w16=0 (4 warps always)" GROVE_TUNE_WARP4=100000 GROVE_TUNE_WARP16=0 timeout 120 python /tmp/c4run.py
TAG="8 warps always" GROVE_TUNE_WARP4=100000 GROVE_TUNE_WARP16=100000 timeout 120 python /tmp/c4run.py
TAG="w16=1500" GROVE_TUNE_WARP16=1500 timeout 120 python /tmp/c4run.py
TAG="entry=2048 w16=1500" GROVE_TUNE_ENTRY=2048 GROVE_TUNE_WARP16=1500 timeout 120 python /tmp/c4run.py
TAG="entry=2048 8 warps" GROVE_TUNE_ENTRY=2048 GROVE_TUNE_WARP4=100000 GROVE_TUNE_WARP16=100000 timeout 120 python /tmp/c4run.py
TAG="entry=512 8 warps" GROVE_TUNE_ENTRY=512 GROVE_TUNE_WARP4=100000 GROVE_TUNE_WARP16=100000 timeout 120 python /tmp/c4run.py
) 2>&1 | tee gpurun_out/dev_sweep2.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_random_parity_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/dev_tests.log
