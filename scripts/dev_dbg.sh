#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export PYTHONPATH="$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
cat > /tmp/c4one.py <<'PY'
import numpy as np, os
from grove_b200 import synth
from grove_b200.engine import PlacementEngine
cfg = synth.config_c4()
g, c, s = cfg["tables"]
with PlacementEngine(cfg["n_levels"]) as e:
    e.load_nodes(cfg["nodes"]); e.submit_gangs(g, c, s)
    st = e.run_cycle()
    e.load_nodes(cfg["nodes"]); st = e.run_cycle()
    print(st)
PY
GROVE_DEBUG_ADMIT=1 timeout 300 python /tmp/c4one.py 2>&1 | tee gpurun_out/dev_dbg.log | grep "warp 0\|cycle:\|packed evals" | cut -c1-330
