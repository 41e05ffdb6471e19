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
    print(st)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_dev.csv python /tmp/c4one.py > gpurun_out/ncu_dev.log 2>&1
tail -3 gpurun_out/ncu_dev.log
