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
GROVE_TUNE_AHEAD=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 700 --csv --log-file gpurun_out/s2_launches.csv python /tmp/c4one.py > gpurun_out/s2_launches.log 2>&1
tail -2 gpurun_out/s2_launches.log | cut -c1-300
