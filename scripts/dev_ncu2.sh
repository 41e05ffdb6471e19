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
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_eval --launch-skip 22 --launch-count 1 -o gpurun_out/k_eval_r8 -f python /tmp/c4one.py > gpurun_out/ncu_eval.log 2>&1
tail -3 gpurun_out/ncu_eval.log
ls -la gpurun_out/*.ncu-rep
