#!/bin/bash
# ncu --set full of one light (round 6) and one heavy (round 9) k_eval launch, source hot spots
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
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_eval --launch-skip ${SKIP_LIGHT:-11} --launch-count 1 -o gpurun_out/s2_eval_light -f python /tmp/c4one.py > gpurun_out/s2_ncu_light.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_eval --launch-skip ${SKIP_HEAVY:-16} --launch-count 1 -o gpurun_out/s2_eval_heavy -f python /tmp/c4one.py > gpurun_out/s2_ncu_heavy.log 2>&1
python scripts/ncu_hotspots.py gpurun_out/s2_eval_light.ncu-rep regex:k_eval > gpurun_out/s2_hot_light.txt 2>&1
python scripts/ncu_hotspots.py gpurun_out/s2_eval_heavy.ncu-rep regex:k_eval > gpurun_out/s2_hot_heavy.txt 2>&1
head -40 gpurun_out/s2_hot_light.txt
