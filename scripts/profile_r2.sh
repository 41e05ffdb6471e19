#!/bin/bash
# round-2 profiles: launch list of one bench step, ncu --set full of K3 (k_eval), K2 (k_score), the publish kernel
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
    ms = e.build_score_matrix()
    print(st, ms)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python /tmp/c4one.py > gpurun_out/r2_launches.log 2>&1
# K3: the 4-warp form in round 6 (a throughput round) and the 8-warp form in round 16 (a latency round)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_eval --launch-skip 11 --launch-count 1 -o gpurun_out/r2_k_eval4 -f python /tmp/c4one.py > gpurun_out/r2_k_eval4.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_eval --launch-skip 30 --launch-count 1 -o gpurun_out/r2_k_eval8 -f python /tmp/c4one.py > gpurun_out/r2_k_eval8.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_score --launch-count 1 -o gpurun_out/r2_k_score -f python /tmp/c4one.py > gpurun_out/r2_k_score.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_apply --launch-skip 5 --launch-count 1 -o gpurun_out/r2_k_apply -f python /tmp/c4one.py > gpurun_out/r2_k_apply.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fit --launch-count 1 -o gpurun_out/r2_k_fit -f python /tmp/c4one.py > gpurun_out/r2_k_fit.log 2>&1
ls -la gpurun_out/*.ncu-rep
