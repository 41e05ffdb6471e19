#!/bin/bash
# session baseline: all GPU tests, the bench line, the launch list of one cycle, per-round debug statistics
cd "$GRAFT_REPO_ROOT" || exit 1
export PYTHONPATH="$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/s2_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; tail -c 3000 gpurun_out/s2_bench.json
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s2_launches.csv python /tmp/c4one.py > gpurun_out/s2_launches.log 2>&1
GROVE_DEBUG_ADMIT=1 timeout 300 python /tmp/c4one.py > gpurun_out/s2_dbg.log 2>&1
tail -3 gpurun_out/s2_dbg.log | cut -c1-400
