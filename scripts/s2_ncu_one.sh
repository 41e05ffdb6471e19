#!/bin/bash
# ncu --set full of one launch of kernel $K (regex), skipping $SKIP launches
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
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip ${SKIP:-0} --launch-count 1 -o gpurun_out/s2_one_$K -f python /tmp/c4one.py > gpurun_out/s2_one_$K.log 2>&1
ncu -i gpurun_out/s2_one_$K.ncu-rep --page details 2>/dev/null | grep -E "Duration|Executed Ipc Active|Registers Per|Achieved Occupancy|Grid Size|Block Size|Executed Instructions  |Warp Cycles Per Issued|Theoretical Occ|L2 Hit|Local" | head -14
python scripts/ncu_regions.py gpurun_out/s2_one_$K.ncu-rep ${FILE:-admit.cuh} 16
