#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_random_parity_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/dev_tests.log
GROVE_DEBUG_ADMIT=1 timeout 300 python - <<'PY' 2>&1 | tail -60 | tee gpurun_out/dev_c4.log
import numpy as np, time
from grove_b200 import synth
from grove_b200.engine import PlacementEngine
cfg = synth.config_c4()
g, c, s = cfg["tables"]
with PlacementEngine(cfg["n_levels"]) as e:
    e.load_nodes(cfg["nodes"]); e.submit_gangs(g, c, s)
    for i in range(2):
        e.load_nodes(cfg["nodes"])
        st = e.run_cycle()
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}, flush=True)
PY
