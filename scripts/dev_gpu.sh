#!/bin/bash
# first contact of the relaxation engine with a GPU: small configs with timeouts, then the test files
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/dev_small.log
import numpy as np, time, sys
from grove_b200 import synth
from grove_b200.engine import PlacementEngine
from oracle import oracle_py as O
for name, cfg in (("C1", synth.config_c1()), ("C2", synth.config_c2()), ("C3s", synth.config_c3(n=2016, g=200)), ("C4s", synth.config_c4(n=5040, g=1000)), ("C3", synth.config_c3())):
    g, c, s = cfg["tables"]
    ref = O.run_cycle(cfg["nodes"], cfg["n_levels"], g, c, s, threads=8)
    with PlacementEngine(cfg["n_levels"]) as e:
        e.load_nodes(cfg["nodes"]); e.submit_gangs(g, c, s)
        st = e.run_cycle()
        ok = (np.array_equal(e.placements(), ref["placements"]), np.array_equal(e.gang_status(), ref["status"]),
              np.array_equal(e.scope_domains(), ref["scope_status"]), np.array_equal(e.nodes(), ref["nodes_after"]))
        st2 = e.run_cycle() if False else None
    print(name, ok, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}, flush=True)
    if not all(ok):
        gs = e.gang_status() if False else None
PY
