# compute-sanitizer over a small cycle (memcheck + racecheck), on the GPU box
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
from grove_b200 import synth
from grove_b200.engine import PlacementEngine
for cfg in (synth.config_c3(n=756, g=120), synth.config_c2(n=200, g=30), synth.config_c4(n=2520, g=800)):
    g, c, s = cfg["tables"]
    with PlacementEngine(cfg["n_levels"]) as e:
        e.load_nodes(cfg["nodes"]); e.submit_gangs(g, c, s); st = e.run_cycle(); print(st["rounds"], st["gangs_admitted"])
        import numpy as np
        idx = np.arange(0, len(cfg["nodes"]), 7, dtype=np.uint32)   # churn path: update a few nodes, read the table back
        e.update_nodes(idx, cfg["nodes"][idx]); e.nodes(); e.submit_gangs(g, c, s); e.run_cycle()
sys.path.insert(0, 'tests')
from test_random_parity_gpu import random_case
for seed in (3001, 3007, 3012):   # Preferred levels at gang / scope / clique, incl. the whole-cluster fallback
    nodes, L, (g, c, s) = random_case(seed, pref=True)
    with PlacementEngine(L) as e:
        e.load_nodes(nodes); e.submit_gangs(g, c, s); st = e.run_cycle(); print("pref", st["rounds"], st["gangs_admitted"])
# the reclaim pass and the sharded score pass
sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import oracle_py as O
from preempt_cases import churned_cluster
from grove_b200.sharded import engine_summary
nodes, L, (g, c, s), running, holdings = churned_cluster(O, 11, n=1260, g_running=200, g_pending=100)
with PlacementEngine(L) as e:
    e.load_nodes(nodes); e.submit_gangs(g, c, s); st = e.run_cycle_preempt(running, holdings); print("preempt", st["gangs_admitted"], len(e.victims()))
for r in range(2):
    with PlacementEngine(L, rank=r, world=2) as e:
        e.load_nodes(nodes); e.submit_gangs(g, c, s); e.run_score_pass(); t = engine_summary(e, torch.device("cuda", 0))(r); print("shard", r, int(t.sum()))
PY
cd "${GRAFT_REPO_ROOT:-.}"; export PYTHONPATH="${GRAFT_REPO_ROOT:-.}"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -2 gpurun_out/sanitizer_$tool.log
done
