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
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -3 gpurun_out/sanitizer_$tool.log
done
