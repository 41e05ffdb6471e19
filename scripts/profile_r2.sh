#!/bin/bash
# round-2 evidence: GPU tests, bench lines, launch list of the bench command, ncu --set full of the kernels of a round + K1 + K2, sanitizer
cd "$GRAFT_REPO_ROOT" || exit 1
export PYTHONPATH="$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/*.ncu-rep
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4 | tee $O/r2_gpu_tests.log; fi
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2_bench_C4.json 2> $O/r2_bench_C4.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
timeout 600 python bench.py --config C5 --steps 30 --warmup 5 > $O/r2_bench_C5.json 2> $O/r2_bench_C5.err
for c in C1 C2 C3; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 > $O/r2_bench_$c.json 2> $O/r2_bench_$c.err; done
python - <<'PY'
import json
for n in ("C4", "reference", "C5", "C1", "C2", "C3"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"], 3), "ms", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["ms_per_step"], 3) if "ms_per_step" in d["e2e"] else "", (d.get("cpu_baseline") or {}).get("outputs_identical_to_gpu"))
    except Exception as ex:
        print(n, "failed", ex)
PY
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_launches.log 2>&1
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
if [ -n "$ONLY_BENCH" ]; then du -sh $O; exit 0; fi
cap() { timeout 600 ncu --set full --import-source on --clock-control none -k regex:$1 --launch-skip $2 --launch-count 1 -o $O/r2_$3 -f python /tmp/c4one.py > $O/r2_$3.log 2>&1; }
cap k_eval 13 k_eval_light     # round 7, a warp per gang
cap k_eval 18 k_eval_heavy     # round 10, eight warps per gang
cap k_apply 6 k_apply
cap k_detect 6 k_detect
cap k_fit 0 k_fit
cap k_shape_plaus 0 k_shape_plaus
cap k_score 0 k_score
for k in k_eval_light k_eval_heavy k_apply k_detect k_fit k_shape_plaus k_score; do
  ncu -i $O/r2_$k.ncu-rep --page details > $O/r2_ncu_${k}_details.txt 2>/dev/null
  ncu -i $O/r2_$k.ncu-rep --page raw --csv > $O/r2_ncu_${k}_raw.csv 2>/dev/null
done
python scripts/ncu_regions.py $O/r2_k_eval_light.ncu-rep admit.cuh 40 > $O/r2_ncu_k_eval_light_source_hotspots.txt 2>&1
python scripts/ncu_regions.py $O/r2_k_eval_heavy.ncu-rep admit.cuh 40 > $O/r2_ncu_k_eval_heavy_source_hotspots.txt 2>&1
python scripts/ncu_regions.py $O/r2_k_apply.ncu-rep relax.cuh 25 > $O/r2_ncu_k_apply_source_hotspots.txt 2>&1
python scripts/ncu_regions.py $O/r2_k_detect.ncu-rep relax.cuh 25 > $O/r2_ncu_k_detect_source_hotspots.txt 2>&1
ls -la $O/r2_*.ncu-rep | awk '{print $5, $9}'
rm -f $O/r2_*.ncu-rep    # (gpurun brings back at most 64 MiB: the text / csv exports above are what is kept)
du -sh $O
