set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print({k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step')}, d['e2e'], d['roofline']['frac'], d['cpu_baseline'])
PY
