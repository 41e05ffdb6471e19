python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for w0 in 16; do
  GROVE_TUNE_WIDTH0=$w0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('w0=$w0', round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items()}, round(d['e2e']['ms_per_step'],2))"
done
