for b in 1 2 4 8; do
  GROVE_TUNE_RESOLVE_BPS=$b python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('bps=$b', round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items()}, round(d['e2e']['ms_per_step'],2))"
done
