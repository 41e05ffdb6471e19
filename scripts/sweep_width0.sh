# first attempt window of the warp-per-gang admission kernel (GROVE_TUNE_WIDTH0)
for w in 24 8 10 12 16 32; do
  GROVE_TUNE_WIDTH0=$w python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('width0=$w', round(d['ms_per_step'],3), {a:round(b,3) for a,b in k.items()}, 'e2e', round(d['e2e']['ms_per_step'],2))"
done
