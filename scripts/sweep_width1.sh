# attempt-window sweep of the CTA-per-gang admission kernels (GROVE_TUNE_WIDTH1 = attempts per warp in the first window)
python -m pytest tests/test_parity_gpu.py tests/test_random_parity_gpu.py -x -q -m gpu 2>&1 | tail -2
for w in 32 1 2 4 8; do
  GROVE_TUNE_WIDTH1=$w python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('width1=$w', round(d['ms_per_step'],3), {a:round(b,3) for a,b in k.items()}, round(d['e2e']['ms_per_step'],2))"
done
