for cfg in "1500 592" "1500 100" "1500 10"; do set -- $cfg
  GROVE_TUNE_WARP_MIN=$1 GROVE_TUNE_WIDE_MAX=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('warp_min=$1 wide_max=$2', round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items()}, round(d['e2e']['ms_per_step'],2))"
done
