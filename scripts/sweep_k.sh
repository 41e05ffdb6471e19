# alternatives per gang (GROVE_TUNE_ALTERNATIVES) x first attempt window (GROVE_TUNE_WIDTH0)
for cfg in "8 24" "6 24" "6 12" "4 24" "4 8" "4 12" "3 8" "2 6"; do set -- $cfg
  GROVE_TUNE_ALTERNATIVES=$1 GROVE_TUNE_WIDTH0=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('K=$1 width0=$2', round(d['ms_per_step'],3), 'rounds', d.get('rounds'), {a:round(b,3) for a,b in k.items()}, 'e2e', round(d['e2e']['ms_per_step'],2))"
done
