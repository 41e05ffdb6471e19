# experiment: alternatives per gang (K) and sub-rounds per round, built as variants of the same sources
# (nvcc ... -DGROVE_MAX_ALTERNATIVES=Ku -DGROVE_SUBROUNDS=Su -o grove_b200/variants/kK_sS.so)
run() {
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$1', round(d['ms_per_step'],3), 'rounds', d['config'].get('rounds', d.get('rounds')), {a:round(b,3) for a,b in k.items()}, 'e2e', round(d['e2e']['ms_per_step'],2), 'admitted', d.get('gangs_admitted', d['config'].get('gangs_admitted')))"
}
run base
for v in k8_s16 k16_s8 k16_s16 k32_s32; do
  GROVE_PLACE_LIB=$PWD/grove_b200/variants/$v.so GROVE_TUNE_WIDTH0=32 run $v
done
GROVE_PLACE_LIB=$PWD/grove_b200/variants/k16_s16.so GROVE_TUNE_WIDTH0=24 run k16_s16_w24
GROVE_PLACE_LIB=$PWD/grove_b200/variants/k32_s32.so GROVE_DEBUG_ADMIT=1 python bench.py --steps 1 --warmup 0 --no-cpu-baseline 2>&1 | grep "^round" | head -8
