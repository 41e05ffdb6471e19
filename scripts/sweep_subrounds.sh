# experiment: conflict-resolution sub-rounds per round (variants built with -DGROVE_SUBROUNDS=Nu)
run() {
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$1', round(d['ms_per_step'],3), 'rounds', d['config']['rounds'], 'admitted', d['config']['admitted'], {a:round(b,3) for a,b in k.items()}, 'e2e', round(d['e2e']['ms_per_step'],2))"
}
run s8
for v in s6 s5 s4; do GROVE_PLACE_LIB=$PWD/grove_b200/variants/$v.so run $v; done
