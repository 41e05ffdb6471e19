# experiment: CTAs per SM of the 8-warp admission kernel (main build: 2, variants/wide1.so: 1) x the gang count
# from which the 4-warp kernel takes over (GROVE_TUNE_WIDE_MAX)
run() {
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$1', round(d['ms_per_step'],3), {a:round(b,3) for a,b in k.items()}, 'e2e', round(d['e2e']['ms_per_step'],2))"
}
for wm in 592 297 149 75; do GROVE_TUNE_WIDE_MAX=$wm run "wide2 wide_max=$wm"; done
for wm in 592 149; do GROVE_PLACE_LIB=$PWD/grove_b200/variants/wide1.so GROVE_TUNE_WIDE_MAX=$wm run "wide1 wide_max=$wm"; done
