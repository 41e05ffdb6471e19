for mb in 6 8 10; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fopenmp -shared -cudart shared -lgomp -DGROVE_ADMIT_MINBLOCKS=$mb -o grove_b200/libgrove_place.so grove_b200/csrc/engine.cu 2>&1 | grep -i error
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('minblocks=$mb', round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items()}, round(d['e2e']['ms_per_step'],2))"
done
