mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for pf in 2; do for w0 in 32; do
  GROVE_TUNE_PREFILTER=$pf GROVE_TUNE_WIDTH0=$w0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('pf=$pf w0=$w0', round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items()})"
done; done
bash gpurun_dbg.sh 2>&1 | grep -E "^round" | tail -8
