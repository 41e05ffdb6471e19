# 1..N GPU strong-scaling run on one box: gpurun --gpus 8 -- "bash scripts/gpu_scale.sh 1 2 4 8"
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = "1" ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_g1.json 2> gpurun_out/scale_g1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_g$n.json 2> gpurun_out/scale_g$n.err; fi
  python -c "
import json
d=json.loads(open('gpurun_out/scale_g$n.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('n_gpus','value','ms_per_step')}, round(d['e2e']['ms_per_step'],2), d['result']['admitted'], d['result']['relaxation_rounds_rank0'])"
done
