set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; tail -5 gpurun_out/bench_g2.err; cat gpurun_out/bench_g2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','kernel_ms_per_step')}, d['e2e'], d['cpu_baseline'])"
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','kernel_ms_per_step')}, d['e2e'])"
