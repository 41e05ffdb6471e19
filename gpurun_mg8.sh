mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$n.json 2> gpurun_out/bench_g$n.err; tail -2 gpurun_out/bench_g$n.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_g$n.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['ms_per_step'], d['config']['admitted'], d['config']['rounds'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 | cut -c1-200
