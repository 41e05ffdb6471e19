mkdir -p gpurun_out
for n in 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_g$n.json 2> gpurun_out/bench_g$n.err; tail -3 gpurun_out/bench_g$n.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_g$n.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e'], d['cpu_baseline']['placements_identical_to_gpu'])"
done
