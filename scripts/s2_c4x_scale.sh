#!/bin/bash
# the sharded score pass at 1, 2, 4, 8 GPUs on one box (gpurun --gpus 8)
cd "$GRAFT_REPO_ROOT" || exit 1
export PYTHONPATH="$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then timeout 300 python bench.py --config C4X --steps 10 --warmup 3 > gpurun_out/r2_bench_C4X_n$n.json 2> gpurun_out/r2_bench_C4X_n$n.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --config C4X --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_C4X_n$n.json 2> gpurun_out/r2_bench_C4X_n$n.err; fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_C4X_n$n.json").read().strip().splitlines()[-1])
    print("C4X n=$n", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["ms_per_step"], 3), "roofline", round(d["roofline"]["frac"], 3), "coll_s", round(d["result"]["collective_s_per_step_max_over_ranks"], 6))
except Exception as ex:
    print("C4X n=$n failed", ex); print(open("gpurun_out/r2_bench_C4X_n$n.err").read()[-800:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_replicas_n8.json 2> gpurun_out/r2_bench_replicas_n8.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_replicas_n8.json").read().strip().splitlines()[-1])
    print("replicas n=8", round(d["ms_per_step"], 3), "ms", round(d["value"]), "gangs/s e2e", round(d["e2e"]["ms_per_step"], 3))
except Exception as ex:
    print("replicas failed", ex); print(open("gpurun_out/r2_bench_replicas_n8.err").read()[-800:])
PY
