#!/bin/bash
# multi-GPU checks (run under gpurun --gpus N): the NCCL tests, the sharded score pass (C4X) at 1..N GPUs, the replicas line at N
cd "$GRAFT_REPO_ROOT" || exit 1
export PYTHONPATH="$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_sharded_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -5 | tee gpurun_out/s2_multi_tests.log
for n in $(seq 1 $N); do
  case $n in 1|2|4|8) ;; *) continue;; esac
  if [ $n = 1 ]; then timeout 600 python bench.py --config C4X --steps 10 --warmup 3 > gpurun_out/s2_c4x_g$n.json 2> gpurun_out/s2_c4x_g$n.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --config C4X --gpus $n --steps 10 --warmup 3 > gpurun_out/s2_c4x_g$n.json 2> gpurun_out/s2_c4x_g$n.err; fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/s2_c4x_g$n.json").read().strip().splitlines()[-1])
    print("C4X n=$n", round(d["ms_per_step"], 3), "ms", "e2e", round(d["e2e"]["ms_per_step"], 3), "ms", "roofline", round(d["roofline"]["frac"], 3), d["result"])
except Exception as ex:
    print("C4X n=$n failed", ex); print(open("gpurun_out/s2_c4x_g$n.err").read()[-1500:])
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s2_rep_g$N.json 2> gpurun_out/s2_rep_g$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/s2_rep_g$N.json").read().strip().splitlines()[-1])
    print("replicas n=$N", round(d["ms_per_step"], 3), "ms", round(d["value"]), "gangs/s")
except Exception as ex:
    print("replicas failed", ex); print(open("gpurun_out/s2_rep_g$N.err").read()[-1500:])
PY
