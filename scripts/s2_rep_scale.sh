#!/bin/bash
# replicas at 1 and N GPUs on one box (weak scaling of the headline line), with the polling depth varied at N
cd "$GRAFT_REPO_ROOT" || exit 1
export PYTHONPATH="$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${N:-4}
run() { # tag, n, env...
  tag=$1; n=$2; shift 2
  if [ $n = 1 ]; then env "$@" timeout 300 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/s2_rs_$tag.json 2> gpurun_out/s2_rs_$tag.err
  else env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/s2_rs_$tag.json 2> gpurun_out/s2_rs_$tag.err; fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/s2_rs_$tag.json").read().splitlines() if l.startswith("{")][-1])
    print("$tag n=$n", round(d["ms_per_step"], 3), "ms", round(d["value"]), "gangs/s e2e", round(d["e2e"]["ms_per_step"], 3), "dev", round(d["kernel_ms_per_step"]["ms_total"], 3))
except Exception as ex:
    print("$tag failed", ex); print(open("gpurun_out/s2_rs_$tag.err").read()[-600:])
PY
}
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"
run n1 1 A=1
run nN $N A=1
run nN_ahead6 $N GROVE_TUNE_AHEAD=6
run nN_threads2 $N GROVE_HOST_THREADS=2
