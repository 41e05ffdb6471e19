#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 3000 gpurun_out/bench_c4.json; tail -5 gpurun_out/bench_c4.err
timeout 300 python bench.py --config C5 --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 1500 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1200 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
