set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; tail -3 gpurun_out/bench_r1_a.err; cat gpurun_out/bench_r1_a.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_a.json; cat gpurun_out/bench_ref_a.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_a.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
