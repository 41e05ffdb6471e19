mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 140 --csv --log-file gpurun_out/launches_c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_c.log 2>&1
tail -1 gpurun_out/ncu_bench_c.log | cut -c1-200
