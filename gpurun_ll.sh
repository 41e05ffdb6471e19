mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 90 --csv --log-file gpurun_out/launches_d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_d.log 2>&1
