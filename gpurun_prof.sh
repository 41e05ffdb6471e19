set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/launches_r1_b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score -s 48 -c 2 -o gpurun_out/prof_score_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_score.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fit -s 48 -c 2 -o gpurun_out/prof_fit_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_admit -s 48 -c 2 -o gpurun_out/prof_admit_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_admit.log 2>&1
ls -la gpurun_out
