mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_ckpt.json 2>gpurun_out/err.txt; cut -c1-400 gpurun_out/bench_r1_ckpt.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ckpt_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 160 --csv --log-file gpurun_out/launches_r1_ckpt.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score -s 64 -c 1 -o gpurun_out/prof_score_r1_ckpt python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_admit -s 64 -c 2 -o gpurun_out/prof_admit_r1_ckpt python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls gpurun_out | head -30
