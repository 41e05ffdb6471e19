# Round-end capture on the GPU box: /usr/local/graft/bin/gpurun --timeout 1200 -- "bash scripts/gpu_capture.sh"
# (tests, smoke, bench, reference arm, ncu launch list + full captures of k_score / k_admit / k_resolve into gpurun_out/)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/final_n1.json 2>gpurun_out/final_n1.err; cut -c1-300 gpurun_out/final_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json; cut -c1-200 gpurun_out/final_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 168 -c 45 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score -s 10 -c 1 -o gpurun_out/final_score python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_admit -s 10 -c 4 -o gpurun_out/final_admit python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 10 -c 1 -o gpurun_out/final_resolve python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls gpurun_out | grep final
