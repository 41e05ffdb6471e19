mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_admit -s 10 -c 2 -o gpurun_out/prof_admit_v2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/prof_admit_v2.ncu-rep
