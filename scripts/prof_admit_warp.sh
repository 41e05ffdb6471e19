mkdir -p gpurun_out
GROVE_TUNE_WIDTH0=24 ncu --set full --clock-control none --import-source on -k regex:k_admit_warp -s 6 -c 1 -o gpurun_out/prof_warp python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/prof_warp.ncu-rep
