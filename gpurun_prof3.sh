mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_admit -s 13 -c 1 -o gpurun_out/prof_admit_late python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/prof_admit_late.ncu-rep
