set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print({k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step')}, d['e2e'], d['roofline']['frac'])
PY
ncu --set full --clock-control none --import-source on -k regex:k_score -s 48 -c 1 -o gpurun_out/prof_score_r1b python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_score.log 2>&1
python - <<'PY'
import torch, time
x=torch.empty(1400*1024*1024, dtype=torch.uint8, device='cuda')
for _ in range(3): x.zero_()
torch.cuda.synchronize()
s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): x.zero_()
e.record(); torch.cuda.synchronize()
print('memset GB/s', x.numel()*10/ (s.elapsed_time(e)*1e-3)/1e9)
y=torch.empty_like(x)
s.record()
for _ in range(10): y.copy_(x)
e.record(); torch.cuda.synchronize()
print('copy GB/s (r+w)', 2*x.numel()*10/ (s.elapsed_time(e)*1e-3)/1e9)
PY
