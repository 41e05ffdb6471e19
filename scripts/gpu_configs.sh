# bench lines of the smaller BASELINE.json configs (parity-test cases; recorded for reference)
mkdir -p gpurun_out
for c in C1 C2 C3; do
  python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_$c.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/bench_$c.json').read().strip().splitlines()[-1]); print('$c', round(d['ms_per_step'],3), 'ms', round(d['value']), 'gangs/s e2e', round(d['e2e']['ms_per_step'],3), 'ms rounds', d['config']['rounds'], 'cpu', round(d['cpu_baseline']['value']), d['cpu_baseline']['placements_identical_to_gpu'])"
done
# C5: steady-state churn (a step = one 100 ms scheduler tick)
python bench.py --config C5 --steps 30 --warmup 5 > gpurun_out/bench_C5.json 2>gpurun_out/bench_C5.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_C5.json').read().strip().splitlines()[-1]); c=d['config']; print('C5 cycle', round(d['ms_per_step'],3), 'ms tick', round(c['tick_ms_mean'],3), 'ms (max', round(c['tick_ms_max'],3), ') pending/tick', round(c['pending_per_tick_mean']), 'admitted', c['admitted'], 'cpu', round(d['cpu_baseline']['value']), d['cpu_baseline']['placements_identical_to_gpu'])" || tail -5 gpurun_out/bench_C5.err
