# Round-end bench evidence without the test-suite (run separately): both bench arms, the ncu launch list.
TAG=${1:-r2b_final}
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$TAG.json
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4),'verified',d['verified'], 'clocks', d['clocks'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='pcie'}); print('pcie', d['e2e'].get('pcie'))
print('cli', d.get('cli_e2e')); print('cpu', d.get('cpu_baseline'))
for k,v in (d.get('others') or {}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('pattern','verified_how','workload','unit','bytes_per_gpu')})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-others > /dev/null 2>&1
