# round 2, third GPU pass: whole GPU suite with durations, then the new bench line (others, e2e+pcie, cli_e2e)
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu --durations=12 2>&1 | tail -40
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2c.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4),'verified',d['verified'], 'copy', d['roofline']['torch_copy_same_size_gbs'], 'smcopy', d['roofline']['torch_elementwise_kernel_same_size_gbs'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='pcie'}); print('pcie', d['e2e'].get('pcie'))
print('cli', d.get('cli_e2e')); print('cpu', d.get('cpu_baseline'))
for k,v in (d.get('others') or {}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('pattern','verified_how')})
PY
TRRE_NO_SWAR=1 python bench.py --steps 30 --warmup 3 --no-others --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair-table k_bytemap', round(d['value'],1), round(d['roofline']['frac'],4), d['verified'])"
for w in num sub del; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value'],1), round(d['roofline']['frac'],4), d['config_detail']['exec'], d['verified'])"; done
