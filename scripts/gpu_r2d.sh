# round 2, fourth GPU pass: the var-length gram front end, cheap DFS cases, ncu captures of the final kernels
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --durations=5 -k "variable_length or dfs or cyclic or dictionary_on_gpu or random_patterns_on_inputs" 2>&1 | tail -12
for w in num sub del; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value'],1), round(d['roofline']['frac'],4), d['config_detail']['exec'], d['verified'], 'e2e', round(d['e2e']['value'],1))"; done
TRRE_NO_GRAM=1 python bench.py --workload sub --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sub nogram', round(d['value'],1), round(d['roofline']['frac'],4))"
P="--steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-others --bytes 1073741824"
ncu --set full --clock-control none --import-source on -k regex:k_gram -s 3 -c 1 -o gpurun_out/prof_gram_cfg5_$TAG -f python bench.py --workload cfg5 $P > gpurun_out/ncu_gram_cfg5_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gram -s 3 -c 1 -o gpurun_out/prof_gram_cfg3_$TAG -f python bench.py --workload cfg3 $P > gpurun_out/ncu_gram_cfg3_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bytemap_swar -s 3 -c 1 -o gpurun_out/prof_swar_cfg2_$TAG -f python bench.py --workload cfg2 $P > gpurun_out/ncu_swar_cfg2_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-others > /dev/null 2>&1
ls -la gpurun_out | tail -6
