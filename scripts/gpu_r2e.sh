# round 2, fifth GPU pass: delta-form k_gram
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu --durations=5 -k "gram or dictionary or cfg1 or cfg3 or cfg5_dictionary_10k or config_patterns or irregular or random_patterns_on_inputs or nft_flavour" 2>&1 | tail -12
B="--bytes 1073741824 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-others"
for w in cfg1 cfg3 cfg5; do python bench.py --workload $w $B 2>gpurun_out/err_$w.log | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['verified'])"; done
TRRE_GRAM_NO_REG=1 python bench.py --workload cfg3 $B 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 smem-form', round(d['value'],1), round(d['roofline']['frac'],4), d['verified'])"
P="--steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-others --bytes 1073741824"
ncu --set full --clock-control none --import-source on -k regex:k_gram -s 3 -c 1 -o gpurun_out/prof_gram_cfg5_$TAG -f python bench.py --workload cfg5 $P > gpurun_out/ncu_gram_cfg5_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gram -s 3 -c 1 -o gpurun_out/prof_gram_cfg3_$TAG -f python bench.py --workload cfg3 $P > gpurun_out/ncu_gram_cfg3_$TAG.log 2>&1
