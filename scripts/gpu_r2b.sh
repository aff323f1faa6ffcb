# round 2, second GPU pass: lazy tables, DFS tiers, reworked k_gram, SWAR launch shapes
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu -k "gram or lazy or dfs or golden or dictionary or config_patterns or irregular" 2>&1 | tail -15
B="--bytes 1073741824 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2', round(d['value'],1),'GB/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), d['config']['exec'], 'verified',d['verified'], 'copy', d['roofline']['torch_copy_same_size_gbs'])"; }
for v in default v1c4 v1c8 v2; do
  case $v in default) E="";; v1c4) E="TRRE_SWAR_VARIANT=1 TRRE_SWAR_CTAS=4";; v1c8) E="TRRE_SWAR_VARIANT=1 TRRE_SWAR_CTAS=8";; v2) E="TRRE_SWAR_VARIANT=2";; esac
  env $E python bench.py --workload cfg2 $B --steps 30 2>gpurun_out/err_$v.log | tail -1 > gpurun_out/bench_cfg2_${v}_$TAG.json; show gpurun_out/bench_cfg2_${v}_$TAG.json cfg2_$v
done
for w in cfg1 cfg3 cfg5; do
  python bench.py --workload $w $B 2>gpurun_out/err_$w.log | tail -1 > gpurun_out/bench_${w}_$TAG.json; show gpurun_out/bench_${w}_$TAG.json $w
done
P="--steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --bytes 1073741824"
ncu --set full --clock-control none --import-source on -k regex:k_gram -s 3 -c 1 -o gpurun_out/prof_gram_cfg5_$TAG -f python bench.py --workload cfg5 $P > gpurun_out/ncu_gram_cfg5_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gram -s 3 -c 1 -o gpurun_out/prof_gram_cfg3_$TAG -f python bench.py --workload cfg3 $P > gpurun_out/ncu_gram_cfg3_$TAG.log 2>&1
ls -la gpurun_out | tail -8
