# round 2, first GPU pass: parity of the new kernels, then 1 GiB bench lines and ncu captures
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
B="--bytes 1073741824 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2', round(d['value'],1),'GB/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), d['config']['exec'], 'verified',d['verified'], 'copy', d['roofline']['torch_copy_same_size_gbs'])"; }
for v in default nosw c1 c3 c4; do
  case $v in default) E="";; nosw) E="TRRE_NO_SWAR=1";; c1) E="TRRE_SWAR_CTAS=1";; c3) E="TRRE_SWAR_CTAS=3";; c4) E="TRRE_SWAR_CTAS=4";; esac
  env $E python bench.py --workload cfg2 $B --steps 30 2>gpurun_out/err_$v.log | tail -1 > gpurun_out/bench_cfg2_${v}_$TAG.json; show gpurun_out/bench_cfg2_${v}_$TAG.json cfg2_$v
done
python bench.py --workload cfg4 $B --steps 30 2>/dev/null | tail -1 > gpurun_out/bench_cfg4_$TAG.json; show gpurun_out/bench_cfg4_$TAG.json cfg4
for w in cfg1 cfg3 cfg5; do
  python bench.py --workload $w $B 2>gpurun_out/err_$w.log | tail -1 > gpurun_out/bench_${w}_$TAG.json; show gpurun_out/bench_${w}_$TAG.json $w
  TRRE_NO_GRAM=1 python bench.py --workload $w $B 2>/dev/null | tail -1 > gpurun_out/bench_${w}_nogram_$TAG.json; show gpurun_out/bench_${w}_nogram_$TAG.json ${w}_nogram
done
P="--steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --bytes 1073741824"
ncu --set full --clock-control none --import-source on -k regex:k_gram -s 3 -c 1 -o gpurun_out/prof_gram_cfg5_$TAG -f python bench.py --workload cfg5 $P > gpurun_out/ncu_gram_cfg5_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gram -s 3 -c 1 -o gpurun_out/prof_gram_cfg3_$TAG -f python bench.py --workload cfg3 $P > gpurun_out/ncu_gram_cfg3_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bytemap_swar -s 3 -c 1 -o gpurun_out/prof_swar_cfg2_$TAG -f python bench.py --workload cfg2 $P > gpurun_out/ncu_swar_cfg2_$TAG.log 2>&1
ls -la gpurun_out | tail -20
