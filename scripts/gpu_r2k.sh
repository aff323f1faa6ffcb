# round 2, k_lit EQ form: parity, 1 GiB bench lines of cfg1 / cfg3, one ncu capture of k_lit on cfg3
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "literal or cfg1 or cfg3 or config_patterns" 2>&1 | tail -5
for w in cfg3 cfg1; do
  timeout 200 python bench.py --workload $w --bytes 1073741824 --steps 20 --warmup 3 --no-others --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_${w}_lit_$TAG.json 2> gpurun_out/bench_${w}_lit_$TAG.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${w}_lit_$TAG.json").read().strip().splitlines()[-1])
print("$w", round(d["value"], 1), d["ms_per_step"], round(d["roofline"]["frac"], 4), d["config_detail"]["exec"], d["verified"])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lit -s 3 -c 1 -o gpurun_out/prof_lit_cfg3_$TAG -f \
    python bench.py --workload cfg3 --bytes 1073741824 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-others > gpurun_out/ncu_lit_cfg3_$TAG.log 2>&1
tail -2 gpurun_out/ncu_lit_cfg3_$TAG.log | cut -c1-300
