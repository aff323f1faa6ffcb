# round 2: k_gram with shifts on the FMA pipe (cfg5), k_lit at 6 CTAs per SM (cfg1 / cfg3): parity, 1 GiB bench lines
TAG=${1:-r2l}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "literal or cfg1 or cfg3 or config_patterns or dictionary or cfg5_dictionary_10k or random_patterns" 2>&1 | tail -5
for w in cfg5 cfg3 cfg1; do
  timeout 200 python bench.py --workload $w --bytes 1073741824 --steps 20 --warmup 3 --no-others --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${w}_$TAG.json").read().strip().splitlines()[-1])
print("$w", round(d["value"], 1), d["ms_per_step"], round(d["roofline"]["frac"], 4), d["config_detail"]["exec"], d["verified"])
PY
done
