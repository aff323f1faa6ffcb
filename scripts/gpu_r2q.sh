# round 2: k_gram with the bitmap regrouped by rest + PRMT/IMAD class lookups: parity, cfg5 1 GiB
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_golden.py -m gpu -x -q -k "dictionary or cfg5_dictionary_10k or random_patterns or config_patterns or golden_scan or position_parallel" 2>&1 | tail -4
for w in cfg5; do
  timeout 200 python bench.py --workload $w --bytes 1073741824 --steps 20 --warmup 3 --no-others --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${w}_$TAG.json").read().strip().splitlines()[-1])
print("$w", round(d["value"], 1), d["ms_per_step"], round(d["roofline"]["frac"], 4), d["config_detail"]["exec"], d["verified"])
PY
done
TRRE_GRAM_SHIFT_CLASS=1 timeout 200 python bench.py --workload cfg5 --bytes 1073741824 --steps 20 --warmup 3 --no-others --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5 shift-class', round(d['value'],1), d['ms_per_step'], round(d['roofline']['frac'],4), d['verified'])"
