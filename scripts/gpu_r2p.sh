# round 2: k_gram delta emission (cfg5): parity, 1 GiB bench line, one ncu capture
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "dictionary or cfg5_dictionary_10k or any_length" 2>&1 | tail -5
for w in cfg5 sub; do
  timeout 200 python bench.py --workload $w --bytes 1073741824 --steps 20 --warmup 3 --no-others --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${w}_$TAG.json").read().strip().splitlines()[-1])
print("$w", round(d["value"], 1), d["ms_per_step"], round(d["roofline"]["frac"], 4), d["config_detail"]["exec"], d["verified"])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gram -s 3 -c 1 -o gpurun_out/prof_gram_cfg5_$TAG -f \
    python bench.py --workload cfg5 --bytes 1073741824 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-others > gpurun_out/ncu_gram_cfg5_$TAG.log 2>&1
tail -1 gpurun_out/ncu_gram_cfg5_$TAG.log | cut -c1-200
