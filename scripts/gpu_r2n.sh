# round 2: the literal word form with outputs of any length: parity, then the off-BASELINE bench lines
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "any_length or run_form or literal or edge_inputs or config_patterns" 2>&1 | tail -5
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
for w in sub del num; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-others --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${w}_$TAG.json").read().strip().splitlines()[-1])
print("$w", round(d["value"], 1), d["ms_per_step"], round(d["roofline"]["frac"], 4), d["config_detail"]["exec"], d["verified"], d["gpu_launches"], (d.get("e2e") or {}).get("value"))
PY
done
