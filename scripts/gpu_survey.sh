python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "match_kernel or variable_length or config or full_size" 2>&1 | tail -3
for w in cfg1 cfg3 sub; do
  echo -n "== $w: "
  python bench.py --workload $w --bytes 1073741824 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'],'GB/s', d['ms_per_step'],'ms', d['verified'])"
done
