python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -3
for w in del sub; do
  echo -n "== $w: "
  python bench.py --workload $w --bytes 1073741824 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'],'GB/s e2e', d['e2e'])"
done
