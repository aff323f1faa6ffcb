mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -4
for w in del; do
  echo "== $w"
  python bench.py --workload $w --bytes 1073741824 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'],'GB/s', d['ms_per_step'],'ms frac',d['roofline']['frac'],'launches',d['gpu_launches'],'verified',d['verified'],'e2e',d['e2e']['value'])"
done
