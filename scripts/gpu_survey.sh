for mb in 8 16 32 64 128; do
  echo -n "chunk $mb MB: "
  TRRE_CHUNK_MB=$mb python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['e2e']['value'])"
done
python -m pytest tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -2
