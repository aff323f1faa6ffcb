for w in num del; do
  echo "== $w staged"
  python bench.py --workload $w --bytes 1073741824 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'],'GB/s', d['ms_per_step'],'ms', d['config']['exec'])"
done
echo "== del via general staged path (TRRE_NO_BMV=1)"
TRRE_NO_BMV=1 python bench.py --workload del --bytes 1073741824 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'],'GB/s', d['ms_per_step'],'ms')"
