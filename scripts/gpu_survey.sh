for v in 0 1 2 3 0; do
  echo -n "variant $v: "
  TRRE_BM_VARIANT=$v python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['verified'])"
done
