# Round-end evidence: full GPU test-suite, smoke, official bench line (both arms), ncu launch list + full capture.
TAG=${1:-r1_final}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bytemap -s 3 -c 1 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
for w in cfg1 cfg3 cfg5 del sub num cfg4; do
  python bench.py --workload $w --bytes 1073741824 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 > gpurun_out/bench_${w}_$TAG.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_${w}_$TAG.json')); print('$w', round(d['value'],1),'GB/s', round(d['ms_per_step'],3),'ms frac',round(d['roofline']['frac'],3), d['config']['exec'], 'verified',d['verified'],'e2e',round(d['e2e']['value'],1))"
done
