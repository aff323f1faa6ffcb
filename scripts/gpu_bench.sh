# usage: bash scripts/gpu_bench.sh [tag]   -> results under gpurun_out/
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
kill $SMI
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
# launch list (cold, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_bench_$TAG.log 2>&1
# full capture of the top kernel
ncu --set full --clock-control none --import-source on -k regex:k_bytemap -s 3 -c 2 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/
