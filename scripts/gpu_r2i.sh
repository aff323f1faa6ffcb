# round 2: compute-sanitizer over every kernel family + differential fuzz on the GPU
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r2_$tool.log 2>&1; echo "$tool rc=$?"
  grep -v "^=========     " gpurun_out/r2_$tool.log | tail -6
done
timeout 600 python scripts/gpu_fuzz.py 20260924 1500 2>&1 | tail -5 | tee gpurun_out/r2_fuzz.log
