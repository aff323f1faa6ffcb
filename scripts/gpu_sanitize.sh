# usage: bash scripts/gpu_sanitize.sh [memcheck|racecheck|initcheck|synccheck]
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/$TOOL.log 2>&1; echo "$TOOL rc=$?"
grep -v "^=========     " gpurun_out/$TOOL.log | tail -12
