# usage: bash scripts/gpu_prof.sh <workload> <kernel regex> <tag> [bytes]
W=$1; K=$2; TAG=$3; B=${4:-1073741824}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_$TAG -f \
    python bench.py --workload $W --bytes $B --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
