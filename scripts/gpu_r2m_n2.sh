# 2 GPUs of one box: trre_scan_multi / TRRE_GPUS against the single-GPU result (the test the 1-GPU round-end run skips),
# and the bench contract under torchrun with the round's new kernels (k_lit: cfg3) - short, headline workload cfg3
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_cli.py -m gpu -q -k "multi_gpu" -rA 2>&1 | tail -6 | tee gpurun_out/multi_gpu_test_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload cfg3 --bytes 1073741824 --steps 10 --warmup 3 --no-others --no-cpu-baseline --e2e-steps 2 > gpurun_out/scale_cfg3_lit_n2.json 2> gpurun_out/scale_cfg3_lit_n2.err
tail -c 1200 gpurun_out/scale_cfg3_lit_n2.json
