mkdir -p gpurun_out
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N "$@" 2>gpurun_out/scale_err_$N.log | tail -1; }
echo "== cfg2 N=8"; run 8 --steps 20 --warmup 3 | tee gpurun_out/scale_cfg2_n8.json | cut -c1-400
echo "== cfg2 N=4"; run 4 --steps 20 --warmup 3 --e2e-steps 2 | tee gpurun_out/scale_cfg2_n4.json | cut -c1-200
echo "== cfg4 N=8 (8 GiB per GPU = 64 GiB)"; run 8 --workload cfg4 --steps 5 --warmup 3 --e2e-steps 0 | tee gpurun_out/scale_cfg4_n8.json | cut -c1-400
echo "== reference arm under torchrun"; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-200
tail -3 gpurun_out/scale_err_8.log
