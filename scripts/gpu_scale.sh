mkdir -p gpurun_out
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N "$@" 2>gpurun_out/scale_err_$N.log | tail -1; }
echo "== cfg2 N=8"; run 8 --steps 20 --warmup 3 --e2e-steps 2 > gpurun_out/scale_cfg2_n8.json; python -c "
import json; d=json.load(open('gpurun_out/scale_cfg2_n8.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gather'])"
echo "== cfg5 N=8 (1 GiB slices)"; run 8 --workload cfg5 --bytes 1073741824 --steps 5 --warmup 3 --e2e-steps 0 > gpurun_out/scale_cfg5_n8.json; python -c "
import json; d=json.load(open('gpurun_out/scale_cfg5_n8.json')); print(d['value'], d['ms_per_step'], d['verified'])"
