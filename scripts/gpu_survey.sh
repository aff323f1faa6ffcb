mkdir -p gpurun_out
for w in cfg1 cfg3 cfg5 del cfg4; do
  echo "== $w"
  python bench.py --workload $w --bytes 1073741824 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'],'GB/s', d['ms_per_step'],'ms frac',d['roofline']['frac'],'launches',d['gpu_launches'],'verified',d['verified'],'e2e',d['e2e'])"
done
echo "== 2 GPUs cfg2"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -2
