python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-steps 1 2>gpurun_out/g2.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'],'GB/s',d['ms_per_step'],'ms', d['gather'])"
tail -3 gpurun_out/g2.err
