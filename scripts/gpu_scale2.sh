run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N "$@" 2>/dev/null | tail -1; }
for w in cfg5 cfg3 sub; do echo "== $w N=2"; run 2 --workload $w --bytes 1073741824 --steps 5 --warmup 3 --e2e-steps 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'],'GB/s',d['ms_per_step'],'ms verified',d['verified'],'sizes',d['output_bytes_per_rank'],'e2e',d['e2e']['value'])"; done
