# usage (under gpurun --gpus N): bash scripts/gpu_scale.sh N [tag]  -> the torchrun bench line of N ranks (+ both arms at N)
N=${1:-8}; TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/topo_$TAG.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((N % 10)) bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 3 \
   > gpurun_out/scale_n${N}_$TAG.json 2> gpurun_out/scale_n${N}_$TAG.err
tail -3 gpurun_out/scale_n${N}_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_n${N}_$TAG.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'verified', d['verified'])
print('e2e', {k:v for k,v in d['e2e'].items() if k not in ('pcie',)}); print('pcie', d['e2e'].get('pcie'))
print('gather', d['gather'])
for k,v in (d.get('others') or {}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('pattern','verified_how')})
PY
