#!/bin/bash
# usage: tools/gpu_ddp.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/ddp_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/ddp_probe.py > gpurun_out/ddp_probe_n$N.txt 2>&1
tail -25 gpurun_out/ddp_probe_n$N.txt
for kind in nvl nccl; do
VDQN_DDP=$kind timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/ddp_bench_${kind}_n$N.json 2> gpurun_out/ddp_bench_${kind}_n$N.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ddp_bench_${kind}_n$N.json'))
    print('$kind N=$N: ms_per_step', round(d['ms_per_step'],4), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp_check', d['dp_check'], d.get('grad_exchange'))
except Exception as e:
    print('$kind failed', e); print(open('gpurun_out/ddp_bench_${kind}_n$N.err').read()[-3000:])
PY
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-inference > gpurun_out/ddp_bench_n1.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/ddp_bench_n1.json')); print('N=1 same box: ms_per_step', round(d['ms_per_step'],4))"
