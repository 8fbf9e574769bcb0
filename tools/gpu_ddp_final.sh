#!/bin/bash
# usage: tools/gpu_ddp_final.sh N  -- the bench line at N GPUs (NVLink exchange) and at N = 1 on the same box
N=${1:-2}
mkdir -p gpurun_out
VDQN_DDP=nvl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench_n1_same_box_as_n$N.json 2>/dev/null
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/final_bench_n$N.json'))
    print('N=$N: ms_per_step', round(d['ms_per_step'],4), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp_check', d['dp_check'], d.get('grad_exchange'))
    print('per rank', d.get('per_rank_ms_without_exchange'), d.get('slowest_rank_bound_on_efficiency'))
except Exception as e:
    print('failed', e); print(open('gpurun_out/final_bench_n$N.err').read()[-3000:])
d=json.load(open('gpurun_out/final_bench_n1_same_box_as_n$N.json')); print('N=1 same box: ms_per_step', round(d['ms_per_step'],4))
PY
