#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
for rep in ${REPS:-1}; do for ov in 1 0; do
VDQN_DDP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 5 --no-e2e > gpurun_out/ddp_ab_${ov}_${rep}.json 2> gpurun_out/ddp_ab_${ov}_${rep}.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ddp_ab_${ov}_${rep}.json'))
    print('overlap=$ov rep $rep N=$N: ms_per_step', round(d['ms_per_step'],4), 'dp', d["dp_check"]["max_param_diff_across_ranks"], d["per_rank_ms_without_exchange"])
except Exception as e:
    print('failed', e); print(open('gpurun_out/ddp_ab_${ov}_${rep}.err').read()[-2000:])
PY
done; done
timeout 300 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-inference > gpurun_out/ddp_bench_n1.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/ddp_bench_n1.json')); print('N=1 same box: ms_per_step', round(d['ms_per_step'],4))"
