#!/bin/bash
# A/B of an environment switch under data parallelism on one box: tools/gpu_ddp_ab_env.sh N VAR A B
N=$1; V=$2; A=$3; B=$4
mkdir -p gpurun_out
for rep in 1 2; do for val in $A $B; do
  env $V=$val VDQN_DDP=nvl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-inference > gpurun_out/ddpab_${val}_${rep}.json 2> gpurun_out/ddpab_${val}_${rep}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ddpab_${val}_${rep}.json')); print('$V=$val rep $rep: ms_per_step', round(d['ms_per_step'],4), 'local', d.get('per_rank_ms_without_exchange'), 'param diff', d['dp_check']['max_param_diff_across_ranks'], 'grad', d['dp_check']['avg_grad_rel_l2_vs_global_batch'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/ddpab_${val}_${rep}.err').read()[-1500:])
PY
done; done
