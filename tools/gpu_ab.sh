#!/bin/bash
# A/B of an engine switch on the same box: tools/gpu_ab.sh ENVVAR [A B]  (runs bench with ENVVAR=A, =B, =A, =B; default 1 0)
V=${1:-VDQN_FUSE_POOL}
A=${2:-1}
B=${3:-0}
mkdir -p gpurun_out
for rep in 1 2; do for val in $A $B; do
  env $V=$val timeout 300 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-inference > gpurun_out/ab_${val}_${rep}.json 2> gpurun_out/ab_${val}_${rep}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_${val}_${rep}.json')); print('$V=$val rep $rep: ms_per_step', round(d['ms_per_step'],4), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/ab_${val}_${rep}.err').read()[-1500:])
PY
done; done
