#!/bin/bash
# ncu launch list of ONE eager step (duration + DRAM bytes + tensor-pipe activity per launch).
# usage: tools/ncu_step.sh <tag> [launches_per_step]
TAG=${1:-step}; LPS=${2:-100}
mkdir -p gpurun_out
SKIP=$((LPS * 5))
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --launch-skip $SKIP -c $LPS --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-inference > /dev/null 2> gpurun_out/ncu_${TAG}.err
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv all > gpurun_out/launches_${TAG}_summary.txt 2>&1
head -40 gpurun_out/launches_${TAG}_summary.txt
