#!/bin/bash
mkdir -p gpurun_out
for cfg in "12 4" "18 6" "18 4" "15 5"; do
  set -- $cfg
  VDQN_NVCC_FLAGS="-DVDQN_POOL_STAGES=$1 -DVDQN_POOL_DEPTH=$2" python video_dqn_b200/build.py --force > /dev/null 2>&1
  echo "stages $1 depth $2:"; timeout 200 python tools/pool_probe.py 2>&1 | grep "768 frames\|no pooling work\|nothing but\|OK\|FAILED"
done
