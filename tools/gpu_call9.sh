#!/bin/bash
bash tools/gpu_check.sh
bash tools/ncu_step.sh r02c 86 > /dev/null 2>&1
head -30 gpurun_out/launches_r02c_summary.txt
