#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/mlp_probe.py > gpurun_out/c6_mlp_probe.txt 2>&1
tail -3 gpurun_out/c6_mlp_probe.txt
timeout 600 python -m pytest tests/test_gpu_teacher_forced.py -m gpu -q -x --no-header 2>&1 | tail -150 > gpurun_out/c6_tf.txt
tail -2 gpurun_out/c6_tf.txt; grep "^E  " gpurun_out/c6_tf.txt | head -5
bash tools/ncu_step.sh r02b 95 > /dev/null 2>&1
grep "mlp_gemm\|split_bf16" gpurun_out/launches_r02b_summary.txt
head -3 gpurun_out/launches_r02b_summary.txt
