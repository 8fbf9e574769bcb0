#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/mlp_probe.py > gpurun_out/c5_mlp_probe.txt 2>&1
cat gpurun_out/c5_mlp_probe.txt | tail -30
timeout 600 python -m pytest tests/test_gpu_teacher_forced.py -m gpu -q -x --no-header 2>&1 | tail -150 > gpurun_out/c5_tf.txt
tail -4 gpurun_out/c5_tf.txt; grep "^E  " gpurun_out/c5_tf.txt | head -5
