#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_teacher_forced.py -m gpu -q -x --no-header 2>&1 | tail -150 > gpurun_out/c4_tf.txt
tail -5 gpurun_out/c4_tf.txt; grep "^E  " gpurun_out/c4_tf.txt | head -5
