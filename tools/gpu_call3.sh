#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_teacher_forced.py -m gpu -q -x --no-header 2>&1 | tail -150 > gpurun_out/c3_tf.txt
timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q --no-header 2>&1 | tail -400 > gpurun_out/c3_full.txt
timeout 120 python tools/td_probe.py > gpurun_out/c3_td_probe.json 2> gpurun_out/c3_err1.txt
timeout 120 python tools/td_bandwidth.py > gpurun_out/c3_td_bw.json 2> gpurun_out/c3_err2.txt
tail -25 gpurun_out/c3_tf.txt; tail -30 gpurun_out/c3_full.txt; cat gpurun_out/c3_td_probe.json
