#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_basic_train.py -m gpu -q --no-header 2>&1 | tail -80 > gpurun_out/c10_basic.txt
grep -n "passed\|failed\|^FAILED\|^E   \|FAIL " gpurun_out/c10_basic.txt | head -30
cat gpurun_out/basic_train_report_f4.txt 2>/dev/null | grep -v "^ok" | head
