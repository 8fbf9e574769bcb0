#!/bin/bash
# round-2 GPU call 1: new parity tests, TD bandwidth, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -x --no-header -rA 2>&1 | tail -60 > gpurun_out/c1_tests_full.txt
timeout 120 python tools/td_bandwidth.py > gpurun_out/c1_td_bulk.json 2> gpurun_out/c1_td_bulk.err
VDQN_TD_BULK=0 timeout 120 python tools/td_bandwidth.py > gpurun_out/c1_td_staged.json 2> gpurun_out/c1_td_staged.err
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err
timeout 1200 python -m pytest tests -m gpu -q --no-header --deselect tests/test_gpu_parity_full.py 2>&1 | tail -30 > gpurun_out/c1_tests_old.txt
tail -5 gpurun_out/c1_tests_full.txt; tail -3 gpurun_out/c1_tests_old.txt; cat gpurun_out/c1_td_bulk.json | head -20
