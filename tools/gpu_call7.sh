#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pool_probe.py > gpurun_out/c7_pool_probe.txt 2>&1
tail -12 gpurun_out/c7_pool_probe.txt
