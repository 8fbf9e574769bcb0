#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/ddp_timeline.py > gpurun_out/ddp_timeline_n$N.txt 2>&1
grep "^rank\|Error\|error" gpurun_out/ddp_timeline_n$N.txt | head -20
