#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/td_probe.py > gpurun_out/c2_td_probe_bulk.json 2> gpurun_out/c2_err1.txt
VDQN_TD_BULK=0 timeout 120 python tools/td_probe.py > gpurun_out/c2_td_probe_staged.json 2> gpurun_out/c2_err2.txt
timeout 300 ncu --set full --clock-control none -k regex:td_epilogue -c 4 --csv --page raw --log-file gpurun_out/c2_ncu_td_bulk.csv python tools/td_bandwidth.py > /dev/null 2> gpurun_out/c2_err3.txt
cat gpurun_out/c2_td_probe_bulk.json gpurun_out/c2_td_probe_staged.json
