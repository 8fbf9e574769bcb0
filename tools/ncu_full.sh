#!/bin/bash
# ncu --set full captures of the kernels of one step (details pages as CSV under gpurun_out/)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-inference"
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/full_$1 -f $CMD > /dev/null 2> gpurun_out/full_$1.err
  ncu -i gpurun_out/full_$1.ncu-rep --page details --csv > gpurun_out/ncu_full_$1.csv 2>/dev/null
  ncu -i gpurun_out/full_$1.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]
keep=[i for i,x in enumerate(h) if x in ('Kernel Name','Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active')]
for row in r[2:]:
    print({h[i]: row[i] for i in keep})
" > gpurun_out/ncu_key_$1.txt
  rm -f gpurun_out/full_$1.ncu-rep
  cat gpurun_out/ncu_key_$1.txt | cut -c1-600
}
# launches of the 4th eager step: skip counts are per kernel-name match
if [ "$1" == "halo" ]; then
  cap halo "halo_conv_kernel" 27 9
  CMD="python tools/td_bandwidth.py"
  cap td "td_epilogue" 2 2
else
  cap igemm "igemm_kernel" 100 6
  cap mlp "mlp_gemm_kernel" 27 9
  cap wgrad "wgrad_kernel" 48 4
fi
