#!/bin/bash
# per-instruction stall samples of one kernel launch: tools/ncu_source.sh NAME REGEX SKIP -- command...
# writes gpurun_out/src_NAME.csv (ncu --page source: SASS + sampling columns) and the details page
NAME=$1; REGEX=$2; SKIP=$3; shift 4
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c 1 -o gpurun_out/src_$NAME -f "$@" > /dev/null 2> gpurun_out/src_$NAME.err
ncu -i gpurun_out/src_$NAME.ncu-rep --page source --csv > gpurun_out/src_$NAME.csv 2>/dev/null
ncu -i gpurun_out/src_$NAME.ncu-rep --page details --csv > gpurun_out/src_${NAME}_details.csv 2>/dev/null
ls -la gpurun_out/src_$NAME.*
rm -f gpurun_out/src_$NAME.ncu-rep
