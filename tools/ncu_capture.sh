#!/bin/bash
# tools/ncu_capture.sh TAG KERNEL_REGEX SKIP "prof_target args" (or CMD="..." to profile another command) -- one `ncu --set full` capture of one launch on the GPU
# box, exported as raw / details / source CSV into gpurun_out/ (the .ncu-rep itself stays on the box: gpurun_out is
# limited to 64 MiB).  Development tool.
TAG=$1; KRE=$2; SKIP=$3; ARGS=$4
REP=/tmp/$TAG
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o $REP -f ${CMD:-python tools/prof_target.py $ARGS} > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ls -la $REP.ncu-rep gpurun_out/${TAG}_*
