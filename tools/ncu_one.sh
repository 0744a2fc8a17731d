#!/bin/bash
# one ncu --set full capture: TAG, kernel regex, stages for prof_stage.py (512^3), launch skip count
TAG=$1; KRE=$2; STAGES=${3:-naive,c1}; SKIP=${4:-1}
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 -f -o $O/${TAG} python tools/prof_stage.py 512 $STAGES 2 > $O/${TAG}.log 2>&1
ncu -i $O/${TAG}.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
ncu -i $O/${TAG}.ncu-rep --page source --csv --print-source sass > $O/${TAG}_sass.csv 2>/dev/null
ls -la $O/${TAG}*
