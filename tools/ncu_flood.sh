#!/bin/bash
# ncu --set full of the flood round kernel on the cfg2 vessel: TAG, distance function
TAG=$1; DF=${2:-1}; O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"flood_round" -c 1 -f -o $O/${TAG} python tools/prof_flood1.py $DF > $O/${TAG}.log 2>&1
ncu -i $O/${TAG}.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
ncu -i $O/${TAG}.ncu-rep --page source --csv --print-source sass > $O/${TAG}_sass.csv 2>/dev/null
tail -3 $O/${TAG}.log
