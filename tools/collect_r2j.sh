#!/bin/bash
# Evidence for the build with the thin-front flood solver (one B200, through gpurun; outputs land in gpurun_out/):
#   GPU test log, bench line (default run), flood timings with and without the front solver, ncu --set full of the front kernel.
# The cfg3 kernels, the reference arm and the launch list of the cfg3 step are those of profiles/r2i_* (unchanged code).
set -u
TAG=${1:-r2j}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/${TAG}_gpu_tests.log 2>&1; echo "pytest rc $?" >> $O/${TAG}_gpu_tests.log
tail -3 $O/${TAG}_gpu_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python tools/show_bench.py $O/${TAG}_bench.json
(timeout 100 python tools/prof_flood.py 256 256; timeout 100 python tools/prof_flood.py 256 256 0; timeout 100 python tools/prof_flood.py 512 128; timeout 100 python tools/prof_flood.py 512 128 0) > $O/${TAG}_flood_timings.txt 2>&1
cat $O/${TAG}_flood_timings.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"flood_front_kernel" -c 1 -f -o $O/${TAG}_front_full python tools/prof_flood1.py 1 > $O/${TAG}_front_full.log 2>&1
ncu -i $O/${TAG}_front_full.ncu-rep --page raw --csv > $O/${TAG}_front_full_raw.csv 2>/dev/null
python tools/ncu_read.py $O/${TAG}_front_full_raw.csv 2>&1 | head -40
