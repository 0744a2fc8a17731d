#!/bin/bash
# Round-2 evidence for profiles/ (one B200, through gpurun; outputs land in gpurun_out/ with the given tag):
#   GPU test log, bench line (default run), reference arm, flood timings, ncu launch list of the bench, ncu --set full of the cfg3 kernels
#   and of the flood round kernel.
set -u
TAG=${1:-r2}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_gpu_tests.log 2>&1; echo "pytest rc $?" >> $O/${TAG}_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 200 python tools/prof_flood.py 256 256 > $O/${TAG}_flood_timings.txt 2>&1
timeout 200 python tools/prof_vessel.py 512 > $O/${TAG}_vessel_stage_timings.txt 2>&1
timeout 120 python tools/prof_voxelize.py 256 512 > $O/${TAG}_voxelize_timings.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 600 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batch --no-vessel > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"naive_brick|certificate|resolve_kernel|stencil_fast|erode_sparse|sweep_sparse|sweep_apply|histogram" -s 10 -c 10 -f -o $O/${TAG}_cfg3_full python tools/prof_stage.py 512 naive,c1,erode,hist 2 > $O/${TAG}_cfg3_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"flood_round" -c 1 -f -o $O/${TAG}_flood_full python tools/prof_flood1.py 1 > $O/${TAG}_flood_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"voxelize_brick" -s 4 -c 1 -f -o $O/${TAG}_vox_full python tools/prof_voxelize.py 512 > $O/${TAG}_vox_full.log 2>&1
ncu -i $O/${TAG}_vox_full.ncu-rep --page raw --csv > $O/${TAG}_vox_full_raw.csv 2>/dev/null
ncu -i $O/${TAG}_cfg3_full.ncu-rep --page raw --csv > $O/${TAG}_cfg3_full_raw.csv 2>/dev/null
ncu -i $O/${TAG}_flood_full.ncu-rep --page raw --csv > $O/${TAG}_flood_full_raw.csv 2>/dev/null
tail -3 $O/${TAG}_gpu_tests.log; python tools/show_bench.py $O/${TAG}_bench.json; tail -c 600 $O/${TAG}_bench_reference.json; cat $O/${TAG}_flood_timings.txt
