#!/bin/bash
# Final check of the build with the thin-front flood solver (default limit 8192, batch contexts on tiles): GPU tests, bench line, flood timings.
set -u
TAG=${1:-r2k}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/${TAG}_gpu_tests.log 2>&1; echo "pytest rc $?" >> $O/${TAG}_gpu_tests.log
tail -3 $O/${TAG}_gpu_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python tools/show_bench.py $O/${TAG}_bench.json
(timeout 100 python tools/prof_flood.py 256 256; timeout 100 python tools/prof_flood.py 256 256 0) > $O/${TAG}_flood_timings.txt 2>&1
cat $O/${TAG}_flood_timings.txt
