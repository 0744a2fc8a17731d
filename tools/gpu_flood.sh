#!/bin/bash
# flood iteration check: TAG; flood / slab / dataset tests, flood timings (cfg2 vessel + dense 256^3)
TAG=$1; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "flood or slab or dataset or cfg2 or cfg4 or cfg5 or round_loop" > $O/${TAG}_tests.log 2>&1; echo "pytest rc $?" >> $O/${TAG}_tests.log
timeout 200 python tools/prof_flood.py 256 256 > $O/${TAG}_flood_timings.txt 2>&1
tail -5 $O/${TAG}_tests.log; cat $O/${TAG}_flood_timings.txt
