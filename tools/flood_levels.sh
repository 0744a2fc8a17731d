#!/bin/bash
# sweep of the flood's per-round distance window (VF_FLOOD_LEVELS) on the cfg2 vessel and a dense 256^3 grid
for k in 8 16 24 32 48 100000; do echo "== levels $k"; VF_FLOOD_LEVELS=$k python tools/prof_flood.py 256 256 2>&1 | grep "ms rounds"; done
