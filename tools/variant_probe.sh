#!/bin/bash
# Builds the library with one compile-time experimental variant, runs the GPU parity suite and the default bench line, then restores the
# default build.  usage (through gpurun): tools/variant_probe.sh VF_CCL_JUMP   [more -D names ...]
# Run-time variants need no rebuild: `VF_C1_DESCENT=1 python bench.py ...` (c1_descent.cu: certificate pass + list work instead of the union-find
# for C1; parity: `VF_TEST_EXPERIMENTAL=1 python -m pytest tests/test_flood_gpu.py -m gpu -k descent`).
# Compile-time variants staged in the sources: VF_CCL_JUMP (ccl.cu: pointer jumping instead of per-lane chain walks in the in-tile flatten phase),
# VF_HIST_WARP (stencil.cu: uniform vectors of the histogram counted per warp with one shared atomic per distinct label),
# VF_FLOOD_GRAPH_BUILD (flood.cu: CUDA-graph round loop, additionally needs VF_FLOOD_GRAPH=1 at run time; see tools/graph_loop_probe.sh).
set -u
O=gpurun_out
mkdir -p $O
tag=$(echo "$*" | tr ' ' '_')
defs=""
for d in "$@"; do defs="$defs -D$d"; done
VF_NVCC_EXTRA="$defs" python build_lib.py --force > $O/variant_${tag}_build.log 2>&1 || { tail -5 $O/variant_${tag}_build.log; exit 1; }
timeout 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/variant_${tag}_tests.log 2>&1
tail -2 $O/variant_${tag}_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/variant_${tag}_bench.json 2> $O/variant_${tag}_bench.err
python -c "import json,sys; d=json.load(open('$O/variant_${tag}_bench.json')); print('$tag', round(d['value'],2), d['stage_ms'], round(d['batch']['value'],1))"
python build_lib.py --force > /dev/null 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/variant_default_bench.json 2> /dev/null
python -c "import json,sys; d=json.load(open('$O/variant_default_bench.json')); print('default', round(d['value'],2), d['stage_ms'], round(d['batch']['value'],1))"
