#!/bin/bash
# GPU check of the thin-front flood solver: parity tests, timings with and without it (tools/prof_flood.py), memcheck of one small case
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_flood_gpu.py tests/test_cleanup_gpu.py -x -q -m gpu > gpurun_out/front_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/front_tests.log
tail -15 gpurun_out/front_tests.log
for lim in 16384 0 65536 4096; do timeout 120 python tools/prof_flood.py 256 256 $lim; done > gpurun_out/front_timings.txt 2>&1
cat gpurun_out/front_timings.txt
timeout 120 python tools/prof_flood.py 512 128 16384 > gpurun_out/front_timings512.txt 2>&1; timeout 120 python tools/prof_flood.py 512 128 0 >> gpurun_out/front_timings512.txt 2>&1
cat gpurun_out/front_timings512.txt
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_flood_gpu.py -x -q -m gpu -k "front_limit and 40-4" > gpurun_out/front_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/front_memcheck.log
tail -12 gpurun_out/front_memcheck.log
