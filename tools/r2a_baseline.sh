#!/bin/bash
# round-2 first call: the round-1 kernels under the new full-size parity tests and the reworked bench line; experimental C1 probe
O=gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > $O/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_gpu_tests.log 2>&1; echo "pytest rc $?" >> $O/r2a_gpu_tests.log
VF_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_flood_gpu.py -m gpu -q -k "descent" > $O/r2a_descent_test.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err
VF_C1_DESCENT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-batch --no-vessel --no-cpu-baseline > $O/r2a_bench_descent.json 2> $O/r2a_bench_descent.err
timeout 120 python tools/prof_flood.py 256 256 > $O/r2a_flood_timings.txt 2>&1
tail -3 $O/r2a_gpu_tests.log; tail -3 $O/r2a_descent_test.log; cat $O/r2a_flood_timings.txt
