#!/bin/bash
# round-2 call b: C1 by descent certificate as the default path — tests, bench, launch list of the cfg3 stages
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2b_gpu_tests.log 2>&1; echo "pytest rc $?" >> $O/r2b_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-batch > $O/r2b_bench.json 2> $O/r2b_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2b_launches_stage.csv python tools/prof_stage.py 512 naive,c1,erode,hist 2 > $O/r2b_prof_stage.log 2>&1
tail -5 $O/r2b_gpu_tests.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stage_ms"], d.get("parity"))
PY
grep -v "^==" $O/r2b_launches_stage.csv | awk -F'","' 'NR>1{print $5, $NF}' | tail -40
