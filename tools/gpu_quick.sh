#!/bin/bash
# quick GPU check used while iterating: TAG, pytest -k selection (or "all"), stages to launch-list at 512^3
TAG=$1; SEL=${2:-all}; STAGES=${3:-naive,c1,erode,hist}
O=gpurun_out
if [ "$SEL" = all ]; then timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; else timeout 900 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/${TAG}_tests.log 2>&1; fi
echo "pytest rc $?" >> $O/${TAG}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-batch --no-vessel --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv python tools/prof_stage.py 512 $STAGES 2 > $O/${TAG}_prof.log 2>&1
tail -4 $O/${TAG}_tests.log
python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["stage_ms"], d.get("parity_checked"), (d.get("parity") or {}).get("mismatching_cells"))
except Exception as e: print("bench parse failed", e); print(open("$O/${TAG}_bench.err").read()[-2000:])
PY
grep -v "^==" $O/${TAG}_launches.csv | awk -F'","' 'NR>1{n=split($5,a,"("); print a[1], $NF}' | tail -16
