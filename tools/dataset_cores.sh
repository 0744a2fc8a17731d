#!/bin/bash
# the native dataset driver (one C call per model) without export on one GPU, all cores / 4 cores
O=gpurun_out
run() { local pre=""; [ -n "$2" ] && pre="taskset -c $2"
  timeout 300 $pre python bench.py --workload dataset --no-export --meshes 128 --mesh-pool 128 --warmup 2 --jobs $3 > $O/dc_$1.json 2> $O/dc_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/dc_$1.json").read().strip().splitlines()[-1])
    c=d["config"]; print("$1:", round(d["value"],1), "models/s", "cpu_s/model", round(c.get("host_cpu_s_per_model_rank0",0),4))
except Exception as e: print("$1 failed", e, open("$O/dc_$1.err").read()[-500:])
PY
}
run all_j16 "" 16
run all_j32 "" 32
run c4_j16 0-3 16
run c4_j32 0-3 32
run c4_j8 0-3 8
