#!/bin/bash
# batch generation (cfg4) on one GPU with the host share of one rank of the 8-GPU box (4 hardware threads): models/s per setting
O=gpurun_out
run() { # label, taskset cores ("" = all), jobs, wait mode
  local pre=""; [ -n "$2" ] && pre="taskset -c $2"
  timeout 300 $pre python bench.py --workload batch --meshes 128 --mesh-pool 0 --warmup 2 --jobs $3 --blocking-sync $4 > $O/bc_$1.json 2> $O/bc_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bc_$1.json").read().strip().splitlines()[-1])
    c=d["config"]; print("$1:", round(d["value"],1), "models/s", "cpu_s/model", round(c.get("host_cpu_s_per_model_rank0",0),4))
except Exception as e: print("$1 failed", e, open("$O/bc_$1.err").read()[-500:])
PY
}
for spec in "$@"; do set -- $(echo $spec | tr , ' '); run $1 "$2" $3 $4; done
