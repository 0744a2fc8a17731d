#!/bin/bash
# batch generation (cfg4) on one GPU under the flood modes: models/s and host CPU seconds per model
O=gpurun_out
for m in 0 1 2; do
  timeout 300 python bench.py --workload batch --meshes ${1:-128} --warmup 2 --flood-mode $m > $O/r2k_batch_mode$m.json 2> $O/r2k_batch_mode$m.err
  python - <<PY
import json
d=json.loads(open("$O/r2k_batch_mode$m.json").read().strip().splitlines()[-1])
print("mode $m", round(d["value"],1), "models/s", {k:v for k,v in d["config"].items() if k!="workload"})
PY
done
