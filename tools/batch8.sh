#!/bin/bash
# cfg4 batch on 8 GPUs: wait modes / jobs per GPU side by side on the same box
O=gpurun_out
for spec in "$@"; do
  set -- $(echo $spec | tr , ' ')
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --workload batch --meshes 1024 --mesh-pool 0 --warmup 2 --jobs $2 --blocking-sync $3 > $O/b8_$1.json 2> $O/b8_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/b8_$1.json").read().strip().splitlines()[-1])
    c=d["config"]; print("$1:", round(d["value"],1), "models/s", "cpu_s/model", round(c.get("host_cpu_s_per_model_rank0",0),4))
except Exception as e: print("$1 failed", e, open("$O/b8_$1.err").read()[-800:])
PY
done
