#!/bin/bash
# end-to-end producer settings at one GPU: spinning / blocking waits, 4 / 6 / 8 producer threads
O=gpurun_out
for cfg in "off 4" "on 4" "on 6" "on 8" "off 8"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 3 --no-batch --no-vessel --no-cpu-baseline --e2e-blocking $1 --e2e-producers $2 > $O/e2e_$1_$2.json 2> $O/e2e_$1_$2.err
  python - <<PY
import json
d=json.loads(open("$O/e2e_$1_$2.json").read().strip().splitlines()[-1])
e=d["e2e"]; print("blocking $1 producers $2:", round(e["value"],1), "Gvoxels/s", [round(x,3) for x in e["passes_ms_per_step"]], "step", round(d["ms_per_step"],4))
PY
done
