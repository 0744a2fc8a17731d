#!/bin/bash
# ncu launch list of one C1 + erode pipeline at 512^3 (per-kernel durations)
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_c1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_c1.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:60]].append(float(r[-1].replace(",", "")) / 1e3)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:62s} n={len(v):3d} avg={sum(v)/len(v):8.1f} us")
PY
