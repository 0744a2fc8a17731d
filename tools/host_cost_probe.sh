#!/bin/bash
# Host-side cost of batch generation (run on the GPU box through gpurun; outputs land in gpurun_out/).
# The 8-GPU batch figure is bound by the host cores per rank (4 on the 8-GPU box): this probe restricts one rank to 4 / 8 / all cores and
# compares the Python-threaded producer (`--workload batch`) with the native driver loop (`--workload dataset --no-export`, one C call per
# model); both lines carry host_cpu_s_per_model_rank0.
set -u
O=gpurun_out
mkdir -p $O
N=${1:-128}
ALL=$(nproc)
for cores in 4 8 $ALL; do
  [ "$cores" -gt "$ALL" ] && continue
  taskset -c 0-$((cores-1)) python bench.py --workload batch --meshes $N --warmup 2 > $O/host_batch_py_${cores}c.json 2> $O/host_batch_py_${cores}c.err
  taskset -c 0-$((cores-1)) python bench.py --workload dataset --no-export --meshes $N --warmup 2 > $O/host_batch_native_${cores}c.json 2> $O/host_batch_native_${cores}c.err
done
# CTAs per SM of a flood round launch (default 4): fewer idle CTAs when 16 jobs share the GPU?
for k in 1 2; do
  VF_FLOOD_CTAS_PER_SM=$k python bench.py --workload batch --meshes $N --warmup 2 > $O/host_batch_py_ctas${k}.json 2> $O/host_batch_py_ctas${k}.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/host_batch_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["value"], 1), "models/s", d["config"].get("jobs_per_gpu"), "jobs", "cpu s/model", round(d["config"].get("host_cpu_s_per_model_rank0") or 0, 4))
    except Exception as e:
        print(f, "unreadable:", e)
PY
