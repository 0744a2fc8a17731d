#!/bin/bash
# Experimental CUDA-graph round loop of the flood (VF_FLOOD_GRAPH=1, csrc/flood.cu run_rounds_graph): parity first, then what it buys.
# Run on the GPU box through gpurun; outputs land in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
# the graph loop is compiled only on request (its device-side cudaGraphSetConditional is resolved by the driver at module load)
VF_NVCC_EXTRA=-DVF_FLOOD_GRAPH_BUILD python build_lib.py --force > $O/graph_build.log 2>&1 || { tail -5 $O/graph_build.log; exit 1; }
VF_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_flood_gpu.py -m gpu -q -k graph_driven > $O/graph_parity.log 2>&1
tail -3 $O/graph_parity.log
grep -q passed $O/graph_parity.log || exit 1
N=${1:-128}
for mode in 0 1; do
  VF_FLOOD_GRAPH=$mode python tools/prof_cfg12.py > $O/graph${mode}_cfg12.txt 2>&1
  VF_FLOOD_GRAPH=$mode python tools/prof_flood.py 256 256 > $O/graph${mode}_flood_timings.txt 2>&1
  VF_FLOOD_GRAPH=$mode python bench.py --workload batch --meshes $N --warmup 2 > $O/graph${mode}_batch.json 2> $O/graph${mode}_batch.err
  VF_FLOOD_GRAPH=$mode taskset -c 0-3 python bench.py --workload batch --meshes $N --warmup 2 > $O/graph${mode}_batch_4c.json 2> $O/graph${mode}_batch_4c.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/graph*_batch*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["value"], 1), "models/s", "cpu s/model", round(d["config"].get("host_cpu_s_per_model_rank0") or 0, 4))
    except Exception as e:
        print(f, "unreadable:", e)
PY
