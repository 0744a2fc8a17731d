#!/bin/bash
# 8-GPU runs of the three workloads (one process per GPU under torchrun); outputs in gpurun_out/
O=gpurun_out; mkdir -p $O
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/scale${N}_cfg3.json 2> $O/scale${N}_cfg3.err
timeout 300 $TR --master-port 29512 bench.py --gpus $N --workload batch --meshes $((64 * N)) --steps 1 --warmup 1 > $O/scale${N}_batch.json 2> $O/scale${N}_batch.err
timeout 400 $TR --master-port 29513 bench.py --gpus $N --workload slab --size 2048 --seeds 256 --steps 2 --warmup 1 > $O/scale${N}_slab.json 2> $O/scale${N}_slab.err
tail -c 300 $O/scale${N}_cfg3.err $O/scale${N}_batch.err $O/scale${N}_slab.err
wc -c $O/scale${N}_*.json
