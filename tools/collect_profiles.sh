#!/bin/bash
# Collects the single-GPU evidence that profiles/ summarises (run on the GPU box through gpurun; outputs land in gpurun_out/).
#   bench line (with cpu_baseline and the short batch measurement), reference arm, batch / dataset / flood / voxelizer timings,
#   ncu launch list, ncu --set full of the pipeline kernels, the flood round kernel and the solid voxelizer
set -u
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --workload batch --meshes 256 --warmup 2 > $O/bench_batch.json 2> $O/bench_batch.err
python bench.py --workload batch --meshes 64 --warmup 2 --jobs 1 > $O/bench_batch_1job.json 2> $O/bench_batch_1job.err
python bench.py --workload dataset --meshes 96 --warmup 2 > $O/bench_dataset.json 2> $O/bench_dataset.err
python tools/prof_flood.py 256 256 > $O/flood_timings.txt 2>&1
python tools/prof_solid.py 128 256 512 > $O/voxelizer_timings.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 600 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batch > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"naive_brick|stencil_fast|ccl_tile|ccl_border|ccl_select|histogram|pointwise" -s 14 -c 14 -o $O/cfg3_full python tools/prof_stage.py 512 naive,c1,erode,hist 2 > $O/cfg3_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"flood_round" -s 6 -c 2 -o $O/flood_full python tools/prof_flood1.py 1 > $O/flood_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"solid_scatter|solid_expand|tetra_setup|voxelize_brick" -s 8 -c 4 -o $O/voxelizer_full python tools/prof_solid.py 256 > $O/voxelizer_full.log 2>&1
ls -la $O | tail -20
