#!/bin/bash
# ncu evidence for the current build: launch list + full capture of the parity and plane kernels
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -o gpurun_out/prof_${TAG}_parity python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -o gpurun_out/prof_${TAG}_plane python bench.py --steps 2 --warmup 3 --mode plane --no-cpu-baseline > gpurun_out/ncu_full_plane.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
ls -la gpurun_out
