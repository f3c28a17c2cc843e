#!/bin/bash
# first GPU contact: parity tests, bench, threshold sweep, launch list, one full ncu capture
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_parity.json 2> gpurun_out/bench_parity.err; tail -c 3000 gpurun_out/bench_parity.json
for T in 8 16 24; do timeout 300 python bench.py --steps 5 --warmup 3 --threshold $T --no-cpu-baseline > gpurun_out/bench_T$T.json 2>> gpurun_out/bench_parity.err; done
timeout 300 python bench.py --steps 5 --warmup 3 --mode plane --no-cpu-baseline > gpurun_out/bench_plane.json 2>> gpurun_out/bench_parity.err
grep -h -o '"value": [0-9.e+]*\|"kernel_ms": {[^}]*}\|"frac": [0-9.e-]*' gpurun_out/bench_T*.json gpurun_out/bench_plane.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -o gpurun_out/prof_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
