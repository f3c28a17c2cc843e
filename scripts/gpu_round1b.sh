#!/bin/bash
# second GPU contact: Nystrom-form kernel; parity tests, register-cap variants, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E 'selftest|dpos|full frame|passed|failed|rc=' gpurun_out/pytest_gpu.log
V=blackhole_geodesic_calculator_b200/lib/variants
for mb in 3 4 5; do
  BHG_LIB=$PWD/$V/libbhgeo_mb$mb.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mb$mb.json 2>> gpurun_out/bench.err
  echo "mb$mb: $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench_mb$mb.json) $(grep -o '"frac": [0-9.e-]*' gpurun_out/bench_mb$mb.json)"
done
for T in 16 24; do timeout 300 python bench.py --steps 5 --warmup 3 --threshold $T --no-cpu-baseline > gpurun_out/bench_T$T.json 2>> gpurun_out/bench.err; echo "T$T: $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench_T$T.json)"; done
timeout 300 python bench.py --steps 5 --warmup 3 --mode plane --no-cpu-baseline > gpurun_out/bench_plane.json 2>> gpurun_out/bench.err
echo "plane: $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench_plane.json)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -o gpurun_out/prof_r1b python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/bench.err
ls -la gpurun_out | head -30
