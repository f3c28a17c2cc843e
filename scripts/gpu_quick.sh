#!/bin/bash
# quick GPU check: parity tests + bench of the lib variants given as arguments (default: in-tree build)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E 'selftest|dpos|full frame|passed|failed|rc=|Error|error' gpurun_out/pytest_gpu.log | head -30
V=blackhole_geodesic_calculator_b200/lib/variants
for v in "$@"; do
  BHG_LIB=$PWD/$V/libbhgeo_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2>> gpurun_out/bench.err
  echo "$v: $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench_$v.json) $(grep -o '"frac": [0-9.e-]*' gpurun_out/bench_$v.json)"
  BHG_LIB=$PWD/$V/libbhgeo_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --mode plane --no-cpu-baseline > gpurun_out/bench_plane_$v.json 2>> gpurun_out/bench.err
  echo "$v plane: $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench_plane_$v.json)"
done
tail -3 gpurun_out/bench.err
