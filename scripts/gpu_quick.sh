#!/bin/bash
# quick GPU check: parity tests, then one bench run per argument; each argument is a string of extra
# bench.py flags (e.g. "--no-tiles" or "--mode plane"); prefix "lib=<variant>;" selects lib/variants/libbhgeo_<variant>.so
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E 'selftest|dpos|full frame|passed|failed|rc=|Error|error' gpurun_out/pytest_gpu.log | head -30
V=$PWD/blackhole_geodesic_calculator_b200/lib/variants
i=0
for spec in "$@"; do
  i=$((i+1)); lib=""; flags="$spec"
  if [[ "$spec" == lib=* ]]; then lib="${spec%%;*}"; lib="${lib#lib=}"; flags="${spec#*;}"; fi
  if [ -n "$lib" ]; then export BHG_LIB=$V/libbhgeo_$lib.so; else unset BHG_LIB; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $flags > gpurun_out/bench_q$i.json 2>> gpurun_out/bench.err
  echo "[$spec] $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench_q$i.json) $(grep -o '"frac": [0-9.e-]*' gpurun_out/bench_q$i.json) $(grep -o '"e2e": {"value": [0-9.e+]*' gpurun_out/bench_q$i.json)"
done
tail -3 gpurun_out/bench.err
