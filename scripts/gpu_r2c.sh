#!/bin/bash
# round 2: pre-pass (prepare_kernel) parity + A/B against in-kernel initialisation
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_gpu.log
grep -E "config|full frame|passed|failed|rc=|Error|assert" gpurun_out/r2c_pytest_gpu.log | head -40
for rep in 1 2 3; do for v in 0 1; do
  BHG_PREP=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/r2c_bench_prep$v.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2c_bench_prep$v.json').read().strip().splitlines()[-1])
print('PREP=$v', d['kernel_ms'], 'step ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cam', d['e2e_camera']['dir_and_status']['value'])
PY
done; done
