#!/bin/bash
# round 2: full GPU test-suite with the conditioning rule + README figure pin + overlay images + baseline bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log
grep -E "config|full frame|fig5|cfg[0-9]:|passed|failed|rc=|Error|assert" gpurun_out/r2b_pytest_gpu.log | head -40
timeout 300 python scripts/readme_overlay.py 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -2 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print(d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])
PY
