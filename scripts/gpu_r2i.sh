#!/bin/bash
# full GPU suite + N=1 bench (with animation_100, f32 camera leg) + smoke
mkdir -p gpurun_out
TAG=${1:-r2i}
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|rc=|Error|assert" gpurun_out/${TAG}_pytest_gpu.log | head -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('kernel', d['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
print('cam', {k:(v['value'] if isinstance(v,dict) else v) for k,v in d['e2e_camera'].items()})
print('anim', d.get('animation_100'))
print('cpu', {k:v for k,v in d.get('cpu_baseline',{}).items() if k!='sample'})
print('parity', d.get('parity_vs_cpu_sample'))
PY
