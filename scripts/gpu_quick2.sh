#!/bin/bash
# quick feedback: hot-loop probe, headline kernel time, core parity tests
mkdir -p gpurun_out
python scripts/hot_loop_probe.py
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('kernel %.4f ms frac %.4f e2e %.3g' % (j['kernel_ms']['mean'], j['roofline']['frac'], j['e2e']['value']))"
if [ "$1" != "notest" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
fi
