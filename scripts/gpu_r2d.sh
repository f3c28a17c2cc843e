#!/bin/bash
# launch list + full ncu capture of the pre-pass build
mkdir -p gpurun_out
TAG=${1:-r2d}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
grep -E "trace_kernel|prepare_kernel" gpurun_out/${TAG}_launches.csv | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -o gpurun_out/${TAG}_prof_parity python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prepare_kernel -s 4 -c 1 -o gpurun_out/${TAG}_prof_prepare python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full2.log 2>&1
for b in ; do
BHG_IDLE_BUDGET=$b timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('budget $b', 'kernel %.4f ms' % j['kernel_ms']['mean'])"
done
