#!/bin/bash
# what the driver runs at round end, in one call: gpu tests, smoke, bench (both arms), launch list + full ncu
mkdir -p gpurun_out
TAG=${1:-final}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -4 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -o gpurun_out/${TAG}_prof_parity python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/${TAG}_nvsmi.csv
ls -la gpurun_out | grep $TAG
