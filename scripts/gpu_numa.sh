#!/bin/bash
# host-buffer e2e at N ranks with and without NUMA placement of the pinned frame buffers
N=${1:-2}
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -i "numa\|^CPU(s)\|Model name\|Socket"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa_node=$(cat $d/numa_node)"; fi; done; nproc; python -c "import os; print(sorted(os.sched_getaffinity(0)))"; } > gpurun_out/numa_topo.txt 2>&1
for bind in 0 1; do
  BHG_NUMA_BIND=$bind timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$bind bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/numa_bind${bind}_n$N.json 2>> gpurun_out/numa.err
  echo "bind=$bind: $(grep -o '"e2e": {"value": [0-9.e+]*' gpurun_out/numa_bind${bind}_n$N.json) $(grep -o '"e2e_f32io": {"value": [0-9.e+]*' gpurun_out/numa_bind${bind}_n$N.json)"
done
head -40 gpurun_out/numa_topo.txt
