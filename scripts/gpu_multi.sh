#!/bin/bash
# multi-GPU check: full gpu test suite (incl. the 2-GPU NCCL gather test) + bench at N=1 and N=$1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_multi.log
tail -5 gpurun_out/pytest_gpu_multi.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2>> gpurun_out/scale.err
for f in gpurun_out/scale_n1.json gpurun_out/scale_n$N.json; do echo "$f: $(grep -o '"value": [0-9.e+]*' $f | head -1) $(grep -o '"ms_per_step": [0-9.e+]*' $f) $(grep -o '"e2e": {"value": [0-9.e+]*' $f)"; done
tail -5 gpurun_out/scale.err
