#!/bin/bash
# strong scaling of one frame incl. delivery to the owner (scripts/strong_frame.py) at N=1 and each N given, plus the NCCL/IPC test
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_adapters.py -m gpu -x -q -k nccl > gpurun_out/pytest_strong.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_strong.log
tail -3 gpurun_out/pytest_strong.log
timeout 300 python scripts/strong_frame.py > gpurun_out/strong_n1.json 2> gpurun_out/strong.err
cat gpurun_out/strong_n1.json
for N in "${@:-2}"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N scripts/strong_frame.py > gpurun_out/strong_n$N.json 2>> gpurun_out/strong.err
  cat gpurun_out/strong_n$N.json
done
tail -5 gpurun_out/strong.err
