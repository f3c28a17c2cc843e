#!/bin/bash
# round 2, first GPU call: issue-model microbenchmark, GPU half of the outlier adjudication, baseline tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_gpu.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/issue_model_bench scripts/issue_model_bench.cu && timeout 300 /tmp/issue_model_bench > gpurun_out/r2a_issue_model.txt 2>&1
cat gpurun_out/r2a_issue_model.txt
timeout 900 python scripts/adjudicate_gpu.py > gpurun_out/r2a_adjudicate_gpu.log 2>&1
tail -40 gpurun_out/r2a_adjudicate_gpu.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2a_pytest_gpu.log
