#!/bin/bash
# does a torchrun job that captured NCCL fences in CUDA graphs exit cleanly? (tight timeout: a hang costs GPU budget)
mkdir -p gpurun_out
BHG_STRONG_GRAPHS=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/strong_frame.py > gpurun_out/graph_teardown.json 2> gpurun_out/graph_teardown.err
echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/graph_teardown.json')); print(d['peer_frame_graph_ms'])"
tail -3 gpurun_out/graph_teardown.err
