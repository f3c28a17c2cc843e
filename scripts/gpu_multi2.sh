#!/bin/bash
# round 2 multi-GPU: 2-GPU tests (NCCL gather, peer frame with flags) + bench at N (strong_frame section)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
if [ "$2" != "notest" ]; then
timeout 900 python -m pytest tests/test_gpu_adapters.py -m gpu -x -q > gpurun_out/r2f_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_multi.log
tail -15 gpurun_out/r2f_pytest_multi.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2f_scale_n$N.json 2> gpurun_out/r2f_scale_n$N.err
tail -3 gpurun_out/r2f_scale_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_scale_n$N.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('strong_frame'), indent=1))
PY
