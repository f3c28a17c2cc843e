#!/bin/bash
# 8-GPU evidence: bench at N (weak value, strong_frame, animation_100, e2e legs) + concurrent PCIe probe
N=${1:-8}
mkdir -p gpurun_out
TAG=${2:-r2k}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
tail -2 gpurun_out/${TAG}_scale_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/pcie_concurrent.py > gpurun_out/${TAG}_pcie_concurrent_n$N.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_scale_n$N.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'cam', {k:(round(v['value']/1e9,3) if isinstance(v,dict) else '') for k,v in d['e2e_camera'].items()}, 'f32io', d['e2e_f32io']['value'])
s=d.get('strong_frame',{}); print('strong', {k:s.get(k) for k in ('route','ms','efficiency_vs_n1','peer_equals_single','gather_equals_single','shard_compute_only_ms','error')}, {k:round(v['ms_median'],3) for k,v in s.get('ms_by_route',{}).items()})
print('anim', d.get('animation_100'))
print(open('gpurun_out/${TAG}_pcie_concurrent_n$N.json').read()[:1500])
PY
