#!/bin/bash
# courier SM count sweep on N GPUs (strong_frame section only matters)
N=${1:-2}
for sms in ${SMS:-4 8 16}; do
BHG_COURIER_SMS=$sms timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['strong_frame']
print('courier_sms $sms', 'n1', round(s['n1_ms'],3), {k:round(v['ms_median'],3) for k,v in s['ms_by_route'].items()}, 'compute-only', round(s['shard_compute_only_ms'],3), 'equal', s['peer_equals_single'])"
done
