#!/bin/bash
mkdir -p gpurun_out
for c in 131072 262144 524288 1048576 1310720 2621440; do
  BHG_CHUNK_RAYS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chunk$c.json 2>> gpurun_out/bench.err
  python - $c <<'PY'
import json,sys
c=sys.argv[1]; j=json.load(open(f'gpurun_out/bench_chunk{c}.json'))
print(c, 'e2e %.0f M' % (j['e2e']['value']/1e6), ' cam_all %.0f  cam_dir %.0f  cam_uv %.0f' % tuple(j['e2e_camera'][k]['value']/1e6 for k in ('all_outputs','dir_and_status','sky_uv_and_status')), 'kernel %.3f ms' % j['kernel_ms']['mean'])
PY
done
