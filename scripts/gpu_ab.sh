#!/bin/bash
# A/B/... library builds on the same box, interleaved: lib/variants/libbhgeo_<name>.so for each name given
V=$PWD/blackhole_geodesic_calculator_b200/lib/variants
NAMES=${NAMES:-"A B"}
for rep in 1 2 3; do for v in $NAMES; do
  BHG_LIB=$V/libbhgeo_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline $@ 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('$v', 'kernel %.4f ms' % j['kernel_ms']['mean'])"
done; done
