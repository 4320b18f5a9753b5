#!/bin/bash
OUT=gpurun_out/${1:-r2y}; mkdir -p $OUT
for cfg in "64 0" "64 1" "32 0" "32 1" "128 1" "16 0"; do
  set -- $cfg
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 8 --e2e-shard $1 --e2e-ramp $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('shard $1 ramp $2: e2e', round(e['value']), 'frac', round(e['frac_of_pcie_bound'],3), 'step_ms', {k:round(v,2) for k,v in e['step_ms'].items()}, 'bound', round(d['pcie']['bound_img_s']))" | tee -a $OUT/e2e.txt
done
