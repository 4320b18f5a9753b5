#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2h}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest.log
timeout 120 python tools/bench_resize.py 2>&1 | tail -5 | tee $OUT/resize.txt
timeout 120 python tools/bench_stack.py 2>&1 | tail -1 | tee $OUT/stack.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python -c "
import json
d=json.load(open('$OUT/bench.json'))
for k in ('value','ms_per_step','step_ms','gpu_launches_per_step','checks'):
    print(k, json.dumps(d.get(k)))
print('eager', d['eager']['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_pcie_bound'])
print('kernels', {k:round(v['avg_us'],1) for k,v in d['kernels'].items()})
print('rooflines', {k:round(v['frac'],3) for k,v in d['rooflines'].items()})
"
