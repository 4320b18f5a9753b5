#!/bin/bash
# Round-2 GPU pass A: parity tests, bench (graph + eager + configs), the two reference arms.
set -u
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt; numactl -H >> $OUT/gpu.txt 2>&1; nvidia-smi topo -m >> $OUT/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -15 $OUT/pytest.log
python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -5 $OUT/bench.err
python bench.py --impl reference-gpu --steps 3 --warmup 3 > $OUT/reference_gpu.json 2> $OUT/reference_gpu.err; echo "ref-gpu exit $?"
cat $OUT/reference_gpu.json; tail -3 $OUT/reference_gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/reference_cpu.json 2> $OUT/reference_cpu.err; echo "ref-cpu exit $?"
python -c "
import json
d=json.load(open('$OUT/bench.json'))
for k in ('value','ms_per_step','step_ms','eager','e2e','pcie','roofline','clocks','gpu_launches_per_step','checks','cpu_baseline','placement'):
    print(k, json.dumps(d.get(k)))
print('kernels', json.dumps(d['kernels']))
print('configs', json.dumps(d['configs']))
"
