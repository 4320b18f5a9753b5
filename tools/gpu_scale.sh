#!/bin/bash
# bench.py at N GPUs of one box, launched the way the driver launches it.  tools/gpu_scale.sh <tag> <N> [extra bench args]
set -u
TAG=$1; N=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt
if [ "$N" = "1" ]; then
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 "$@" > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 "$@" > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
fi
echo "bench N=$N exit $?"; tail -5 $OUT/bench_n$N.err
python -c "
import json
d=json.loads([l for l in open('$OUT/bench_n$N.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','step_ms','eager','clocks','placement','checks'):
    print(k, json.dumps(d.get(k)))
e=d['e2e']; print('e2e', e['value'], e['frac_of_pcie_bound'], e['step_ms']); print('pcie', d['pcie'])
print('kernels', {k:round(v['avg_us'],1) for k,v in d['kernels'].items()})
c=d['configs']
for k in ('cfg3_opt_d4_images','cfg4_pointcloud_so3','cfg5_nbody_e3'):
    print(k, {t:(round(c[k][t]['samples_per_s']), round(c[k][t]['us'],1), c[k][t]['mode']) for t in ('strong','weak')})
"
