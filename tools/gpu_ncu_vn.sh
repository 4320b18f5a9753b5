#!/bin/bash
# ncu --set full of the VNSmall kernel inside bench.py's cfg4 leg.  tools/gpu_ncu_vn.sh <tag>
set -u
OUT=gpurun_out/${1:-r3o}; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vnsmall_kernel -s 3 -c 1 -f -o $OUT/full_vnsmall \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $OUT/ncu_vn.log 2>&1
echo "exit $?"; tail -3 $OUT/ncu_vn.log
