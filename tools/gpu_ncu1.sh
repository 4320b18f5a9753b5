#!/bin/bash
# ncu --set full of ONE kernel (regex) in a small driver script.  tools/gpu_ncu1.sh <tag> <kernel-regex> <python script + args...>
set -u
TAG=$1; K=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
SAFE=$(echo $K | tr -c 'A-Za-z0-9_' '_')
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/full_$SAFE python "$@" > $OUT/ncu_$SAFE.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_$SAFE.log; ls -la $OUT
