#!/bin/bash
# One full ncu capture of a kernel (regex) from a short bench run.  tools/gpu_ncu.sh <tag> <kernel-regex> [skip] [count]
set -u
TAG=$1; K=$2; SKIP=${3:-2}; CNT=${4:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SAFE=$(echo $K | tr -c 'A-Za-z0-9_' '_')
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o $OUT/full_$SAFE \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full_$SAFE.log 2>&1
echo "ncu full $K exit $?"
ls -la $OUT
