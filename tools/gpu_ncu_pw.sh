#!/bin/bash
# ncu --set full of the training tensor-core kernels (tools/check_pw.py drives them).  tools/gpu_ncu_pw.sh <tag>
set -u
OUT=gpurun_out/${1:-r3v}; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pw_conv_kernel|pw_wgrad_kernel" -s 30 -c 8 -f -o $OUT/full_pw \
    python tools/check_pw.py > $OUT/ncu_pw.log 2>&1
echo "exit $?"; tail -3 $OUT/ncu_pw.log
