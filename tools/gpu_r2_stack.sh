#!/bin/bash
# A/B of the pair kernels + timeline traces + parity tests of the stack.  tools/gpu_r2_stack.sh <tag> [env settings to sweep...]
set -u
OUT=gpurun_out/${1:-r2d}; mkdir -p $OUT; shift
for cfg in "EQB_TC_PAIR2=0" "$@"; do
  env $cfg timeout 120 python tools/bench_stack.py 2>&1 | tail -1 | tee -a $OUT/ab.txt
done
timeout 120 python tools/trace_stack.py 10 3 > $OUT/trace_p2.txt 2>&1
head -2 $OUT/trace_p2.txt | tail -1
timeout 600 python -m pytest tests -m gpu -x -q -k "stack or tcgen05 or golden or capture or smoke or batch_mates" 2>&1 | tail -3 | tee $OUT/pytest.log
python -c "
from equiadapt_b200 import native
import ctypes
o=(ctypes.c_int*5)(); print('stall', native.lib().eqb_debug_last_stall(o), list(o))"
