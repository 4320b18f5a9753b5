#!/bin/bash
# A/B of the pair-kernel pipeline knobs + timeline traces + parity tests of the stack.
set -u
OUT=gpurun_out/${1:-r2d}; mkdir -p $OUT
for cfg in "EQB_TC_LIFT_ORDER=0 EQB_TC_EPI2_PIPE=0" "EQB_TC_LIFT_ORDER=0 EQB_TC_EPI2_PIPE=1" "EQB_TC_LIFT_ORDER=1 EQB_TC_EPI2_PIPE=0" "EQB_TC_LIFT_ORDER=1 EQB_TC_EPI2_PIPE=1"; do
  env $cfg timeout 120 python tools/bench_stack.py 2>&1 | tail -1 | tee -a $OUT/ab.txt
done
EQB_TC_LIFT_ORDER=0 EQB_TC_EPI2_PIPE=0 timeout 120 python tools/trace_stack.py 10 3 > $OUT/trace_00.txt 2>&1
EQB_TC_LIFT_ORDER=1 EQB_TC_EPI2_PIPE=0 timeout 120 python tools/trace_stack.py 10 3 > $OUT/trace_10.txt 2>&1
EQB_TC_LIFT_ORDER=1 EQB_TC_EPI2_PIPE=1 timeout 120 python tools/trace_stack.py 10 3 > $OUT/trace_11.txt 2>&1
head -2 $OUT/trace_00.txt | tail -1; head -2 $OUT/trace_10.txt | tail -1; head -2 $OUT/trace_11.txt | tail -1
timeout 600 python -m pytest tests -m gpu -x -q -k "stack or tcgen05 or golden or capture or smoke" 2>&1 | tail -2 | tee $OUT/pytest.log
