#!/bin/bash
OUT=gpurun_out/${1:-r2e}; mkdir -p $OUT
EQB_TC_LIFT_ORDER=0 EQB_TC_EPI2_PIPE=0 timeout 120 python tools/trace_stack.py 10 3 > $OUT/trace_base.txt 2>&1
EQB_TC_LIFT_ORDER=1 EQB_TC_EPI2_PIPE=1 timeout 120 python tools/trace_stack.py 10 3 > $OUT/trace_new.txt 2>&1
head -3 $OUT/trace_base.txt; head -3 $OUT/trace_new.txt
