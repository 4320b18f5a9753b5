#!/bin/bash
OUT=gpurun_out/${1:-r2g}; mkdir -p $OUT
timeout 60 tools/tmem_probe | tee $OUT/tmem_probe.txt
timeout 120 python tools/bench_stack.py 2>&1 | tail -1 | tee $OUT/stack.txt
