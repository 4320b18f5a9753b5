#!/bin/bash
# Round-2 evidence pass: bench line, ncu launch list of the same command, --set full captures of the bench kernels, an
# NVTX-filtered launch list (the ranges around the C-ABI calls exist), reference arms.
set -u
TAG=${1:-r2w}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -2 $OUT/bench.err
timeout 300 python bench.py --impl reference-gpu --steps 3 --warmup 3 > $OUT/reference_gpu.json 2> $OUT/reference_gpu.err; echo "ref-gpu exit $?"; cat $OUT/reference_gpu.json | head -c 600; echo; tail -2 $OUT/reference_gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/reference_cpu.json 2> $OUT/reference_cpu.err; echo "ref-cpu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-graph > $OUT/ncu_list.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --nvtx --nvtx-include "eqb_warp_canonicalize/" --metrics gpu__time_duration.sum --clock-control none -c 6 --csv --log-file $OUT/nvtx_warp_canonicalize.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-graph > $OUT/ncu_nvtx.log 2>&1; echo "nvtx list exit $?"
for K in gconv_stack_pair2 resample_tma crop_resize_tma; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o $OUT/full_$K \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-graph > $OUT/ncu_full_$K.log 2>&1
  echo "ncu full $K exit $?"
done
ls -la $OUT
