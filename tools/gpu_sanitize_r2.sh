#!/bin/bash
# Round 2: compute-sanitizer memcheck over the whole small GPU suite (new kernels: pair2 stack, TMA crop+resize, fused finish/select,
# geometry table), racecheck over the new shared-memory kernels that are not mbarrier pipelines.
mkdir -p gpurun_out/sanitize_r2
SKIP='not full_size and not full_batch'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SKIP" > gpurun_out/sanitize_r2/memcheck.log 2>&1; echo "memcheck exit $?"
tail -4 gpurun_out/sanitize_r2/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "tma_resize or select_fused or batch_mates" > gpurun_out/sanitize_r2/racecheck.log 2>&1; echo "racecheck exit $?"
tail -6 gpurun_out/sanitize_r2/racecheck.log
