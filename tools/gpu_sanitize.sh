#!/bin/bash
# compute-sanitizer passes over the small GPU parity tests (memcheck; racecheck on the shared-memory pipelines)
mkdir -p gpurun_out/sanitize
SKIP='not full_size and not full_batch'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SKIP" > gpurun_out/sanitize/memcheck.log 2>&1; echo "memcheck exit $?"
tail -4 gpurun_out/sanitize/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "vnsmall or vndeepsets or escnn_expanded or group_pool or frames" > gpurun_out/sanitize/racecheck.log 2>&1; echo "racecheck exit $?"
tail -4 gpurun_out/sanitize/racecheck.log
