#!/bin/bash
# Quick GPU pass: parity tests + bench line (no ncu).  tools/gpu_pass2.sh <tag> [pytest -k expr]
set -u
TAG=${1:-pass}; KEXPR=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
else
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
fi
tail -15 $OUT/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -5 $OUT/bench.err
cat $OUT/bench.json | head -c 4000; echo
