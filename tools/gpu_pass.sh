#!/bin/bash
# One GPU pass (run under gpurun from the repo root): parity tests, bench line, ncu launch list, full captures.
#   tools/gpu_pass.sh <tag> [kernel-regex-for-full-capture ...]
set -u
TAG=${1:-pass}; shift || true
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -5 $OUT/pytest.log
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json | head -c 3000; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_list.log 2>&1
for K in "$@"; do
  SAFE=$(echo $K | tr -c 'A-Za-z0-9_' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 2 -f -o $OUT/full_$SAFE \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$SAFE.log 2>&1
  echo "ncu full $K exit $?"
done
ls -la $OUT
