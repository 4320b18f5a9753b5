timeout 300 python -m pytest tests -m gpu -x -q -k "tcgen05" 2>&1 | grep -a "passed\|failed"
for g in 2 1; do EQB_TC_EPI1_GROUPS=$g timeout 120 python tools/bench_stack.py; done
