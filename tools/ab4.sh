timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -a "passed\|failed"
for i in 1 2 3; do timeout 120 python tools/bench_stack.py; done
