for g in 1 2; do for l in 0 1; do EQB_TC_EPI1_GROUPS=$g EQB_TC_LIFT_EARLY=$l python tools/bench_stack.py; done; done
