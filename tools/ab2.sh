timeout 300 python -m pytest tests -m gpu -x -q -k "tcgen05 or full_pipeline or golden" 2>&1 | tail -4
python -c "
import ctypes, sys
sys.path.insert(0,'.')
from equiadapt_b200 import native
o=(ctypes.c_int*5)(); print('stall', native.lib().eqb_debug_last_stall(o), list(o))"
for g in 2 1; do for l in 1 0; do EQB_TC_EPI1_GROUPS=$g EQB_TC_LIFT_EARLY=$l timeout 120 python tools/bench_stack.py; done; done
EQB_TC_PAIR=0 timeout 120 python tools/bench_stack.py
