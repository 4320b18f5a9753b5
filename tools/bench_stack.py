"""Time eqb_gconv_stack_run alone at the bench configuration (development aid; CUDA events, 20 calls)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from equiadapt_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
net = bench.make_layers().cuda()
x = torch.rand(B, 3, 96, 96, device="cuda")
with torch.no_grad():
    for _ in range(3): act = net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): act = net(x)
    e1.record(); torch.cuda.synchronize()
print({k: v for k, v in os.environ.items() if k.startswith("EQB_")}, "stack call us:", round(1e3 * e0.elapsed_time(e1) / 20, 1), "checksum", float(act.double().sum()))
