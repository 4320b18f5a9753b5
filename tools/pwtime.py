import os, sys, torch
sys.path.insert(0, '/root/repo')
from equiadapt_b200 import ops
ops.event_log = {}
B, H, W = 64, 92, 92
dev='cuda'
x = torch.randn(B, 256, H, W, device=dev); w = torch.randn(256, 256, 1, 1, device=dev) / 16
bias = torch.randn(256, device=dev); mask = torch.randn(B, 256, H, W, device=dev)
for name, args in (("fwd", (x, w, bias, True, None)), ("dgrad", (x, w, None, False, mask))):
    for _ in range(2): ops.conv2d_forward(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.conv2d_forward(*args)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get('EQB_PW_DEBUG','0'), name, f"{e0.elapsed_time(e1)/5*1e3:.0f} us")
