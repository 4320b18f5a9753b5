"""Time the continuous (SO(2)) canonicalize warp at 512 x 3 x 224 x 224 (development aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200 import ops
B = 512
x = torch.rand(B, 3, 224, 224, device="cuda")
ang = torch.rand(B, device="cuda") * 6.2831853
m = torch.stack([torch.stack([torch.cos(ang), -torch.sin(ang)], 1), torch.stack([torch.sin(ang), torch.cos(ang)], 1)], 1).contiguous()
for _ in range(3): y = ops.warp_affine(x, m, None, True, 112, 112.0, 112.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): y = ops.warp_affine(x, m, None, True, 112, 112.0, 112.0)
e1.record(); torch.cuda.synchronize()
us = 1e3 * e0.elapsed_time(e1) / 20
print({k: v for k, v in os.environ.items() if k.startswith("EQB_")}, f"warp_affine 512x3x224x224: {us:.1f} us, {2 * x.numel() * 4 / us / 1e3:.0f} GB/s")
