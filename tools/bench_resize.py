"""Time the pre-network crop + antialiased resize (a3) at the benchmark's shape and check the quad variant against the
scalar one bit for bit (development aid).  EQB_RESIZE_NO_QUAD=1 selects the scalar kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200 import ops
torch.manual_seed(0)
for (B, H, crop, out) in ((512, 224, 180, 96), (512, 32, 29, 32), (64, 224, 224, 128)):
    x = torch.rand(B, 3, H, H, device="cuda")
    off = int(round((H - crop) / 2.0))
    f = lambda: ops.crop_resize_aa(x, off, off, crop, crop, out, out)
    for _ in range(3): y = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): y = f()
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 20
    byt = (B * 3 * crop * crop + y.numel()) * 4
    print({k: v for k, v in os.environ.items() if k.startswith("EQB_")}, f"crop_resize {B}x3x{H}^2 crop {crop} -> {out}: {us:.1f} us, {byt / us / 1e3:.0f} GB/s, checksum {float(y.double().sum()):.10f} {float(y.double().square().sum()):.10f}")
