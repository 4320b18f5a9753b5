"""HBM-bound warp kernels at the SURVEY.md 8d sizes (development aid): invert regular (C = 8), orbit expand (cfg3)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200 import ops
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6536.4

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n

dev = "cuda"
idx = torch.randint(0, 8, (512,), device=dev, dtype=torch.int32)
f = torch.randn(512, 8, 224, 224, device=dev)
us = timeit(lambda: ops.warp_invert(f, idx, 8, False, True))
print(f"invert regular C8, 512x8x224x224: {us:.1f} us, {2 * f.numel() * 4 / us / 1e3:.0f} GB/s = {2 * f.numel() * 4 / us / 1e3 / peak:.3f} of HBM peak")
x3 = torch.rand(512, 3, 224, 224, device=dev)
us = timeit(lambda: ops.warp_invert(x3, idx, 8, False, False))
print(f"invert scalar, 512x3x224x224: {us:.1f} us, {2 * x3.numel() * 4 / us / 1e3 / peak:.3f} of HBM peak")
us = timeit(lambda: ops.warp_canonicalize(x3, idx, 8, False))
print(f"canonicalize, 512x3x224x224: {us:.1f} us, {2 * x3.numel() * 4 / us / 1e3 / peak:.3f} of HBM peak")
for b in (32, 256):
    xr = torch.rand(b, 3, 96, 96, device=dev)
    us = timeit(lambda: ops.orbit_expand(xr, 48, 96, 4, True))
    byt = xr.numel() * 4 * (1 + 8)
    print(f"orbit expand D4, {b}x3x96x96 -> {8 * b}x3x96x96: {us:.1f} us, {byt / us / 1e3:.0f} GB/s = {byt / us / 1e3 / peak:.3f} of HBM peak")
# the bench step's case: every sample on an odd C8 element (all-bilinear tiles)
idx_odd = (torch.randint(0, 4, (512,), device=dev, dtype=torch.int32) * 2 + 1)
us = timeit(lambda: ops.warp_canonicalize(x3, idx_odd, 8, False))
print(f"canonicalize (all odd C8 elements), 512x3x224x224: {us:.1f} us, {2 * x3.numel() * 4 / us / 1e3 / peak:.3f} of HBM peak")
us = timeit(lambda: ops.warp_invert(x3, idx_odd, 8, False, False))
print(f"invert scalar (all odd C8 elements), 512x3x224x224: {us:.1f} us, {2 * x3.numel() * 4 / us / 1e3 / peak:.3f} of HBM peak")
