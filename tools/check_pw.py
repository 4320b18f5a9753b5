"""eqb_conv2d_forward with k = 1, 256 -> 256 channels: tcgen05 kernel (tc::pw) against an fp64 torch reference and the SIMT kernel; timing at
the training shape (64 x 256 x 92 x 92).  Development aid.  tools/check_pw.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def ref(x, w, bias, relu, mask):
    y = torch.einsum("nc,bchw->bnhw", w[:, :, 0, 0].double(), x.double())
    if bias is not None: y = y + bias.double()[None, :, None, None]
    if relu: y = y.clamp_min(0)
    if mask is not None: y = torch.where(mask > 0, y, torch.zeros_like(y))
    return y


for (B, H, W) in ((2, 16, 16), (3, 20, 20), (2, 92, 92), (1, 4, 5 * 4)):
    for relu, use_mask, use_bias in ((True, False, True), (False, True, False)):
        # heterogeneous ranges per image: the operand scale is per image
        x = torch.randn(B, 256, H, W, device=dev) * torch.logspace(-2, 2, B, device=dev)[:, None, None, None]
        w = torch.randn(256, 256, 1, 1, device=dev) / 16
        bias = torch.randn(256, device=dev) if use_bias else None
        mask = torch.randn(B, 256, H, W, device=dev) if use_mask else None
        y = ops.conv2d_forward(x, w, bias, relu, mask)
        torch.cuda.synchronize()
        r = ref(x, w, bias, relu, mask)
        scale = r.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30)
        err = ((y.double() - r).abs() / scale).max().item()
        print(f"B={B} HxW={H}x{W} relu={relu} mask={use_mask}: max err / max|y| per image = {err:.3e}", "OK" if err < 2e-6 else "FAIL")

B, H, W = 64, 92, 92
x = torch.randn(B, 256, H, W, device=dev)
w = torch.randn(256, 256, 1, 1, device=dev) / 16
bias = torch.randn(256, device=dev)
mask = torch.randn(B, 256, H, W, device=dev)
for name, args in (("forward relu+bias", (x, w, bias, True, None)), ("dgrad with mask", (x, w, None, False, mask))):
    for _ in range(2): ops.conv2d_forward(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.conv2d_forward(*args)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 5 * 1e3:.0f} us per call (EQB_TRAIN_TC={os.environ.get('EQB_TRAIN_TC', '1')})")

# ---- k x k with cin * k * k <= 256 (the lift of the flagship network: 3 -> 256 channels, 5 x 5) ---------------------------
for (B, cin, H, W, k) in ((2, 3, 20, 24, 5), (3, 3, 96, 96, 5), (2, 7, 15, 13, 3), (1, 256, 9, 9, 1)):
    x = torch.randn(B, cin, H, W, device=dev) * torch.logspace(-1, 1, B, device=dev)[:, None, None, None]
    w = torch.randn(256, cin, k, k, device=dev) / 8
    bias = torch.randn(256, device=dev)
    y = ops.conv2d_forward(x, w, bias, True)
    torch.cuda.synchronize()
    r = torch.nn.functional.conv2d(x.double(), w.double(), bias.double()).clamp_min(0)
    scale = r.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30)
    err = ((y.double() - r).abs() / scale).max().item()
    print(f"conv B={B} cin={cin} {H}x{W} k={k}: max err / max|y| per image = {err:.3e}", "OK" if err < 4e-6 else "FAIL")
x = torch.randn(64, 3, 96, 96, device=dev); w = torch.randn(256, 3, 5, 5, device=dev) / 8; bias = torch.randn(256, device=dev)
for _ in range(2): ops.conv2d_forward(x, w, bias, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.conv2d_forward(x, w, bias, True)
e1.record(); torch.cuda.synchronize()
print(f"lift forward 64x3x96x96 -> 256 ch: {e0.elapsed_time(e1) / 5 * 1e3:.0f} us per call (EQB_TRAIN_TC={os.environ.get('EQB_TRAIN_TC', '1')})")

# ---- weight gradient ---------------------------------------------------------------------------------------------------
for (B, H, W) in ((2, 16, 16), (3, 20, 20), (2, 92, 92), (5, 4, 20)):
    x = torch.randn(B, 256, H, W, device=dev) * torch.logspace(-1, 1, B, device=dev)[:, None, None, None]
    dy = torch.randn(B, 256, H, W, device=dev) * torch.logspace(1, -1, B, device=dev)[:, None, None, None]
    dw = ops.conv2d_weight_grad(dy, x, 1)
    torch.cuda.synchronize()
    r = torch.einsum("bnhw,bchw->nc", dy.double(), x.double())
    err = ((dw[:, :, 0, 0].double() - r).abs().max() / r.abs().max()).item()
    print(f"wgrad B={B} HxW={H}x{W}: max err / max|dw| = {err:.3e}", "OK" if err < 4e-6 else "FAIL")
B, H, W = 64, 92, 92
x = torch.randn(B, 256, H, W, device=dev)
dy = torch.randn(B, 256, H, W, device=dev)
for _ in range(2): ops.conv2d_weight_grad(dy, x, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): dw = ops.conv2d_weight_grad(dy, x, 1)
e1.record(); torch.cuda.synchronize()
r = torch.einsum("bnp,bcp->nc", dy.flatten(2).double(), x.flatten(2).double())
print(f"wgrad 64x256x92x92: {e0.elapsed_time(e1) / 5 * 1e3:.0f} us per call, err {((dw[:, :, 0, 0].double() - r).abs().max() / r.abs().max()).item():.3e}")

# ---- weight gradient of k x k filters (patch operand gathered) ----------------------------------------------------------
for (B, cin, H, W, k) in ((2, 3, 20, 24, 5), (3, 3, 96, 96, 5), (2, 7, 14, 12, 3)):
    x = torch.randn(B, cin, H, W, device=dev) * torch.logspace(-1, 1, B, device=dev)[:, None, None, None]
    dy = torch.randn(B, 256, H - k + 1, W - k + 1, device=dev)
    dw = ops.conv2d_weight_grad(dy, x, k)
    torch.cuda.synchronize()
    wd = torch.zeros(256, cin, k, k, dtype=torch.float64, device=dev, requires_grad=True)
    (torch.nn.functional.conv2d(x.double(), wd) * dy.double()).sum().backward()
    err = ((dw.double() - wd.grad).abs().max() / wd.grad.abs().max()).item()
    print(f"wgrad k x k: B={B} cin={cin} {H}x{W} k={k}: max err / max|dw| = {err:.3e}", "OK" if err < 4e-6 else "FAIL")
x = torch.randn(64, 3, 96, 96, device=dev); dy = torch.randn(64, 256, 92, 92, device=dev)
for _ in range(2): ops.conv2d_weight_grad(dy, x, 5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.conv2d_weight_grad(dy, x, 5)
e1.record(); torch.cuda.synchronize()
print(f"lift wgrad 64x3x96x96, dy 64x256x92x92: {e0.elapsed_time(e1) / 5 * 1e3:.0f} us per call (EQB_TRAIN_TC={os.environ.get('EQB_TRAIN_TC', '1')})")
