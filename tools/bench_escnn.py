"""Time the a7 expanded-filter conv stack at the reference's default example configuration (development aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200.images.canonicalization_networks.escnn_networks import ESCNNEquivariantNetwork
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = ESCNNEquivariantNetwork((3, 96, 96), 32, 5, "rotation", 4, 3, device="cuda").eval()
x = torch.rand(B, 3, 96, 96, device="cuda")
flop = 2 * 128 * (92 * 92 * 75 + 88 * 88 * 3200 + 84 * 84 * 3200)
with torch.no_grad():
    for _ in range(2): act = net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): act = net(x)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"escnn C4 oc32 k5 L3 96x96 B={B}: {ms:.2f} ms/call, {B / ms * 1e3:.0f} img/s, {flop * B / ms / 1e9:.1f} TFLOP/s fp32")
