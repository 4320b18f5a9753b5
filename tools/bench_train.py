"""Time one training step of the flagship canonicalizer (C8, 3x224x224, CustomEquivariantNetwork 32 ch / k5 / 3 layers,
crop 0.8, resize 96) on this package's training path: forward (layer-wise, feature maps kept) + task loss through the
canonicalize warp + prior loss + backward.  Development aid; the benchmarked path is inference (bench.py)."""
import os, sys, time, torch
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200 import ops
from equiadapt_b200.images.canonicalization.discrete_group import GroupEquivariantImageCanonicalization
from equiadapt_b200.images.canonicalization_networks.custom_equivariant_networks import CustomEquivariantNetwork

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
net = CustomEquivariantNetwork((3, 96, 96), 32, 5, "rotation", 8, 3, device="cuda")
can = GroupEquivariantImageCanonicalization(net, SimpleNamespace(beta=1.0, input_crop_ratio=0.8, resize_shape=96), (3, 224, 224)).train()
x = torch.rand(B, 3, 224, 224, device="cuda")
w = torch.randn(B, 3, 224, 224, device="cuda")
ops.event_log = {}


def step():
    for p in can.parameters():
        p.grad = None
    y = can(x)
    loss = (y * w).mean() + 100.0 * can.get_prior_regularization_loss()
    loss.backward()
    return loss


for _ in range(2): step()
torch.cuda.synchronize()
ops.event_log = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for _ in range(n): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"training step, batch {B}: {ms:.2f} ms = {B / ms * 1e3:.0f} img/s")
for name, evs in sorted(ops.event_log.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1])):
    tot = sum(a.elapsed_time(b) for a, b in evs) / n
    print(f"  {name:40s} {len(evs) // n:3d} calls/step {tot:8.3f} ms/step")
with torch.no_grad():
    can.eval()
    for _ in range(2): can(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n): can(x)
    e1.record(); torch.cuda.synchronize()
print(f"inference (fused stack), same batch: {e0.elapsed_time(e1) / n:.2f} ms")
