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

# ---- the same step captured as ONE CUDA graph (forward + loss + backward; gradients land in static .grad buffers) -----------
if "--graph" in sys.argv or os.environ.get("EQB_TRAIN_GRAPH", "1") == "1":
    try:
        can.train()
        ops.event_log = None
        params = [p for p in can.parameters() if p.requires_grad]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                for p in params:
                    p.grad = None
                y = can(x)
                loss = (y * w).mean() + 100.0 * can.get_prior_regularization_loss()
                loss.backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        for p in params:
            p.grad = None
        # (captured on the stream the warm-up ran on: the training kernels keep their scratch per (device, stream))
        with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
            y = can(x)
            loss_g = (y * w).mean() + 100.0 * can.get_prior_regularization_loss()
            loss_g.backward()
        grads_ref = None
        for _ in range(3): graph.replay()
        torch.cuda.synchronize()
        e0.record()
        n = 20
        for _ in range(n): graph.replay()
        e1.record(); torch.cuda.synchronize()
        ms_g = e0.elapsed_time(e1) / n
        # same gradients as the eager step
        g_graph = [p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        step()
        torch.cuda.synchronize()
        g_eager = [p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        loss_e = step()
        torch.cuda.synchronize()
        names = [n for n, p in can.named_parameters() if p.requires_grad]
        print(f"training step as one CUDA graph replay, batch {B}: {ms_g:.2f} ms = {B / ms_g * 1e3:.0f} img/s; "
              f"loss {float(loss_g.detach()):.6f} (eager {float(loss_e.detach()):.6f})")
        for nme, a, b2, p in zip(names, g_graph, g_eager, params):
            ref = float(b2.abs().max())
            print(f"  {nme:45s} |grad|max {ref:.3e}  graph vs eager {float((a - b2).abs().max()) / max(ref, 1e-30):.1e}  "
                  f"eager vs eager {float((p.grad - b2).abs().max()) / max(ref, 1e-30):.1e}")
    except Exception as exc:  # development aid: report, do not hide
        print("graph capture of the training step failed:", repr(exc)[:300])
