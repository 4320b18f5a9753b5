"""Time eqb_vnsmall_forward (BASELINE configs[3] shape: 1024-point clouds, n_knn 20).  tools/bench_vnsmall.py [B] [N]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
from equiadapt_b200.pointcloud.canonicalization_networks.equivariant_networks import VNSmall

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.manual_seed(0)
net = VNSmall(SimpleNamespace(n_knn=20, pooling="mean")).cuda().eval()
x = torch.randn(B, 3, N, device="cuda")
with torch.no_grad():
    for _ in range(3): out = net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): out = net(x)
    e1.record(); torch.cuda.synchronize()
print(dict(os.environ.items() & {("EQB_VN_THREADS", os.environ.get("EQB_VN_THREADS", ""))}), f"vnsmall B={B} N={N}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per call, checksum {out.double().sum().item():.10f}")
