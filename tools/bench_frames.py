"""Time BASELINE configs[3] / [4] end to end on one GPU (development aid): clouds 128 x 3 x 1024, 10 000 systems."""
import os, sys, torch
from types import SimpleNamespace as NS
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200.pointcloud.canonicalization.continuous_group import EquivariantPointcloudCanonicalization
from equiadapt_b200.pointcloud.canonicalization_networks.equivariant_networks import VNSmall
from equiadapt_b200.nbody.canonicalization.euclidean_group import EuclideanGroupNBody
from equiadapt_b200.nbody.canonicalization_networks.custom_equivariant_networks import VNDeepSets

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

torch.manual_seed(0)
dev = "cuda"
can = EquivariantPointcloudCanonicalization(VNSmall(NS(n_knn=20, pooling="mean")).to(dev).eval(), NS()).eval()
x = torch.randn(128, 3, 1024, device=dev)
with torch.no_grad():
    ms = timeit(lambda: (can(x), can.get_prior_regularization_loss()))
print(f"cfg4 pointcloud SO(3) B=128 N=1024: {ms:.3f} ms/batch, {128 / ms * 1e3:.0f} clouds/s")
x16 = x[:16].contiguous()
with torch.no_grad():
    ms = timeit(lambda: (can(x16), can.get_prior_regularization_loss()))
print(f"cfg4 per-GPU shard of the 8-GPU configuration, B=16 N=1024: {ms:.3f} ms/batch, {16 / ms * 1e3:.0f} clouds/s")
S = 10000
hp = NS(out_dim=4, hidden_dim=16, layer_pooling="mean", final_pooling="mean", num_layers=4, nonlinearity="relu",
        canon_feature="p", canon_translation=False, angular_feature=0, dropout=0.5, batch_size=S)
nb = EuclideanGroupNBody(VNDeepSets(hp, device=dev).eval()).eval()
loc, vel = torch.randn(5 * S, 3, device=dev), torch.randn(5 * S, 3, device=dev)
ch = (torch.randint(0, 2, (5 * S, 1), device=dev) * 2 - 1).float()
base = torch.tensor([[i, j] for i in range(5) for j in range(5) if i != j], device=dev).t()
edges = (base[:, None, :] + 5 * torch.arange(S, device=dev)[None, :, None]).reshape(2, -1)
nodes = torch.sqrt(torch.sum(vel ** 2, dim=1)).unsqueeze(1)
with torch.no_grad():
    def step():
        cl, cv = nb(nodes, None, loc=loc, edges=edges, vel=vel, edge_attr=None, charges=ch)
        back = nb.invert_canonicalization(cl)
        return nb.get_prior_regularization_loss()
    ms = timeit(step)
print(f"cfg5 n-body E(3) 10k systems (canonicalize + invert + prior): {ms:.3f} ms/batch, {S / ms * 1e3:.0f} systems/s")
