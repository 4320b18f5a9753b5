"""cfg5 step (canonicalize + invert + prior) for several system counts on one GPU (development aid)."""
import os, sys, torch
from types import SimpleNamespace as NS
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200 import ops
from equiadapt_b200.nbody.canonicalization.euclidean_group import EuclideanGroupNBody
from equiadapt_b200.nbody.canonicalization_networks.custom_equivariant_networks import VNDeepSets
dev = "cuda"
for S in (1250, 2500, 4000, 5000, 6000, 7500, 10000, 20000):
    torch.manual_seed(0)
    hp = NS(out_dim=4, hidden_dim=16, layer_pooling="mean", final_pooling="mean", num_layers=4, nonlinearity="relu",
            canon_feature="p", canon_translation=False, angular_feature=0, dropout=0.5, batch_size=S)
    nb = EuclideanGroupNBody(VNDeepSets(hp, device=dev).eval()).eval()
    loc, vel = torch.randn(5 * S, 3, device=dev), torch.randn(5 * S, 3, device=dev)
    ch = (torch.randint(0, 2, (5 * S, 1), device=dev) * 2 - 1).float()
    base = torch.tensor([[i, j] for i in range(5) for j in range(5) if i != j], device=dev).t()
    edges = (base[:, None, :] + 5 * torch.arange(S, device=dev)[None, :, None]).reshape(2, -1)
    nodes = torch.sqrt(torch.sum(vel ** 2, dim=1)).unsqueeze(1)
    def step():
        cl, cv = nb(nodes, None, loc=loc, edges=edges, vel=vel, edge_attr=None, charges=ch)
        back = nb.invert_canonicalization(cl)
        return nb.get_prior_regularization_loss()
    with torch.no_grad():
        for _ in range(3): step()
        torch.cuda.synchronize()
        ops.event_log = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): step()
        e1.record(); torch.cuda.synchronize()
        log, ops.event_log = ops.event_log, None
    per = {k: round(1e3 * sum(a.elapsed_time(b) for a, b in v) / len(v), 1) for k, v in log.items()}
    print(f"S={S}: {1e3 * e0.elapsed_time(e1) / 20:.1f} us/step", per)
