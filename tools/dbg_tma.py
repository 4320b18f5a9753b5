import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200 import ops
from oracle import reference_path as O
mode = sys.argv[1]; size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
x = torch.rand(2, 3, size, size, generator=torch.Generator().manual_seed(0))
g = {"exact": 2, "bilinear": 1, "ident": 0}[mode]
idx = torch.full((2,), g, dtype=torch.int32)
y = ops.warp_canonicalize(x.to(dev), idx.to(dev), 8, False)
torch.cuda.synchronize()
ang = torch.linspace(0.0, 360.0, 9)[:8][idx.long()]
ref = O.canonicalize_image(x, ang, None)
print(mode, size, "max err", float((y.cpu() - ref).abs().max()))
