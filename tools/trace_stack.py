"""Pipeline timeline of the CTA-pair stack kernel (development aid): eqb_debug_stack_trace -> per-tile event times in SM cycles.
   python tools/trace_stack.py [first_tile] [n_tiles]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from equiadapt_b200 import native
T0 = int(sys.argv[1]) if len(sys.argv) > 1 else 10
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 6
TILES = 64
net = bench.make_layers().cuda()
x = torch.rand(512, 3, 96, 96, device="cuda")
buf = torch.zeros(TILES * 64, dtype=torch.int64, device="cuda")
with torch.no_grad():
    for _ in range(2): net(x)
    torch.cuda.synchronize()
    native.lib().eqb_debug_stack_trace(buf.data_ptr(), TILES)
    net(x)
    torch.cuda.synchronize()
    native.lib().eqb_debug_stack_trace(None, 0)
t = buf.cpu().view(TILES, 64)
names = {0: "I1 wait D2EMPTY", 1: "I1 got D2EMPTY", 10: "L wait", 11: "L got D1EMPTY(+G2)", 17: "E1g0 D1FULL", 26: "E1g1 D1FULL",
         35: "E2g0 D2FULL", 36: "E2g0 release", 37: "E2g0 done", 38: "E2g1 D2FULL", 39: "E2g1 release", 40: "E2g1 done"}
for k in range(8): names[2 + k] = f"I1 A1FULL atom{k}"
for k in range(5): names[12 + k] = f"L A0FULL slab{k}"; names[41 + k] = f"IM A0EMPTY slab{k}"; names[46 + k] = f"L pass1 slab{k}"
for k in range(4):
    names[18 + k] = f"E1g0 ld atom{2*k}"; names[22 + k] = f"E1g0 A1EMPTY atom{2*k}"
    names[27 + k] = f"E1g1 ld atom{2*k+1}"; names[31 + k] = f"E1g1 A1EMPTY atom{2*k+1}"
print({k: v for k, v in os.environ.items() if k.startswith("EQB_")})
per = (t[T0 + NT, 1] - t[T0, 1]).item() / NT
print(f"tile period over tiles {T0}..{T0+NT}: {per:.0f} cycles (MMA floor 8064)")
d = [(t[i + 1, 1] - t[i, 1]).item() for i in range(2, TILES - 2)]
print("per-tile periods (I1 got D2EMPTY -> next), tiles 2..:", d)
print(f"mean over tiles 4..{TILES-4}: {(t[TILES-4, 1] - t[4, 1]).item() / (TILES - 8):.0f} cycles")
origin = t[T0, 1].item()
ev = []
for tile in range(T0, T0 + NT):
    for s, nm in names.items():
        v = t[tile, s].item()
        if v: ev.append((v - origin, tile, nm))
ev.sort()
for dt, tile, nm in ev:
    print(f"{dt:8d}  t{tile:<3d} {nm}")
