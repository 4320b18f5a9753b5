import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200.images.canonicalization_networks.custom_equivariant_networks import CustomEquivariantNetwork
from oracle import reference_path as O
dev = torch.device("cuda:0")
cout, k, n, B, R = [int(v) for v in (sys.argv[1:6] + [None] * 5)[:5]] if len(sys.argv) > 5 else (32, 5, 8, 4, 96)
torch.manual_seed(0)
net = CustomEquivariantNetwork((3, R, R), cout, k, "rotation", n, 3, device="cpu")
with torch.no_grad():
    for m in net.eqv_network:
        if hasattr(m, "bias"):
            m.bias.uniform_(-0.05, 0.05)
lay = [(m.weights.detach().clone(), m.bias.detach().clone()) for m in net.eqv_network if hasattr(m, "weights")]
x = torch.rand(B, 3, R, R, generator=torch.Generator().manual_seed(1))
def run(no_tc):
    if no_tc: os.environ["EQB_NO_TC"] = "1"
    else: os.environ.pop("EQB_NO_TC", None)
    torch.manual_seed(0)
    net2 = CustomEquivariantNetwork((3, R, R), cout, k, "rotation", n, 3, device="cpu")
    net2.load_state_dict(net.state_dict())
    net2 = net2.to(dev)
    with torch.no_grad():
        a = net2(x.to(dev)); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): a = net2(x.to(dev))
        torch.cuda.synchronize()
    return a.cpu(), (time.perf_counter() - t0) / 3
a_simt, t_simt = run(True)
print("simt ok", t_simt * 1e3, "ms", flush=True)
try:
    a_tc, t_tc = run(False)
except Exception as ex:
    import ctypes
    from equiadapt_b200 import native
    out = (ctypes.c_int * 5)()
    flag = native.lib().eqb_debug_last_stall(out)
    print("TC FAILED:", str(ex).splitlines()[0], "| stall report flag,block,warp,barrier,parity =", list(out), flush=True)
    sys.exit(1)
print("tc ok", t_tc * 1e3, "ms", flush=True)
a64 = O.custom_equivariant_network(x.double(), [(w.double(), b.double()) for w, b in lay], n, False)
e = lambda a: float((a.double() - a64).abs().max() / a64.abs().max())
print("rel err vs fp64: simt %.3e  tc %.3e   tc vs simt %.3e" % (e(a_simt), e(a_tc), float((a_tc - a_simt).abs().max() / a_simt.abs().max())))
print(a_tc[0], a_simt[0], a64[0].float())
