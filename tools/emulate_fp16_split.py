"""CPU emulation of the fp16 hi/lo split used by the tcgen05 stack: error of the group activations vs fp64.
(development aid; imports the oracle, never used by the product path)"""
import math, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_path as O
import bench

def pow2_scale(m):
    if not (m > 0): return 1.0
    f, e = math.frexp(m)
    return math.ldexp(1.0, 14 - e)

def split(x):
    hi = x.half()
    lo = (x - hi.float()).half()
    return hi.double(), lo.double()

def mm3(a, w):
    ah, al = split(a); wh, wl = split(w)
    return (ah @ wh + al @ wh + ah @ wl)  # fp64 accumulate: isolates the operand-split error

torch.manual_seed(0)
net = bench.make_layers()
layers = [(m.weights.detach(), m.bias.detach()) for m in net.eqv_network if hasattr(m, "weights")]
B = 6
x = bench.host_batch(B, 1)
xp = O.pre_network_transform(x, bench.IN_SHAPE, bench.CROP, bench.RESIZE)
act32 = O.custom_equivariant_network(xp, layers, 8, False)
act64 = O.custom_equivariant_network(xp.double(), [(w.double(), b.double()) for w, b in layers], 8, False)
W0 = O.lift_filter_orbit(layers[0][0], 8, False)          # (256,3,5,5)
W1 = O.regular_filter_orbit(layers[1][0], 8, False)[:, :, 0, 0]   # (256,256)
W2 = O.regular_filter_orbit(layers[2][0], 8, False)[:, :, 0, 0]
b0 = layers[0][1].repeat_interleave(8); b1 = layers[1][1].repeat_interleave(8); b2 = layers[2][1].repeat_interleave(8)
A0 = torch.nn.functional.unfold(xp, 5).transpose(1, 2)     # (B, P, 75)
W0m = W0.reshape(256, 75)
amax = float(xp.abs().max()); sx = pow2_scale(amax); sw0 = pow2_scale(float(W0m.abs().max())); sw1 = pow2_scale(float(W1.abs().max()))
R0 = float(W0m.abs().sum(1).max()); s1 = pow2_scale(amax * R0 + float(b0.abs().max()))
print("scales", sx, sw0, sw1, s1)
d1 = mm3((A0 * sx).reshape(-1, 75), (W0m * sw0).t()) / (sx * sw0)
a1 = torch.clamp(d1.float() + b0, min=0)                  # fp32 epilogue
d2 = mm3(a1 * s1, (W1 * sw1).t()) / (s1 * sw1)
a2 = torch.clamp(d2.float() + b1, min=0).double().reshape(B, -1, 256)
S = a2.sum(1)                                              # (B,256)
fold = W2.double().reshape(32, 8, 256).sum(0)              # (8,256)
act = S @ fold.t() / (32 * a2.shape[1]) + b2.double().mean()
print("|fp32 ref - fp64|max", float((act32.double() - act64).abs().max()))
print("|fp16split - fp64|max", float((act - act64).abs().max()))
print("act scale", float(act64.abs().mean()), "top2 gaps", (act64.sort(1).values[:, -1] - act64.sort(1).values[:, -2]).tolist())
