"""CPU study for the next step of the training path: would the fp16 hi/lo operand split of the
inference stack carry the weight-gradient GEMM  dW[n,c] = sum_k dY[n,k] X[c,k]  (K = B*H*W up to 541 696 at the benchmark
shape) within the 2e-5 tolerance the training tests hold the fp32 SIMT kernels to?  Compares, against fp64:
  fp32      : float32 products and float32 accumulation in 128-wide K chunks (what an fp32 kernel does, roughly)
  split3    : a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with exact products, float32 accumulation per 128-wide chunk (TMEM), chunks
              summed in float32
  split3+64 : the same, chunk results summed in float64 (a split-K epilogue that keeps partials in fp64)
Operands are scaled by exact powers of two into the fp16 range first, as tc_header_kernel does.  (development aid)"""
import math, sys, torch

def pow2_scale(m):
    f, e = math.frexp(m)
    return math.ldexp(1.0, 14 - e)

def split(x):
    hi = x.half(); lo = (x - hi.float()).half()
    return hi.float(), lo.float()

def chunked(a, b, chunk, acc_dtype):
    out = torch.zeros(a.shape[0], b.shape[0], dtype=acc_dtype)
    for k0 in range(0, a.shape[1], chunk):
        out += (a[:, k0:k0 + chunk].double() @ b[:, k0:k0 + chunk].double().t()).float().to(acc_dtype)   # exact products, fp32-rounded chunk
    return out

torch.manual_seed(0)
N = C = 64
for K in (8464, 8464 * 8, 8464 * 64):
    X = torch.relu(torch.randn(C, K))                        # post-ReLU feature map
    dY = torch.randn(N, K) * (torch.rand(N, K) > 0.5) / K    # masked upstream gradient of a mean
    ref = dY.double() @ X.double().t()
    sa, sb = pow2_scale(float(dY.abs().max())), pow2_scale(float(X.abs().max()))
    ah, al = split(dY * sa); bh, bl = split(X * sb)
    rel = lambda v: float((v.double() - ref).abs().max() / ref.abs().max())
    fp32 = chunked(dY, X, 128, torch.float32)
    s32 = (chunked(ah, bh, 128, torch.float32) + chunked(al, bh, 128, torch.float32) + chunked(ah, bl, 128, torch.float32)) / (sa * sb)
    s64 = (chunked(ah, bh, 128, torch.float64) + chunked(al, bh, 128, torch.float64) + chunked(ah, bl, 128, torch.float64)) / (sa * sb)
    print(f"K = {K:7d}: fp32 {rel(fp32):.2e}   split3 {rel(s32):.2e}   split3 + fp64 partial sums {rel(s64):.2e}")
