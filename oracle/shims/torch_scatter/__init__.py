"""Shim: torch_scatter.scatter(src, index, dim, reduce=) via index_add_ (segment sum / mean)."""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0, "shim covers the reference's call sites only (dim=0)"
    n = int(index.max().item()) + 1 if dim_size is None else dim_size
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    res.index_add_(0, index, src)
    if reduce == "mean":
        cnt = torch.zeros(n, dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp(min=1).view((n,) + (1,) * (src.dim() - 1))
        res = res / cnt
    elif reduce not in ("sum", "add"):
        raise NotImplementedError(reduce)
    return res
