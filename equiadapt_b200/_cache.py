"""Parameter-derived device buffers (packed filter orbits, flat parameter blocks, folded batch norms) cached per
parameter VERSION.

The reference rebuilds its filter orbits on every forward (custom_group_equivariant_layers.py:103, :349-351), so it can
never go stale; this package builds them once per parameter version.  The cache key is (data_ptr, _version) of every
tensor that feeds the buffer, which torch bumps on every in-place op on the tensor itself (optimizer steps, `copy_`,
`load_state_dict`, initialisers under no_grad) -- but NOT on writes through `.data` (`p.data.normal_()`, some EMA /
weight-averaging code).  Hence:
  * `invalidate_packed()` drops the cache explicitly; it is called from `_apply` (.to / .cuda / .half ...),
    `load_state_dict` and `train` / `eval`;
  * `cache_packed = False` on a module restores the reference's behaviour (rebuild on every forward) for callers that
    write through `.data` and cannot call `invalidate_packed()`.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch


class PackedParameterCache:
    """Mixin for nn.Module subclasses; put it BEFORE nn.Module in the bases."""

    cache_packed: bool = True

    def invalidate_packed(self) -> None:
        self.__dict__["_packed_key"] = None

    def _packed_current(self, tensors: Iterable[Optional[torch.Tensor]]) -> bool:
        """True when the buffer built for `tensors` is still valid; records the new key otherwise."""
        key = tuple((t.data_ptr(), t._version) for t in tensors if t is not None)
        if self.cache_packed and self.__dict__.get("_packed_key") == key:
            return True
        self.__dict__["_packed_key"] = key
        return False

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_packed()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_packed()
        return out

    def train(self, mode: bool = True):
        self.invalidate_packed()
        return super().train(mode)
