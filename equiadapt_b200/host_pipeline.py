"""Host-buffer front end of the image hot path: pinned host batch in, pinned host batch out.

The reference is called with whatever tensor the data loader produced; on a GPU box that is a pinned
host batch that has to cross PCIe before `canonicalize()` and cross it again after
`invert_canonicalization()`.  Done naively (copy the whole batch, run, copy back) the three phases
serialise and the step is bounded by H2D + compute + D2H.  `HostStreamedCanonicalizer` splits the batch
into shards of `shard` images and runs a three-stream pipeline

    h2d stream     : x_host[shard i+1] -> device slot            (copy engine 0)
    compute stream : canonicalize(x_i) -> fn -> invert(...)      (the sm_100a kernels, current-stream API)
    d2h stream     : result_i -> out_host[shard i]               (copy engine 1)

with event hand-offs per slot, so the step is bounded by max(H2D, compute, D2H) plus one shard of
fill/drain.  Every op of the path is per-sample independent, so sharding the batch
changes no result; the prior statistic [sum CE, sum identity, B] is summed over the shards on the
device and all-reduced once (distributed.allreduce_stats) exactly like the un-sharded call.

Calls made per shard are the public ones of the reference surface (basecanonicalization.py:43-93):
`canonicalizer(x)`, `canonicalizer.invert_canonicalization(out, induced_rep_type=...)`.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch

from . import distributed as D


class HostStreamedCanonicalizer:
    """Pipelined `canonicalize -> fn -> invert_canonicalization` over a pinned host batch.

    canonicalizer : a discrete-group image canonicalizer of this package
    fn            : the caller's prediction network on the canonicalized shard (None = identity);
                    runs on the compute stream, must return a (b, C', H, W) feature map
    shard         : images per pipeline stage
    slots         : device input buffers in flight (>= 2)
    ramp          : shorten the first and last shards (shard/4, shard/2, shard, ..., shard/2, shard/4): the pipeline's
                    fill (first H2D before any kernel can run) and drain (last D2H after the last kernel) shrink from one
                    full shard each to a quarter shard
    """

    def __init__(self, canonicalizer, fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 induced_rep_type: str = "scalar", shard: int = 64, slots: int = 3,
                 device: Optional[torch.device] = None, ramp: bool = False):
        if slots < 2:
            raise ValueError("need at least two slots to overlap copies with compute")
        self.can = canonicalizer
        self.fn = fn
        self.rep = induced_rep_type
        self.shard = int(shard)
        self.slots = int(slots)
        self.ramp = bool(ramp)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.s_h2d = torch.cuda.Stream(self.device)
        self.s_cmp = torch.cuda.Stream(self.device)
        self.s_d2h = torch.cuda.Stream(self.device)
        self._x_dev = None
        self.last_stats: Optional[torch.Tensor] = None

    def shard_bounds(self, B: int):
        """[(lo, hi)] of the pipeline stages over a batch of B samples."""
        sizes = []
        if self.ramp and self.shard >= 4 and B >= 4 * self.shard:
            head = [self.shard // 4, self.shard // 2]
            body = B - 2 * sum(head)
            sizes = head + [self.shard] * (body // self.shard) + ([body % self.shard] if body % self.shard else []) + head[::-1]
        else:
            sizes = [self.shard] * (B // self.shard) + ([B % self.shard] if B % self.shard else [])
        out, lo = [], 0
        for n in sizes:
            out.append((lo, lo + n))
            lo += n
        return out

    def _ensure_slots(self, shape, dtype):
        want = (self.slots, self.shard) + tuple(shape)
        if self._x_dev is None or tuple(self._x_dev.shape) != want or self._x_dev.dtype != dtype:
            self._x_dev = torch.empty(want, dtype=dtype, device=self.device)

    def __call__(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        """x_host (B,C,H,W) pinned -> out_host (B,C',H,W) pinned; returns (prior loss, identity metric) as
        0-d device tensors.  The call returns with all work enqueued; the caller's current stream waits on
        the last D2H copy, so reading `out_host` after synchronising that stream (or after `.item()` on
        either returned tensor followed by a stream sync) is safe.  out_host = None: nothing but the two scalars
        leaves the device (the training-loop shape: the canonicalized batch feeds a device-side consumer)."""
        if x_host.is_cuda or (out_host is not None and out_host.is_cuda):
            raise ValueError("HostStreamedCanonicalizer takes HOST tensors; call the canonicalizer directly on device tensors")
        if not (x_host.is_pinned() and (out_host is None or out_host.is_pinned())):
            raise ValueError("host buffers must be pinned (torch.Tensor.pin_memory) for asynchronous copies")
        B = x_host.shape[0]
        if out_host is not None and out_host.shape[0] != B:
            raise ValueError("output buffer must hold one result per input sample")
        self._ensure_slots(x_host.shape[1:], x_host.dtype)
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_h2d, self.s_cmp, self.s_d2h):
            s.wait_stream(cur)                      # buffers prepared on the caller's stream are visible
        bounds = self.shard_bounds(B)
        h2d_done = [None] * self.slots              # input slot filled
        cmp_done = [None] * self.slots              # input slot consumed
        d2h_done = [None] * self.slots              # result of the shard that used this slot copied out
        keep = [None] * self.slots                  # results stay referenced until their D2H finished
        stats_sum = None
        prefetch, self.can.prefetch_prior_allreduce = self.can.prefetch_prior_allreduce, False   # one collective at the end
        for i, (lo, hi) in enumerate(bounds):
            s = i % self.slots
            xd = self._x_dev[s, : hi - lo]
            with torch.cuda.stream(self.s_h2d):
                if cmp_done[s] is not None:
                    self.s_h2d.wait_event(cmp_done[s])
                xd.copy_(x_host[lo:hi], non_blocking=True)
                h2d_done[s] = self.s_h2d.record_event()
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(h2d_done[s])
                if d2h_done[s] is not None:
                    self.s_cmp.wait_event(d2h_done[s])   # the result that lived in this slot has left the device
                keep[s] = None
                y = self.can(xd)
                f = y if self.fn is None else self.fn(y)
                z = self.can.invert_canonicalization(f, induced_rep_type=self.rep)
                st = self.can._discrete_stats()
                stats_sum = st[:3].clone() if stats_sum is None else stats_sum.add_(st[:3])
                cmp_done[s] = self.s_cmp.record_event()
                keep[s] = (y, f, z)
            if out_host is not None:
                with torch.cuda.stream(self.s_d2h):
                    self.s_d2h.wait_event(cmp_done[s])
                    out_host[lo:hi].copy_(z, non_blocking=True)
                    d2h_done[s] = self.s_d2h.record_event()
        self.can.prefetch_prior_allreduce = prefetch
        with torch.cuda.stream(self.s_cmp):
            for ev in d2h_done:
                if ev is not None:
                    self.s_cmp.wait_event(ev)
            keep = None
            total = D.allreduce_stats(stats_sum) if self.can.sync_prior_across_ranks else stats_sum
            loss, ident = total[0] / total[2], total[1] / total[2]
            for t in (total, loss, ident):
                t.record_stream(cur)
        cur.wait_stream(self.s_cmp)
        cur.wait_stream(self.s_d2h)
        self.last_stats = total
        return loss, ident
