"""CUDA-graph capture of a whole canonicalization step.

The public classes enqueue a step as 4-7 kernels through ctypes; on a busy host (eight ranks on one box, a data
loader, a logger) the gaps between those launches, not the kernels, decide the step time, and with a collective in
the step every rank advances at the pace of the slowest host thread of any rank.  `CapturedStep` records the step
ONCE -- kernels, the scratch they use, and (when torch.distributed is up) the 3-float NCCL all-reduce of the prior
statistic, forked onto NCCL's own stream inside the graph -- and replays it with ONE launch per step.

    step = canonicalizer.capture_step(x_example, induced_rep_type="scalar")     # or graphed.capture(fn, (x,))
    y, z, loss, metric = step(x)          # x is copied into the captured input buffer unless it IS that buffer

Semantics are the reference's (basecanonicalization.py:43-93): the captured callable is exactly
`canonicalizer(x)` -> `fn` -> `invert_canonicalization` -> `get_prior_regularization_loss` / `get_identity_metric`;
the returned tensors and `canonicalization_info_dict` are static buffers that every replay overwrites.
All ranks must capture and replay the same number of times when the prior statistic is synchronised.

A captured step is an INFERENCE artefact: it replays the kernels on the buffers that existed at capture time, among them the
packed filter orbits of the canonicalization network (networks_images.py).  After a parameter update (optimizer step,
load_state_dict) capture again -- the eager calls repack by themselves, a graph cannot.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch


def _flatten(out):
    if torch.is_tensor(out):
        return [out]
    return [t for t in out if torch.is_tensor(t)]


class CapturedStep:
    """`fn(*static_inputs)` recorded as one CUDA graph.

    fn            : any callable built from this package's ops (and torch.distributed collectives); must be
                    shape-static and must not synchronise the host
    static_inputs : example tensors; their storage becomes the graph's input buffers (copy-in on call unless the
                    caller passes the very same tensors)
    warmup        : eager runs on a side stream before capture (lazy one-time work: packed operands, function
                    attributes, NCCL channel set-up, is not capturable)
    """

    def __init__(self, fn: Callable, static_inputs: Sequence[torch.Tensor], warmup: int = 3):
        if not static_inputs or not all(torch.is_tensor(t) and t.is_cuda for t in static_inputs):
            raise RuntimeError("CapturedStep needs CUDA example inputs (there is no CPU fallback)")
        self.fn = fn
        self.inputs = tuple(static_inputs)
        dev = self.inputs[0].device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(int(warmup), 1)):
                fn(*self.inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (NCCL's watchdog, an NVML sampler) may legally touch the CUDA API
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"), torch.no_grad():
            self.outputs = fn(*self.inputs)
        self.replays = 0

    def __call__(self, *inputs: torch.Tensor):
        if inputs:
            if len(inputs) != len(self.inputs):
                raise ValueError(f"captured with {len(self.inputs)} inputs, called with {len(inputs)}")
            for dst, src in zip(self.inputs, inputs):
                if src is dst:
                    continue
                if src.shape != dst.shape or src.dtype != dst.dtype:
                    raise ValueError(f"captured for {tuple(dst.shape)} {dst.dtype}, got {tuple(src.shape)} {src.dtype}: "
                                     "a CUDA graph is shape-static, capture another step for this shape")
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.outputs


def capture(fn: Callable, static_inputs: Sequence[torch.Tensor], warmup: int = 3) -> CapturedStep:
    return CapturedStep(fn, static_inputs, warmup)


def capture_image_step(canonicalizer, x_example: torch.Tensor,
                       fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                       induced_rep_type: str = "scalar", with_prior: bool = True, warmup: int = 3) -> CapturedStep:
    """canonicalize -> fn -> invert_canonicalization [-> prior loss, identity metric] of an image canonicalizer."""

    def step(x):
        y = canonicalizer(x)
        f = y if fn is None else fn(y)
        z = canonicalizer.invert_canonicalization(f, induced_rep_type=induced_rep_type)
        if not with_prior:
            return y, z
        return y, z, canonicalizer.get_prior_regularization_loss(), canonicalizer.get_identity_metric()

    return CapturedStep(step, (x_example,), warmup)
