"""Tensor-level calls into the C ABI (include/equiadapt_b200.h).

PyTorch is only the allocator and the stream here: every function checks its tensors, takes
`data_ptr()`s and the CURRENT torch stream, and enqueues the hand-written sm_100a kernels.
Nothing synchronises the host and nothing falls back to torch arithmetic: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import native

# kernels launched through this module since import (bench.py reports the count per step)
launch_count = 0


def _need_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "equiadapt_b200 runs on CUDA tensors only (there is no CPU fallback); got a tensor on " + str(t.device))
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} and {t.device}")
    return dev


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"equiadapt_b200 computes in float32; got {t.dtype}")
    return t.detach().contiguous()


def _no_backward(what: str, *tensors: Optional[torch.Tensor]) -> None:
    """Fail loudly where the reference would differentiate and this path has no backward kernel (instead of silently
    returning a graph-less result)."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(f"{what} has no backward on the B200 path: detach() the argument or run under torch.no_grad()")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


# optional per-call device timing: set to a dict {abi name: [(start_event, end_event), ...]}; events are
# recorded on the launching stream (used by bench.py for the live roofline figure, never on by default)
event_log = None


def _call(name: str, n_launches: int, dev: torch.device, *args):
    global launch_count
    with torch.cuda.device(dev):
        if event_log is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(dev))
            native.check(getattr(native.lib(), name)(*args), name)
            e1.record(torch.cuda.current_stream(dev))
            event_log.setdefault(name, []).append((e0, e1))
        else:
            native.check(getattr(native.lib(), name)(*args), name)
    launch_count += n_launches


# ---- a3 -----------------------------------------------------------------------------------------
def crop_resize_aa(x: torch.Tensor, top: int, left: int, crop_h: int, crop_w: int, out_h: int, out_w: int,
                   with_absmax: bool = False) -> torch.Tensor:
    """CenterCrop + antialiased resize.  with_absmax: the kernel also reduces max |y[b]| per image and the result carries
    it as `y._eqb_absmax` (B floats) -- the per-image operand scale CustomEquivariantNetwork hands to the conv stack, so no
    separate pass over y is launched."""
    dev = _need_cuda(x)
    _no_backward("crop_resize_aa (gradient with respect to the image)", x)
    x = _f32(x)
    b, c, h, w = x.shape
    y = torch.empty((b, c, out_h, out_w), dtype=torch.float32, device=dev)
    if with_absmax:
        amax = torch.empty((max(b, 1),), dtype=torch.float32, device=dev)
        _call("eqb_crop_resize_aa_absmax", 1, dev, _ptr(x), _ptr(y), _ptr(amax), b, c, h, w, top, left, crop_h, crop_w, out_h,
              out_w, _stream(dev))
        y._eqb_absmax = amax
        return y
    _call("eqb_crop_resize_aa", 1, dev, _ptr(x), _ptr(y), b, c, h, w, top, left, crop_h, crop_w, out_h, out_w, _stream(dev))
    return y


# ---- a4 / a5 ------------------------------------------------------------------------------------
def lift_filter_orbit(w: torch.Tensor, num_rotations: int, reflect: bool) -> torch.Tensor:
    dev = _need_cuda(w)
    w = _f32(w)
    cout, cin, k, k2 = w.shape
    assert k == k2
    g = num_rotations * (2 if reflect else 1)
    out = torch.empty((cout * g, cin, k, k), dtype=torch.float32, device=dev)
    _call("eqb_lift_filter_orbit", 1, dev, _ptr(w), _ptr(out), cout, cin, k, num_rotations, int(reflect), _stream(dev))
    return out


def regular_filter_orbit(w: torch.Tensor, num_rotations: int, reflect: bool) -> torch.Tensor:
    dev = _need_cuda(w)
    w = _f32(w)
    cout, cin, g, k, k2 = w.shape
    assert k == k2 and g == num_rotations * (2 if reflect else 1)
    out = torch.empty((cout * g, cin * g, k, k), dtype=torch.float32, device=dev)
    _call("eqb_regular_filter_orbit", 1, dev, _ptr(w), _ptr(out), cout, cin, k, num_rotations, int(reflect), _stream(dev))
    return out


# ---- N3: training kernels of the group-conv network -------------------------------------------------
def _known_amax(t: torch.Tensor) -> Optional[torch.Tensor]:
    """Per-sample max |t| recorded by the op that produced `t` (valid only while `t` has not been written in place since)."""
    rec = getattr(t, "_eqb_amax", None)
    if rec is not None and rec[1] == t._version and rec[0].device == t.device:
        return rec[0]
    return None


def _record_amax(t: torch.Tensor, amax: torch.Tensor) -> None:
    t._eqb_amax = (amax, t._version)


def conv2d_forward(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], relu: bool,
                   mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[relu](conv2d(x, w) + bias), valid k x k, optionally zeroed where mask <= 0.  The result carries its per-sample max |y|
    (the 1x1 tensor-core layers scale their operands per sample; chained calls hand the maxima on, eqb_conv2d_forward_scaled)."""
    dev = _need_cuda(x, w)
    x_amax = _known_amax(x)
    x, w = _f32(x), _f32(w)
    b, cin, h, wd = x.shape
    n, cin2, k, k2 = w.shape
    if cin2 != cin or k != k2:
        raise ValueError(f"filter {tuple(w.shape)} does not match input {tuple(x.shape)}")
    bias = _f32(bias) if bias is not None else None
    mask = _f32(mask) if mask is not None else None
    y = torch.empty((b, n, h - k + 1, wd - k + 1), dtype=torch.float32, device=dev)
    if mask is not None and mask.shape != y.shape:
        raise ValueError("mask must have the output's shape")
    y_amax = torch.empty((b,), dtype=torch.float32, device=dev)
    _call("eqb_conv2d_forward_scaled", 1 if b else 0, dev, _ptr(x), _ptr(w), _ptr(bias) if bias is not None else None,
          _ptr(mask) if mask is not None else None, _ptr(y), b, cin, h, wd, n, k, int(relu),
          _ptr(x_amax) if x_amax is not None else None, _ptr(y_amax), _stream(dev))
    _record_amax(y, y_amax)
    return y


def conv2d_weight_grad(dy: torch.Tensor, x: torch.Tensor, k: int, with_bias_grad: bool = False):
    """dw[n, ci, ky, kx] = sum_{b,y,x} dy[b,n,y,x] x[b,ci,y+ky,x+kx]; with_bias_grad: also sum_{b,y,x} dy[b,n,y,x] (N,), from the
    same pass over dy -> (dw, dbias)."""
    dev = _need_cuda(dy, x)
    dy_amax, x_amax = _known_amax(dy), _known_amax(x)
    dy, x = _f32(dy), _f32(x)
    b, cin, h, wd = x.shape
    n = dy.shape[1]
    if tuple(dy.shape) != (b, n, h - k + 1, wd - k + 1):
        raise ValueError("dy does not match a valid k x k convolution of x")
    dw = torch.empty((n, cin, k, k), dtype=torch.float32, device=dev)
    db = torch.empty((n,), dtype=torch.float32, device=dev) if with_bias_grad else None
    _call("eqb_conv2d_weight_grad_scaled", 1 if b else 0, dev, _ptr(dy), _ptr(x), _ptr(dw), b, cin, h, wd, n, k,
          _ptr(dy_amax) if dy_amax is not None else None, _ptr(x_amax) if x_amax is not None else None,
          _ptr(db) if db is not None else None, _stream(dev))
    return (dw, db) if with_bias_grad else dw


def plane_sums(x: torch.Tensor) -> torch.Tensor:
    """(B, C, H, W) -> (B, C) sums over the plane."""
    dev = _need_cuda(x)
    x = _f32(x)
    b, c = x.shape[:2]
    out = torch.empty((b, c), dtype=torch.float32, device=dev)
    _call("eqb_plane_sums", 1 if b * c else 0, dev, _ptr(x), b * c, x[0, 0].numel(), _ptr(out), _stream(dev))
    return out


def group_mean_backward(dact: torch.Tensor, cout: int, out_hw: Tuple[int, int]) -> torch.Tensor:
    dev = _need_cuda(dact)
    dact = _f32(dact)
    b, g = dact.shape
    dy = torch.empty((b, cout * g, out_hw[0], out_hw[1]), dtype=torch.float32, device=dev)
    _call("eqb_group_mean_backward", 1 if b else 0, dev, _ptr(dact), _ptr(dy), b, cout, g, out_hw[0] * out_hw[1], _stream(dev))
    # dy[b, o*|G|+g, p] = dact[b, g] / (cout * P): its per-sample maximum costs nothing to state
    _record_amax(dy, (dact.abs().amax(dim=1) / float(cout * out_hw[0] * out_hw[1])).contiguous())
    return dy


def lift_filter_orbit_adjoint(dorbit: torch.Tensor, cout: int, num_rotations: int, reflect: bool) -> torch.Tensor:
    dev = _need_cuda(dorbit)
    dorbit = _f32(dorbit)
    n, cin, k, _ = dorbit.shape
    dw = torch.empty((cout, cin, k, k), dtype=torch.float32, device=dev)
    _call("eqb_lift_filter_orbit_adjoint", 1, dev, _ptr(dorbit), _ptr(dw), cout, cin, k, num_rotations, int(reflect), _stream(dev))
    return dw


def regular_filter_orbit_adjoint(dorbit: torch.Tensor, cout: int, num_rotations: int, reflect: bool) -> torch.Tensor:
    dev = _need_cuda(dorbit)
    dorbit = _f32(dorbit)
    g = num_rotations * (2 if reflect else 1)
    n, cing, k, _ = dorbit.shape
    cin = cing // g
    dw = torch.empty((cout, cin, g, k, k), dtype=torch.float32, device=dev)
    _call("eqb_regular_filter_orbit_adjoint", 1, dev, _ptr(dorbit), _ptr(dw), cout, cin, k, num_rotations, int(reflect), _stream(dev))
    return dw


class _GConvStackTrain(torch.autograd.Function):
    """CustomEquivariantNetwork.forward as an autograd node for TRAINING (custom_equivariant_networks.py:80-93): layer-wise
    forward that keeps the post-ReLU feature maps, backward on the N3 kernels (weight gradients, masked 1x1 data
    gradients, bias sums, filter-orbit adjoints).  Arguments: x, then (weights, bias) per layer, flattened."""

    @staticmethod
    def forward(ctx, x, num_rotations, reflect, *params):
        g = num_rotations * (2 if reflect else 1)
        layers = [(params[i], params[i + 1]) for i in range(0, len(params), 2)]
        cout = layers[0][0].shape[0]
        feats, filters = [], []
        h = x
        for li, (w, bias) in enumerate(layers):
            wx = lift_filter_orbit(w.detach(), num_rotations, reflect) if li == 0 else regular_filter_orbit(w.detach(), num_rotations, reflect)
            bx = bias.detach().repeat_interleave(g) if bias is not None else None        # channel = o*|G| + g
            feats.append(h)
            filters.append(wx)
            h = conv2d_forward(h, wx, bx, relu=li < len(layers) - 1)
        sums = plane_sums(h)                                                               # (B, cout*|G|)
        act = sums.reshape(-1, cout, g).sum(1) / float(cout * h.shape[-2] * h.shape[-1])
        ctx.cfg = (num_rotations, reflect, cout, tuple(h.shape[-2:]), [b is not None for _, b in layers])
        ctx.amax = [_known_amax(f) for f in feats]      # (python attributes do not survive save_for_backward)
        ctx.save_for_backward(*feats, *filters)
        return act

    @staticmethod
    def backward(ctx, dact):
        num_rotations, reflect, cout, out_hw, has_bias = ctx.cfg
        g = num_rotations * (2 if reflect else 1)
        nl = len(has_bias)
        feats, filters = ctx.saved_tensors[:nl], ctx.saved_tensors[nl:]
        dy = group_mean_backward(dact.contiguous(), cout, out_hw)
        grads = [None] * (2 * nl)
        for li in range(nl - 1, -1, -1):
            xin, wx = feats[li], filters[li]
            if ctx.amax[li] is not None:
                _record_amax(xin, ctx.amax[li])
            k = wx.shape[-1]
            want_b = has_bias[li] and ctx.needs_input_grad[3 + 2 * li + 1]
            want_w = ctx.needs_input_grad[3 + 2 * li]
            if want_b and not want_w:
                grads[2 * li + 1] = plane_sums(dy).sum(0).reshape(cout, g).sum(1)
            if want_w:
                dwx = conv2d_weight_grad(dy, xin, k, with_bias_grad=want_b)
                if want_b:
                    dwx, db = dwx                       # (the bias gradient comes out of the same pass over dy)
                    grads[2 * li + 1] = db.reshape(cout, g).sum(1)
                grads[2 * li] = (lift_filter_orbit_adjoint(dwx, cout, num_rotations, reflect) if li == 0
                                 else regular_filter_orbit_adjoint(dwx, cout, num_rotations, reflect))
            if li > 0:
                # data gradient through the 1x1 layer and the ReLU in front of it: conv with W^T, masked by the saved input
                wt = wx.reshape(wx.shape[0], wx.shape[1]).t().contiguous().reshape(wx.shape[1], wx.shape[0], 1, 1)
                dy = conv2d_forward(dy, wt, None, relu=False, mask=xin)
        return (None, None, None, *grads)


def gconv_stack_train(x: torch.Tensor, layers, num_rotations: int, reflect: bool) -> torch.Tensor:
    """Differentiable (in the layer parameters) group activations; `layers` = [(weights, bias or None), ...]."""
    for li, (w, _) in enumerate(layers):
        if li > 0 and w.shape[-1] != 1:
            raise NotImplementedError("training path: regular layers are 1 x 1 (all CustomEquivariantNetwork builds)")
    flat = [t for pair in layers for t in pair]
    return _GConvStackTrain.apply(x, num_rotations, reflect, *flat)


# ---- a4..a6 -------------------------------------------------------------------------------------
def _stack_dims(lift_w, reg_w, num_rotations, reflect):
    cout, cin, k, k2 = lift_w.shape
    g = num_rotations * (2 if reflect else 1)
    for rw in reg_w:
        if tuple(rw.shape) != (cout, cout, g, 1, 1):
            raise ValueError(f"regular layer weights must be ({cout},{cout},{g},1,1), got {tuple(rw.shape)}")
    return cout, cin, k, g, 1 + len(reg_w)


def gconv_stack_pack(lift_w: torch.Tensor, lift_b: Optional[torch.Tensor], reg_w: Sequence[torch.Tensor],
                     reg_b: Sequence[Optional[torch.Tensor]], num_rotations: int, reflect: bool) -> torch.Tensor:
    """Filter orbits as zero-padded K-major GEMM operands + expanded biases + folded last layer (uint8 buffer)."""
    dev = _need_cuda(lift_w, lift_b, *reg_w, *reg_b)
    lift_w = _f32(lift_w)
    lift_b = None if lift_b is None else _f32(lift_b)
    reg_w = [_f32(w) for w in reg_w]
    reg_b = [None if b is None else _f32(b) for b in reg_b]
    cout, cin, k, g, n_layers = _stack_dims(lift_w, reg_w, num_rotations, reflect)
    nbytes = native.lib().eqb_gconv_stack_packed_bytes(cin, cout, k, num_rotations, int(reflect), n_layers)
    if nbytes < 0:
        native.check(int(nbytes), "eqb_gconv_stack_packed_bytes")
    packed = torch.empty((int(nbytes),), dtype=torch.uint8, device=dev)
    n = max(len(reg_w), 1)
    wp = (C.c_void_p * n)(*[_ptr(t) for t in reg_w]) if reg_w else (C.c_void_p * 1)(None)
    bp = (C.c_void_p * n)(*[_ptr(t) for t in reg_b]) if reg_w else (C.c_void_p * 1)(None)
    n_gemm = max(n_layers - 1, 1)
    _call("eqb_gconv_stack_pack", 4 + 2 * n_gemm, dev, _ptr(lift_w), _ptr(lift_b), wp, bp, cin, cout, k, num_rotations,
          int(reflect), n_layers, _ptr(packed), int(nbytes), _stream(dev))
    return packed


def gconv_stack_run(x: torch.Tensor, packed: torch.Tensor, last_bias: Optional[torch.Tensor], cout: int, k: int,
                    num_rotations: int, reflect: bool, n_layers: int, x_absmax: Optional[torch.Tensor] = None,
                    select: bool = False) -> torch.Tensor:
    """x (B,Cin,H,W) -> group activations (B,|G|) with pre-packed parameters.
    x_absmax (B floats: max |x[b]|, as left by crop_resize_aa(with_absmax=True)) saves the per-image max pass.
    select: the finish kernel also does the group pool / select of group_pool_select() on its own output and the result
    carries it as `act._eqb_selection = (num_rotations, reflect, idx, rotation, reflection, onehot, stats)`, so the
    canonicalizer needs no select launch: 2 launches per call (fused stack, finish + select)."""
    dev = _need_cuda(x, packed, last_bias, x_absmax)
    _no_backward("gconv_stack_run (the fused inference stack)", x)
    x = _f32(x)
    last_bias = None if last_bias is None else _f32(last_bias)
    b, cin, h, w = x.shape
    g = num_rotations * (2 if reflect else 1)
    lib = native.lib()
    total = lib.eqb_gconv_stack_workspace_bytes(b, cin, h, w, cout, k, num_rotations, int(reflect), n_layers)
    if total < 0:
        native.check(int(total), "eqb_gconv_stack_workspace_bytes")
    scratch_bytes = int(total) - packed.numel()
    if scratch_bytes < 0:
        raise ValueError("packed parameter buffer does not belong to this network configuration")
    scratch = torch.empty((max(scratch_bytes, 16),), dtype=torch.uint8, device=dev)
    act = torch.empty((b, g), dtype=torch.float32, device=dev)
    if x_absmax is not None and (x_absmax.dtype != torch.float32 or x_absmax.numel() < b):
        raise ValueError("x_absmax must hold one float32 per image")
    if select:
        idx = torch.empty((b,), dtype=torch.int32, device=dev)
        rot = torch.empty((b,), dtype=torch.float32, device=dev)
        refl = torch.empty((b,), dtype=torch.float32, device=dev) if reflect else None
        onehot = torch.empty((b, g), dtype=torch.float32, device=dev)
        stats = torch.empty((5,), dtype=torch.float32, device=dev)
        _call("eqb_gconv_stack_run_select", 2 if x_absmax is not None else 3, dev, _ptr(x), _ptr(x_absmax), b, cin, h, w,
              _ptr(packed), _ptr(last_bias), cout, k, num_rotations, int(reflect), n_layers, _ptr(act), _ptr(idx), _ptr(rot),
              _ptr(refl), _ptr(onehot), _ptr(stats), _ptr(scratch), scratch_bytes, _stream(dev))
        act._eqb_selection = (num_rotations, bool(reflect), idx, rot, refl, onehot, stats)
    elif x_absmax is not None:
        _call("eqb_gconv_stack_run_scaled", 2, dev, _ptr(x), _ptr(x_absmax), b, cin, h, w, _ptr(packed), _ptr(last_bias), cout, k,
              num_rotations, int(reflect), n_layers, _ptr(act), _ptr(scratch), scratch_bytes, _stream(dev))
    else:
        _call("eqb_gconv_stack_run", 3, dev, _ptr(x), b, cin, h, w, _ptr(packed), _ptr(last_bias), cout, k, num_rotations,
              int(reflect), n_layers, _ptr(act), _ptr(scratch), scratch_bytes, _stream(dev))
    return act


def gconv_stack_forward(x: torch.Tensor, lift_w: torch.Tensor, lift_b: Optional[torch.Tensor],
                        reg_w: Sequence[torch.Tensor], reg_b: Sequence[Optional[torch.Tensor]],
                        num_rotations: int, reflect: bool) -> torch.Tensor:
    """pack + run in one go (what the reference does on every forward)."""
    if x.shape[1] != lift_w.shape[1]:
        raise ValueError(f"lift weights expect {lift_w.shape[1]} input channels, image has {x.shape[1]}")
    packed = gconv_stack_pack(lift_w, lift_b, reg_w, reg_b, num_rotations, reflect)
    cout, _, k, _, n_layers = _stack_dims(lift_w, reg_w, num_rotations, reflect)
    last_b = reg_b[-1] if len(reg_w) else None
    return gconv_stack_run(x, packed, last_b, cout, k, num_rotations, reflect, n_layers)


# ---- a7 -----------------------------------------------------------------------------------------
def conv_stack_forward(x: torch.Tensor, filters: Sequence[torch.Tensor], biases: Sequence[Optional[torch.Tensor]],
                       scales: Sequence[Optional[torch.Tensor]], shifts: Sequence[Optional[torch.Tensor]],
                       num_group: int) -> torch.Tensor:
    """x (B,Cin,H,W) -> group activations (B,|G|) through L valid k x k convs with expanded filters
    (N, Cin_l, k, k), N = Cout*|G|; affine + ReLU after every layer but the last."""
    dev = _need_cuda(x, *filters, *biases, *scales, *shifts)
    x = _f32(x)
    filters = [_f32(f) for f in filters]
    L = len(filters)
    if L < 1:
        raise ValueError("at least one layer expected")
    n, cin, k, k2 = filters[0].shape
    if k != k2 or n % num_group or cin != x.shape[1]:
        raise ValueError(f"first filter {tuple(filters[0].shape)} does not fit input {tuple(x.shape)} / |G| = {num_group}")
    for f in filters[1:]:
        if tuple(f.shape) != (n, n, k, k):
            raise ValueError(f"inner filters must be ({n},{n},{k},{k}), got {tuple(f.shape)}")

    def vecs(seq, count):
        seq = list(seq) + [None] * (count - len(seq))
        out = []
        for v in seq[:count]:
            if v is not None:
                v = _f32(v).reshape(-1)
                if v.numel() != n:
                    raise ValueError(f"per-channel vectors must have {n} entries")
            out.append(v)
        return out

    biases, scales, shifts = vecs(biases, L), vecs(scales, L), vecs(shifts, L)
    b, _, h, w = x.shape
    cout = n // num_group
    nbytes = native.lib().eqb_conv_stack_workspace_bytes(b, cin, h, w, cout, k, num_group, L)
    if nbytes < 0:
        native.check(int(nbytes), "eqb_conv_stack_workspace_bytes")
    ws = torch.empty((max(int(nbytes), 16),), dtype=torch.uint8, device=dev)
    act = torch.empty((b, num_group), dtype=torch.float32, device=dev)
    arr = lambda ts: (C.c_void_p * L)(*[_ptr(t) for t in ts])
    _call("eqb_conv_stack_forward", 3 * L + 1, dev, _ptr(x), b, cin, h, w, arr(filters), arr(biases), arr(scales), arr(shifts),
          cout, k, num_group, L, _ptr(act), _ptr(ws), int(nbytes), _stream(dev))
    return act


# ---- a9 + a13 -----------------------------------------------------------------------------------
def group_pool_select(act: torch.Tensor, num_rotations: int, reflect: bool, want_onehot: bool = True):
    """-> idx (B) int32, rotation (B) degrees, reflection (B) or None, onehot (B,|G|) or None,
    stats (5) = [sum CE, sum identity, B, mean CE, mean identity]."""
    dev = _need_cuda(act)
    act = _f32(act)
    b, g = act.shape
    if g != num_rotations * (2 if reflect else 1):
        raise ValueError(f"activations have {g} columns, group has {num_rotations * (2 if reflect else 1)} elements")
    idx = torch.empty((b,), dtype=torch.int32, device=dev)
    rot = torch.empty((b,), dtype=torch.float32, device=dev)
    refl = torch.empty((b,), dtype=torch.float32, device=dev) if reflect else None
    onehot = torch.empty((b, g), dtype=torch.float32, device=dev) if want_onehot else None
    stats = torch.empty((5,), dtype=torch.float32, device=dev)
    _call("eqb_group_pool_select", 1, dev, _ptr(act), b, num_rotations, int(reflect), _ptr(idx),
          _ptr(rot), _ptr(refl), _ptr(onehot), _ptr(stats), _stream(dev))
    return idx, rot, refl, onehot, stats


# ---- a10 / a11 / a12 ------------------------------------------------------------------------------
def _idx32(idx: torch.Tensor) -> torch.Tensor:
    if idx.dtype != torch.int32:
        idx = idx.to(torch.int32)
    return idx.contiguous()


def warp_canonicalize(x: torch.Tensor, idx: torch.Tensor, num_rotations: int, reflect: bool) -> torch.Tensor:
    dev = _need_cuda(x, idx)
    x = _f32(x)
    idx = _idx32(idx)
    b, c, h, w = x.shape
    if idx.numel() != b:
        raise ValueError("one group element per sample expected")
    y = torch.empty_like(x)
    _call("eqb_warp_canonicalize", 1, dev, _ptr(x), _ptr(y), _ptr(idx), b, c, h, w, num_rotations, int(reflect), _stream(dev))
    return y


def warp_invert(f: torch.Tensor, idx: torch.Tensor, num_rotations: int, reflect: bool, regular: bool) -> torch.Tensor:
    dev = _need_cuda(f, idx)
    f = _f32(f)
    idx = _idx32(idx)
    b, c, h, w = f.shape
    if idx.numel() != b:
        raise ValueError("one group element per sample expected")
    out = torch.empty_like(f)
    _call("eqb_warp_invert", 1, dev, _ptr(f), _ptr(out), _ptr(idx), b, c, h, w, num_rotations, int(reflect),
          native.REP_REGULAR if regular else native.REP_SCALAR, _stream(dev))
    return out


def warp_adjoint(grad_out: torch.Tensor, idx: torch.Tensor, num_rotations: int, reflect: bool, mode: int) -> torch.Tensor:
    """grad of warp_canonicalize (mode 0) / warp_invert scalar (1) / regular (2) with respect to its image argument."""
    dev = _need_cuda(grad_out, idx)
    grad_out = _f32(grad_out)
    idx = _idx32(idx)
    b, c, h, w = grad_out.shape
    grad_in = torch.empty_like(grad_out)
    _call("eqb_warp_adjoint", 1, dev, _ptr(grad_out), _ptr(grad_in), _ptr(idx), b, c, h, w, num_rotations, int(reflect),
          int(mode), _stream(dev))
    return grad_in


def warp_element_grad(x: torch.Tensor, grad_out: torch.Tensor, idx: torch.Tensor, num_rotations: int, reflect: bool,
                      mode: int):
    """(d loss / d rotation [degrees], d loss / d reflection or None) of warp_canonicalize (mode 0) / warp_invert (1, 2)."""
    dev = _need_cuda(x, grad_out, idx)
    x, grad_out, idx = _f32(x), _f32(grad_out), _idx32(idx)
    b, c, h, w = x.shape
    g_rot = torch.empty(b, dtype=torch.float32, device=dev)
    g_ref = torch.empty(b, dtype=torch.float32, device=dev) if reflect else None
    _call("eqb_warp_element_grad", 1 if b else 0, dev, _ptr(x), _ptr(grad_out), _ptr(idx), b, c, h, w, num_rotations,
          int(reflect), int(mode), _ptr(g_rot), _ptr(g_ref) if reflect else None, _stream(dev))
    return g_rot, g_ref


class _WarpFunction(torch.autograd.Function):
    """The two discrete warps as an autograd node: gradients with respect to the image (eqb_warp_adjoint) and to the
    group element -- rotation in degrees and reflection indicator (eqb_warp_element_grad) -- i.e. what torch autograd
    derives through kornia's rotate and the flip blend in the reference.  The integer index carries no gradient."""

    @staticmethod
    def forward(ctx, x, idx, rotation, reflection, num_rotations, reflect, mode):
        ctx.cfg = (num_rotations, reflect, mode)
        ctx.need_element = (rotation is not None and rotation.requires_grad) or (reflection is not None and reflection.requires_grad)
        ctx.save_for_backward(idx, x if ctx.need_element else None)
        if mode == 0:
            return warp_canonicalize(x, idx, num_rotations, reflect)
        return warp_invert(x, idx, num_rotations, reflect, mode == 2)

    @staticmethod
    def backward(ctx, grad_out):
        idx, x = ctx.saved_tensors
        num_rotations, reflect, mode = ctx.cfg
        grad_out = grad_out.contiguous()
        g_x = warp_adjoint(grad_out, idx, num_rotations, reflect, mode) if ctx.needs_input_grad[0] else None
        g_rot = g_ref = None
        if ctx.need_element and (ctx.needs_input_grad[2] or ctx.needs_input_grad[3]):
            g_rot, g_ref = warp_element_grad(x, grad_out, idx, num_rotations, reflect, mode)
            if not ctx.needs_input_grad[2]:
                g_rot = None
            if not ctx.needs_input_grad[3]:
                g_ref = None
        return g_x, None, g_rot, g_ref, None, None, None


def _element_grads(rotation, reflection):
    rot = rotation if (rotation is not None and rotation.requires_grad) else None
    ref = reflection if (reflection is not None and reflection.requires_grad) else None
    return rot, ref


def warp_canonicalize_autograd(x: torch.Tensor, idx: torch.Tensor, num_rotations: int, reflect: bool,
                               rotation: Optional[torch.Tensor] = None, reflection: Optional[torch.Tensor] = None) -> torch.Tensor:
    """warp_canonicalize that records an autograd node when the image or the (straight-through) element requires
    grad; inference calls stay node-free.  `rotation` (B, degrees) / `reflection` (B) only receive gradients: the
    forward value is decided by `idx`."""
    rot, ref = _element_grads(rotation, reflection)
    if torch.is_grad_enabled() and (x.requires_grad or rot is not None or ref is not None):
        return _WarpFunction.apply(x, idx, rot, ref.reshape(-1) if ref is not None else None, num_rotations, reflect, 0)
    return warp_canonicalize(x, idx, num_rotations, reflect)


def warp_invert_autograd(f: torch.Tensor, idx: torch.Tensor, num_rotations: int, reflect: bool, regular: bool,
                         rotation: Optional[torch.Tensor] = None, reflection: Optional[torch.Tensor] = None) -> torch.Tensor:
    rot, ref = _element_grads(rotation, reflection)
    if torch.is_grad_enabled() and (f.requires_grad or rot is not None or ref is not None):
        return _WarpFunction.apply(f, idx, rot, ref.reshape(-1) if ref is not None else None, num_rotations, reflect,
                                   2 if regular else 1)
    return warp_invert(f, idx, num_rotations, reflect, regular)


def orbit_expand(x: torch.Tensor, pad: int, out_size: int, num_rotations: int, reflect: bool) -> torch.Tensor:
    dev = _need_cuda(x)
    _no_backward("orbit_expand (gradient with respect to the image)", x)
    x = _f32(x)
    b, c, h, w = x.shape
    g = num_rotations * (2 if reflect else 1)
    oh, ow = (h, w) if c == 1 else (out_size, out_size)
    out = torch.empty((g * b, c, oh, ow), dtype=torch.float32, device=dev)
    _call("eqb_orbit_expand", 1, dev, _ptr(x), _ptr(out), b, c, h, w, pad, out_size, num_rotations, int(reflect), _stream(dev))
    return out


def orbit_rotate_nearest(x: torch.Tensor, num_rotations: int, reflect: bool) -> torch.Tensor:
    """(B,C,H,W) -> (|G|,B,C,H,W): the evaluation orbit of GroupInference (inference_utils.py:97-122), one launch."""
    dev = _need_cuda(x)
    x = _f32(x)
    b, c, h, w = x.shape
    g = num_rotations * (2 if reflect else 1)
    out = torch.empty((g, b, c, h, w), dtype=torch.float32, device=dev)
    _call("eqb_orbit_rotate_nearest", 1 if b else 0, dev, _ptr(x), _ptr(out), b, c, h, w, num_rotations, int(reflect), _stream(dev))
    return out


def _cosine_group_activations_raw(vec: torch.Tensor, ref: torch.Tensor, num_group: int) -> torch.Tensor:
    dev = _need_cuda(vec, ref)
    vec = _f32(vec)
    ref = _f32(ref).reshape(-1)
    rows, v = vec.shape
    if rows % num_group or ref.numel() != v:
        raise ValueError("vector_out must be (|G|*B, V) and the reference vector (1, V)")
    b = rows // num_group
    act = torch.empty((b, num_group), dtype=torch.float32, device=dev)
    _call("eqb_cosine_group_activations", 1, dev, _ptr(vec), _ptr(ref), _ptr(act), b, num_group, v, _stream(dev))
    return act


class _CosineActivations(torch.autograd.Function):
    """cosine_group_activations as an autograd node (the optimisation-based variant trains ANY torch network and the
    reference vector through it: discrete_group.py:475-481)."""

    @staticmethod
    def forward(ctx, vec, ref, num_group):
        ctx.save_for_backward(vec, ref)
        ctx.num_group = num_group
        return _cosine_group_activations_raw(vec, ref, num_group)

    @staticmethod
    def backward(ctx, dact):
        vec, ref = ctx.saved_tensors
        dev = vec.device
        v32, r32 = _f32(vec), _f32(ref).reshape(-1)
        rows, v = v32.shape
        b = rows // ctx.num_group
        dvec = torch.empty_like(v32) if ctx.needs_input_grad[0] else None
        dref = torch.empty(v, dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        _call("eqb_cosine_group_activations_backward", 1 if b else 0, dev, _ptr(v32), _ptr(r32), _ptr(_f32(dact)),
              _ptr(dvec) if dvec is not None else None, _ptr(dref) if dref is not None else None, b, ctx.num_group, v,
              _stream(dev))
        return dvec, (dref.reshape(ref.shape) if dref is not None else None), None


def cosine_group_activations(vec: torch.Tensor, ref: torch.Tensor, num_group: int) -> torch.Tensor:
    if torch.is_grad_enabled() and (vec.requires_grad or ref.requires_grad):
        return _CosineActivations.apply(vec, ref, num_group)
    return _cosine_group_activations_raw(vec, ref, num_group)


# ---- a14 .. a17 -----------------------------------------------------------------------------------
def _gram_schmidt3_raw(v: torch.Tensor, modified: bool = False) -> torch.Tensor:
    dev = _need_cuda(v)
    v = _f32(v)
    if v.dim() != 3 or v.shape[1:] != (3, 3):
        raise ValueError(f"expected (B,3,3) vectors, got {tuple(v.shape)}")
    out = torch.empty_like(v)
    _call("eqb_gram_schmidt3", 1, dev, _ptr(v), _ptr(out), v.shape[0], int(modified), _stream(dev))
    return out


def _so3_apply_raw(x: torch.Tensor, rot: torch.Tensor) -> torch.Tensor:
    dev = _need_cuda(x, rot)
    x = _f32(x)
    rot = _f32(rot)
    b, three, n = x.shape
    if three != 3 or tuple(rot.shape) != (b, 3, 3):
        raise ValueError("expected x (B,3,N) and rotation (B,3,3)")
    y = torch.empty_like(x)
    _call("eqb_so3_apply", 1, dev, _ptr(x), _ptr(rot), _ptr(y), b, n, _stream(dev))
    return y


def _e3_apply_raw(loc: torch.Tensor, vel: torch.Tensor, rot: torch.Tensor, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    dev = _need_cuda(loc, vel, rot, t)
    loc, vel, rot, t = _f32(loc), _f32(vel), _f32(rot), _f32(t)
    m = loc.shape[0]
    if tuple(loc.shape) != (m, 3) or tuple(vel.shape) != (m, 3) or tuple(rot.shape) != (m, 3, 3) or tuple(t.shape) != (m, 3):
        raise ValueError("expected loc, vel, t (M,3) and rotation (M,3,3)")
    lc, vc = torch.empty_like(loc), torch.empty_like(vel)
    _call("eqb_e3_apply", 1, dev, _ptr(loc), _ptr(vel), _ptr(rot), _ptr(t), _ptr(lc), _ptr(vc), m, _stream(dev))
    return lc, vc


def _e3_invert_raw(x: torch.Tensor, rot: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    dev = _need_cuda(x, rot, t)
    x, rot, t = _f32(x), _f32(rot), _f32(t)
    m = x.shape[0]
    if tuple(x.shape) != (m, 3) or tuple(rot.shape) != (m, 3, 3) or tuple(t.shape) != (m, 3):
        raise ValueError("expected x, t (M,3) and rotation (M,3,3)")
    y = torch.empty_like(x)
    _call("eqb_e3_invert", 1, dev, _ptr(x), _ptr(rot), _ptr(t), _ptr(y), m, _stream(dev))
    return y


def _optr(t):
    return _ptr(t) if t is not None else None


def _wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class _GramSchmidt3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, modified):
        ctx.save_for_backward(v)
        ctx.modified = modified
        return _gram_schmidt3_raw(v, modified)

    @staticmethod
    def backward(ctx, dR):
        (v,) = ctx.saved_tensors
        v32, g = _f32(v), _f32(dR)
        dv = torch.empty_like(v32)
        _call("eqb_gram_schmidt3_backward", 1 if v32.shape[0] else 0, v32.device, _ptr(v32), _ptr(g), _ptr(dv), v32.shape[0],
              int(ctx.modified), _stream(v32.device))
        return dv, None


def gram_schmidt3(v: torch.Tensor, modified: bool = False) -> torch.Tensor:
    return _GramSchmidt3.apply(v, modified) if _wants_grad(v) else _gram_schmidt3_raw(v, modified)


class _So3Apply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rot):
        ctx.save_for_backward(x, rot)
        return _so3_apply_raw(x, rot)

    @staticmethod
    def backward(ctx, dy):
        x, rot = ctx.saved_tensors
        x32, r32, g = _f32(x), _f32(rot), _f32(dy)
        b, _, n = x32.shape
        dx = torch.empty_like(x32) if ctx.needs_input_grad[0] else None
        dr = torch.empty_like(r32) if ctx.needs_input_grad[1] else None
        _call("eqb_so3_apply_backward", 1 if b * n else 0, x32.device, _ptr(x32), _ptr(r32), _ptr(g), _optr(dx), _optr(dr), b, n,
              _stream(x32.device))
        if dr is not None and b * n == 0:
            dr.zero_()
        return dx, dr


def so3_apply(x: torch.Tensor, rot: torch.Tensor) -> torch.Tensor:
    return _So3Apply.apply(x, rot) if _wants_grad(x, rot) else _so3_apply_raw(x, rot)


class _E3Apply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loc, vel, rot, t):
        ctx.save_for_backward(loc, vel, rot, t)
        return _e3_apply_raw(loc, vel, rot, t)

    @staticmethod
    def backward(ctx, dlc, dvc):
        loc, vel, rot, t = (_f32(a) for a in ctx.saved_tensors)
        m, dev = loc.shape[0], loc.device
        need = ctx.needs_input_grad
        dloc = torch.empty_like(loc) if need[0] else None
        dvel = torch.empty_like(vel) if need[1] else None
        dr = torch.empty_like(rot) if need[2] else None
        dt = torch.empty_like(t) if need[3] else None
        _call("eqb_e3_apply_backward", 1 if m else 0, dev, _ptr(loc), _ptr(vel), _ptr(rot), _ptr(t),
              _ptr(_f32(dlc)) if dlc is not None else None, _ptr(_f32(dvc)) if dvc is not None else None,
              _optr(dloc), _optr(dvel), _optr(dr), _optr(dt), m, _stream(dev))
        return dloc, dvel, dr, dt


def e3_apply(loc: torch.Tensor, vel: torch.Tensor, rot: torch.Tensor, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return _E3Apply.apply(loc, vel, rot, t) if _wants_grad(loc, vel, rot, t) else _e3_apply_raw(loc, vel, rot, t)


class _E3Invert(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rot, t):
        ctx.save_for_backward(x, rot)
        return _e3_invert_raw(x, rot, t)

    @staticmethod
    def backward(ctx, dy):
        x, rot = (_f32(a) for a in ctx.saved_tensors)
        m, dev = x.shape[0], x.device
        need = ctx.needs_input_grad
        g = _f32(dy)
        dx = torch.empty_like(x) if need[0] else None
        dr = torch.empty_like(rot) if need[1] else None
        dt = torch.empty_like(x) if need[2] else None
        _call("eqb_e3_invert_backward", 1 if m else 0, dev, _ptr(x), _ptr(rot), _ptr(g), _optr(dx), _optr(dr), _optr(dt), m,
              _stream(dev))
        return dx, dr, dt


def e3_invert(x: torch.Tensor, rot: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    return _E3Invert.apply(x, rot, t) if _wants_grad(x, rot, t) else _e3_invert_raw(x, rot, t)


def prior_stats_continuous(rep: torch.Tensor) -> torch.Tensor:
    """-> stats (5) = [sum (R-I)^2, B*d*d, 0, mse, 1 - mse]."""
    dev = _need_cuda(rep)
    rep = _f32(rep)
    b, d, d2 = rep.shape
    if d != d2:
        raise ValueError("expected square group-element matrices")
    stats = torch.empty((5,), dtype=torch.float32, device=dev)
    _call("eqb_prior_stats_continuous", 1, dev, _ptr(rep), b, d, _ptr(stats), _stream(dev))
    return stats


def regular_roll_shift(r: int, num_rotations: int) -> int:
    return int(native.lib().eqb_regular_roll_shift(r, num_rotations))


# ---- N1: frame-predicting networks ------------------------------------------------------------------
def vnsmall_forward(x: torch.Tensor, params: torch.Tensor, n_knn: int, bn_eps: float = 1e-5) -> torch.Tensor:
    """x (B,3,N) clouds -> (B,3,3) equivariant vectors; `params` = the flat block eqb_vnsmall_forward documents."""
    dev = _need_cuda(x, params)
    x, params = _f32(x), _f32(params)
    if x.dim() != 3 or x.shape[1] != 3:
        raise ValueError(f"expected clouds (B,3,N), got {tuple(x.shape)}")
    if params.numel() != native.lib().eqb_vnsmall_param_count():
        raise ValueError("parameter block has the wrong size")
    b, _, n = x.shape
    out = torch.empty((b, 3, 3), dtype=torch.float32, device=dev)
    ws = torch.empty((max(int(native.lib().eqb_vnsmall_workspace_bytes(b, n)), 8),), dtype=torch.uint8, device=dev)
    _call("eqb_vnsmall_forward", 2, dev, _ptr(x), b, n, _ptr(params), int(n_knn), float(bn_eps), _ptr(out), _ptr(ws),
          ws.numel(), _stream(dev))
    return out


def vndeepsets_forward(loc: torch.Tensor, vel: Optional[torch.Tensor], charges: Optional[torch.Tensor], edges: torch.Tensor,
                       params: torch.Tensor, in_dim: int, hidden: int, num_layers: int, feat_v: bool, feat_a: bool,
                       feat_c: bool, nonlinearity: int, layer_pool_mean: bool, final_pool_mean: bool,
                       canon_translation: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> rotation vectors (M,3,3), translation (M,3) for M = 5 S particle rows."""
    dev = _need_cuda(loc, vel, charges, edges, params)
    loc, params = _f32(loc), _f32(params)
    vel = None if vel is None else _f32(vel)
    charges = None if charges is None else _f32(charges).reshape(-1)
    m = loc.shape[0]
    if loc.dim() != 2 or loc.shape[1] != 3 or m % 5:
        raise ValueError(f"expected loc (5*S,3), got {tuple(loc.shape)}")
    s = m // 5
    if edges is None:
        edges = torch.zeros((2, 0), dtype=torch.int64, device=dev)
    if edges.dtype != torch.int64:
        edges = edges.to(torch.int64)
    if isinstance(edges, (list, tuple)):
        edges = torch.stack(list(edges))
    edges = edges.contiguous()
    e = edges.shape[1]
    lib = native.lib()
    if params.numel() != lib.eqb_vndeepsets_param_count(in_dim, hidden, num_layers):
        raise ValueError("parameter block has the wrong size")
    ws = torch.empty((int(lib.eqb_vndeepsets_workspace_bytes(s)),), dtype=torch.uint8, device=dev)
    rot = torch.empty((m, 3, 3), dtype=torch.float32, device=dev)
    trans = torch.empty((m, 3), dtype=torch.float32, device=dev)
    _call("eqb_vndeepsets_forward", 2, dev, _ptr(loc), _ptr(vel), _ptr(charges), _ptr(edges), e, s, _ptr(params), in_dim,
          hidden, num_layers, int(feat_v), int(feat_a), int(feat_c), int(nonlinearity), int(layer_pool_mean),
          int(final_pool_mean), int(canon_translation), _ptr(rot), _ptr(trans), _ptr(ws), ws.numel(), None, _stream(dev))
    return rot, trans


# ---- N2: continuous image warps ------------------------------------------------------------------------
def warp_affine(x: torch.Tensor, mats: torch.Tensor, refl: Optional[torch.Tensor], mats_forward: bool, pad: int,
                cx: float, cy: float) -> torch.Tensor:
    """y(dst) = x sampled at c + A (dst - c); mats (B,2,2) = M with A = M^-1 (mats_forward) or A itself."""
    dev = _need_cuda(x, mats, refl)
    x, mats = _f32(x), _f32(mats)
    refl = None if refl is None else _f32(refl).reshape(-1)
    b, c, h, w = x.shape
    if tuple(mats.shape) != (b, 2, 2) or (refl is not None and refl.numel() != b):
        raise ValueError("one 2x2 matrix (and one reflection flag) per sample expected")
    y = torch.empty_like(x)
    _call("eqb_warp_affine", 1, dev, _ptr(x), _ptr(y), _ptr(mats), _ptr(refl), int(mats_forward), b, c, h, w, int(pad),
          float(cx), float(cy), _stream(dev))
    return y



class _WarpAffineCanon(torch.autograd.Function):
    """eqb_warp_affine (canonicalize direction) as an autograd node in the sampling map `theta` (B,2,3: destination ->
    source, un-padded pixel coordinates) and the flip indicator.  The forward VALUE is the tested kernel on the rotation
    matrices; theta only routes the gradient (the caller builds it with torch algebra from the same matrices, so torch
    autograd continues through the matrix inverse and the reference's alpha / beta translation column)."""

    @staticmethod
    def forward(ctx, x, mats, refl_flags, pad, cx, cy, theta, refl_soft):
        y = warp_affine(x, mats.detach(), refl_flags, True, pad, cx, cy)
        ctx.save_for_backward(x, theta.detach(), refl_flags)
        ctx.pad = pad
        ctx.want_refl = refl_soft is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, theta, refl_flags = ctx.saved_tensors
        dev = x.device
        x32, g, th = _f32(x), _f32(dy), _f32(theta).reshape(-1, 6)
        b, c, h, w = x32.shape
        rf = _f32(refl_flags).reshape(-1) if refl_flags is not None else None
        gth = torch.empty((b, 6), dtype=torch.float32, device=dev)
        grf = torch.empty(b, dtype=torch.float32, device=dev) if (ctx.want_refl and rf is not None) else None
        _call("eqb_warp_affine_grad", 1 if b else 0, dev, _ptr(x32), _ptr(g), _ptr(th), _optr(rf), b, c, h, w, int(ctx.pad),
              _ptr(gth), _optr(grf), _stream(dev))
        return None, None, None, None, None, None, gth.reshape(b, 2, 3), grf


def warp_affine_canonicalize_autograd(x: torch.Tensor, mats: torch.Tensor, refl: Optional[torch.Tensor], pad: int, cx: float,
                                      cy: float, theta: torch.Tensor) -> torch.Tensor:
    """Continuous canonicalize warp whose gradient flows into `theta` (and `refl` when it requires grad)."""
    if x.requires_grad:
        raise NotImplementedError("the continuous warp is not differentiable with respect to the image yet")
    flags = None if refl is None else refl.detach().reshape(-1)
    soft = refl.reshape(-1) if (refl is not None and refl.requires_grad) else None
    return _WarpAffineCanon.apply(x, mats, flags, pad, cx, cy, theta, soft)
