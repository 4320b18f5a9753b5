"""Group-equivariant canonicalization network on the fused sm_100a conv stack.

Same module / parameter layout as the reference so checkpoints load unchanged
(`eqv_network.{0,2,4,...}.weights` / `.bias`):
  CustomEquivariantNetwork                      custom_equivariant_networks.py:14-93
  RotationEquivariantConvLift / RotoReflectionEquivariantConvLift / RotationEquivariantConv /
  RotoReflectionEquivariantConv                 custom_group_equivariant_layers.py:9-538
The layer modules own the parameters and expose the filter-orbit builders (eqb_*_filter_orbit);
the network's forward is ONE fused call (eqb_gconv_stack_forward) instead of
orbit-rebuild + conv2d + bias + ReLU per layer.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn as nn

from . import ops
from ._cache import PackedParameterCache


class _GroupConvParams(nn.Module):
    """Parameter holder shared by the four layer classes (init as the reference: kaiming-uniform a=sqrt(5)
    on the un-expanded weight, zero bias; custom_group_equivariant_layers.py:46-52, :266-274)."""

    reflect = False
    regular = False

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, num_rotations: int = 4,
                 stride: int = 1, padding: int = 0, bias: bool = True, device: str = "cuda"):
        super().__init__()
        if stride != 1 or padding != 0:
            raise NotImplementedError("the fused stack covers stride 1, padding 0 (all the reference network uses)")
        g = num_rotations * (2 if self.reflect else 1)
        shape = ((out_channels, in_channels, g, kernel_size, kernel_size) if self.regular
                 else (out_channels, in_channels, kernel_size, kernel_size))
        self.weights = nn.Parameter(torch.empty(*shape).to(device))
        torch.nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels).to(device))
            torch.nn.init.zeros_(self.bias)
        else:
            self.bias = None  # type: ignore
        self.in_channels, self.out_channels = in_channels, out_channels
        self.stride, self.padding = stride, padding
        self.num_rotations, self.kernel_size = num_rotations, kernel_size
        self.num_group_elements = g

    def filter_orbit(self) -> torch.Tensor:
        """(Cout*|G|, Cin[*|G|], k, k) expanded filter bank, channel = o*|G| + g."""
        if self.regular:
            return ops.regular_filter_orbit(self.weights, self.num_rotations, self.reflect)
        return ops.lift_filter_orbit(self.weights, self.num_rotations, self.reflect)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(
            "stand-alone layer forward is outside the B200 hot path: the layers run fused inside "
            "CustomEquivariantNetwork.forward (eqb_gconv_stack_forward)")


class RotationEquivariantConvLift(_GroupConvParams):
    def get_rotated_weights(self, weights: torch.Tensor = None, num_rotations: int = None) -> torch.Tensor:
        return self.filter_orbit()


class RotoReflectionEquivariantConvLift(_GroupConvParams):
    reflect = True

    def get_rotoreflected_weights(self, weights: torch.Tensor = None, num_rotations: int = None) -> torch.Tensor:
        return self.filter_orbit()


class RotationEquivariantConv(_GroupConvParams):
    regular = True

    def get_rotated_permuted_weights(self, weights: torch.Tensor = None, num_rotations: int = None) -> torch.Tensor:
        return self.filter_orbit()


class RotoReflectionEquivariantConv(_GroupConvParams):
    reflect = True
    regular = True

    def get_rotoreflected_permuted_weights(self, weights: torch.Tensor = None, num_rotations: int = None) -> torch.Tensor:
        return self.filter_orbit()


class CustomEquivariantNetwork(PackedParameterCache, nn.Module):
    """Lift(k x k) -> [ReLU -> GroupConv(1x1)] x (L-1) -> mean over (C,H,W) => (B,|G|).

    Unlike the reference class it also sets `group_type` / `num_rotations`, which
    GroupEquivariantImageCanonicalization reads (discrete_group.py:290-291; the reference omits them and
    raises AttributeError with network_type "custom": SURVEY.md A.4-1).
    """

    def __init__(self, in_shape: Tuple[int, int, int], out_channels: int, kernel_size: int,
                 group_type: str = "rotation", num_rotations: int = 4, num_layers: int = 1,
                 device: str = "cuda" if torch.cuda.is_available() else "cpu"):
        super().__init__()
        if group_type == "rotation":
            lift_cls, conv_cls = RotationEquivariantConvLift, RotationEquivariantConv
        elif group_type == "roto-reflection":
            lift_cls, conv_cls = RotoReflectionEquivariantConvLift, RotoReflectionEquivariantConv
        else:
            raise ValueError("group_type must be rotation or roto-reflection for now.")
        layers = [lift_cls(in_shape[0], out_channels, kernel_size, num_rotations, device=device)]
        for _ in range(num_layers - 1):
            layers.append(nn.ReLU())
            layers.append(conv_cls(out_channels, out_channels, 1, num_rotations, device=device))
        self.eqv_network = nn.Sequential(*layers)
        self.group_type = group_type
        self.num_rotations = num_rotations
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.fuse_select = True       # group pool / select rides on the finish kernel of the fused stack (ops.gconv_stack_run)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B,Cin,H,W) -> group activations (B,|G|): custom_equivariant_networks.py:80-93, one fused call."""
        mods = [m for m in self.eqv_network if isinstance(m, _GroupConvParams)]
        lift, regs = mods[0], mods[1:]
        reflect = self.group_type == "roto-reflection"
        # The packed operands (filter orbits, expanded biases, folded last layer) depend on the parameters
        # only: rebuilt when a parameter was modified in place or replaced, reused otherwise.  (The reference
        # rebuilds its orbits on every forward: custom_group_equivariant_layers.py:103, :349-351.)
        params = [p for m in mods for p in (m.weights, m.bias) if p is not None]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            # a gradient is wanted (the reference differentiates in train() AND eval()): layer-wise path that keeps the
            # feature maps, backward on the N3 kernels (gconv_train.cu).  Under no_grad / with frozen parameters: the
            # fused inference stack (no autograd graph).  The gradient with respect to the input image is not produced
            # (the reference's examples never use it): asking for it raises instead of returning silence.
            if x.requires_grad:
                raise NotImplementedError("CustomEquivariantNetwork on the B200 path does not differentiate with respect "
                                          "to its input image; detach() it (the reference's training loops do not need it)")
            return ops.gconv_stack_train(x, [(m.weights, m.bias) for m in mods], self.num_rotations, reflect)
        if not self._packed_current(params):
            # (key: data_ptr + version of every parameter -- see _cache.py for what that does and does not catch)
            self._packed = ops.gconv_stack_pack(lift.weights, lift.bias, [m.weights for m in regs],
                                                [m.bias for m in regs], self.num_rotations, reflect)
        last_bias = regs[-1].bias if regs else None
        # per-image max |x| left on the tensor by the canonicalizer's crop + resize kernel (ops.crop_resize_aa), if any
        amax = getattr(x, "_eqb_absmax", None)
        # the finish kernel also selects the group element (what the canonicalizer would launch eqb_group_pool_select for) and
        # leaves the selection on the result; a caller that only wants the activations ignores it
        return ops.gconv_stack_run(x, self._packed, last_bias, lift.out_channels, lift.kernel_size,
                                   self.num_rotations, reflect, len(mods), x_absmax=amax, select=self.fuse_select)


class ESCNNEquivariantNetwork(PackedParameterCache, nn.Module):
    """The reference's e2cnn network (escnn_networks.py:8-117) on the expanded-filter conv stack (eqb_conv_stack_forward).

    Reference layout: R2Conv(k) -> InnerBatchNorm -> ReLU -> PointwiseDropout(.5), (L-2) more such blocks,
    a final R2Conv(k); output reshaped to (B, Cout, |G|, H', W') and averaged over (Cout, H', W').
    In eval() every module is a dense op on tensors e2cnn caches (`R2Conv.filter`, `R2Conv.expanded_bias`), batch
    norm is one affine per field shared by its |G| channels, dropout is the identity.  This class holds exactly
    those tensors per layer:
        filters[l] (Cout*|G|, Cin_l, k, k), biases[l] (Cout*|G|),
        bn_weight / bn_bias / bn_running_mean / bn_running_var [l] (Cout)  for l < L-1
    and runs them as one C-ABI call.  e2cnn is not available to this build, so
      * a fresh instance is initialised with a group-symmetrised random filter bank (the filter-orbit construction
        of the custom layers, custom_group_equivariant_layers.py:62-90, :298-334, applied to k x k kernels), which is
        equivariant by construction; its values differ from e2cnn's steerable-basis initialisation;
      * `load_e2cnn(reference_network)` copies the cached tensors out of a reference ESCNNEquivariantNetwork in eval
        mode when e2cnn IS importable (checkpoint hand-over);
      * training-mode behaviour (batch statistics, dropout) is not implemented: forward raises in train().
    Parity of e2cnn's own basis expansion is unpinned; what is pinned is the dense arithmetic on
    the expanded tensors (oracle.reference_path.expanded_conv_network).
    """

    bn_eps = 1e-5

    def __init__(self, in_shape: tuple, out_channels: int, kernel_size: int, group_type: str = "rotation",
                 num_rotations: int = 4, num_layers: int = 1, device: str = "cuda" if torch.cuda.is_available() else "cpu"):
        super().__init__()
        if group_type not in ("rotation", "roto-reflection"):
            raise ValueError("group_type must be rotation or roto-reflection for now.")
        self.in_channels = in_shape[0]
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.group_type = group_type
        self.num_rotations = num_rotations
        self.num_layers = num_layers
        self.num_group_elements = num_rotations if group_type == "rotation" else 2 * num_rotations
        g, n = self.num_group_elements, out_channels * self.num_group_elements
        self.filters = nn.ParameterList()
        self.biases = nn.ParameterList()
        self.bn_weight = nn.ParameterList()
        self.bn_bias = nn.ParameterList()
        for l in range(num_layers):
            cin = self.in_channels if l == 0 else n
            self.filters.append(nn.Parameter(torch.zeros(n, cin, kernel_size, kernel_size, device=device)))
            self.biases.append(nn.Parameter(torch.zeros(n, device=device)))
            if l < num_layers - 1:
                self.bn_weight.append(nn.Parameter(torch.ones(out_channels, device=device)))
                self.bn_bias.append(nn.Parameter(torch.zeros(out_channels, device=device)))
                self.register_buffer(f"bn_running_mean_{l}", torch.zeros(out_channels, device=device))
                self.register_buffer(f"bn_running_var_{l}", torch.ones(out_channels, device=device))
        self._symmetrised_init(device)

    def _symmetrised_init(self, device: str) -> None:
        """filters[0] = lift orbit, filters[l>0] = regular orbit of kaiming-uniform base weights (needs CUDA: the
        orbit builders are kernels); on a CPU-only host the filters stay zero until loaded."""
        if not torch.device(device).type == "cuda":
            return
        reflect = self.group_type == "roto-reflection"
        g, k = self.num_group_elements, self.kernel_size
        with torch.no_grad():
            for l in range(self.num_layers):
                if l == 0:
                    base = torch.empty(self.out_channels, self.in_channels, k, k, device=device)
                    torch.nn.init.kaiming_uniform_(base, a=math.sqrt(5))
                    self.filters[l].copy_(ops.lift_filter_orbit(base, self.num_rotations, reflect))
                else:
                    base = torch.empty(self.out_channels, self.out_channels, g, k, k, device=device)
                    torch.nn.init.kaiming_uniform_(base, a=math.sqrt(5))
                    self.filters[l].copy_(ops.regular_filter_orbit(base, self.num_rotations, reflect))

    def load_e2cnn(self, reference_network: nn.Module) -> "ESCNNEquivariantNetwork":
        """Copy `filter` / `expanded_bias` / batch-norm statistics out of a reference ESCNNEquivariantNetwork
        (escnn_networks.py:66-91) that was put in eval() (e2cnn fills those buffers there)."""
        convs = [m for m in reference_network.eqv_network if hasattr(m, "expanded_bias")]
        bns = [m for m in reference_network.eqv_network if m.__class__.__name__ == "InnerBatchNorm"]
        if len(convs) != self.num_layers:
            raise ValueError(f"reference network has {len(convs)} conv layers, this one {self.num_layers}")
        with torch.no_grad():
            for l, c in enumerate(convs):
                self.filters[l].copy_(c.filter)
                self.biases[l].copy_(c.expanded_bias)
            for l, b in enumerate(bns):
                # one torch BatchNorm3d per representation size; all fields here are regular -> a single group
                bn = next(m for m in b.children() if hasattr(m, "running_mean"))
                self.bn_weight[l].copy_(bn.weight)
                self.bn_bias[l].copy_(bn.bias)
                getattr(self, f"bn_running_mean_{l}").copy_(bn.running_mean)
                getattr(self, f"bn_running_var_{l}").copy_(bn.running_var)
                self.bn_eps = bn.eps
        return self

    def folded_affine(self):
        """Per-channel (scale, shift) of every inner layer's batch norm in eval mode, expanded over the group axis;
        rebuilt only when a batch-norm tensor changed (it was ~8 eager torch launches on every forward)."""
        g = self.num_group_elements
        srcs = []
        for l in range(self.num_layers - 1):
            srcs += [self.bn_weight[l], self.bn_bias[l], getattr(self, f"bn_running_mean_{l}"), getattr(self, f"bn_running_var_{l}")]
        if self._packed_current(srcs) and getattr(self, "_folded", None) is not None and self._folded[2] == self.bn_eps:
            return self._folded[0], self._folded[1]
        scales, shifts = [], []
        for l in range(self.num_layers - 1):
            mean, var = getattr(self, f"bn_running_mean_{l}"), getattr(self, f"bn_running_var_{l}")
            sc = self.bn_weight[l] / torch.sqrt(var + self.bn_eps)
            scales.append(sc.repeat_interleave(g))
            shifts.append((self.bn_bias[l] - mean * sc).repeat_interleave(g))
        scales, shifts = [t.detach() for t in scales], [t.detach() for t in shifts]
        self._folded = (scales, shifts, self.bn_eps)
        return scales, shifts

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B,Cin,H,W) -> group activations (B,|G|): escnn_networks.py:93-117, one fused call per batch."""
        if self.training:
            raise NotImplementedError("ESCNNEquivariantNetwork on the B200 path is inference-only: call .eval() "
                                      "(batch statistics and dropout of train mode are not implemented)")
        scales, shifts = self.folded_affine()
        return ops.conv_stack_forward(x, list(self.filters), list(self.biases), scales, shifts, self.num_group_elements)
