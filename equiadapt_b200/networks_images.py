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


class CustomEquivariantNetwork(nn.Module):
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

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B,Cin,H,W) -> group activations (B,|G|): custom_equivariant_networks.py:80-93, one fused call."""
        mods = [m for m in self.eqv_network if isinstance(m, _GroupConvParams)]
        lift, regs = mods[0], mods[1:]
        reflect = self.group_type == "roto-reflection"
        # The packed operands (filter orbits, expanded biases, folded last layer) depend on the parameters
        # only: rebuilt when a parameter was modified in place or replaced, reused otherwise.  (The reference
        # rebuilds its orbits on every forward: custom_group_equivariant_layers.py:103, :349-351.)
        params = [p for m in mods for p in (m.weights, m.bias) if p is not None]
        key = tuple((p.data_ptr(), p._version) for p in params)
        if getattr(self, "_packed_key", None) != key:
            self._packed = ops.gconv_stack_pack(lift.weights, lift.bias, [m.weights for m in regs],
                                                [m.bias for m in regs], self.num_rotations, reflect)
            self._packed_key = key
        last_bias = regs[-1].bias if regs else None
        return ops.gconv_stack_run(x, self._packed, last_bias, lift.out_channels, lift.kernel_size,
                                   self.num_rotations, reflect, len(mods))
