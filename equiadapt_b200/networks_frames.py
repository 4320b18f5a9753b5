"""Frame-predicting vector-neuron networks on fused sm_100a kernels, inference only.

Same module tree and parameter names as the reference, so its checkpoints load unchanged:
  VNSmall        equiadapt/pointcloud/canonicalization_networks/equivariant_networks.py:79-150
                 (conv_pos / conv1 / conv2 = VNLinearLeakyReLU, bn1 = VNBatchNorm: vector_neuron_layers.py:210-324)
  VNDeepSets     equiadapt/nbody/canonicalization_networks/custom_equivariant_networks.py:13-172
                 (first_set_layer / set_layers.i = VNDeepSetLayer :175-252; VNLeakyReLU / VNSoftplus of
                 custom_group_equivariant_layers.py:7-99)
The sub-modules only HOLD parameters (their own forward raises): a network's forward is one C-ABI call
(eqb_vnsmall_forward / eqb_vndeepsets_forward) on a flat parameter block that is rebuilt when a parameter changes.
"""
from __future__ import annotations

from typing import Any, List, Tuple

import torch
import torch.nn as nn

from . import ops
from ._cache import PackedParameterCache


class _Holder(nn.Module):
    def forward(self, *a: Any, **k: Any):
        raise NotImplementedError("parameter holder: the layer runs fused inside its network's forward on the B200 path")


class VNBatchNorm(_Holder):
    """vector_neuron_layers.py:276-324 (bn1d for dim 3 / 4, bn2d for dim 5)."""

    def __init__(self, num_features: int, dim: int):
        super().__init__()
        self.dim = dim
        if dim in (3, 4):
            self.bn1d = nn.BatchNorm1d(num_features)
        elif dim == 5:
            self.bn2d = nn.BatchNorm2d(num_features)

    @property
    def bn(self) -> nn.Module:
        return self.bn2d if self.dim == 5 else self.bn1d


class VNLinearLeakyReLU(_Holder):
    """vector_neuron_layers.py:210-273."""

    def __init__(self, in_channels: int, out_channels: int, dim: int = 5, share_nonlinearity: bool = False,
                 negative_slope: float = 0.2):
        super().__init__()
        if share_nonlinearity:
            raise NotImplementedError("share_nonlinearity is not used by VNSmall and not covered")
        self.dim = dim
        self.negative_slope = negative_slope
        self.map_to_feat = nn.Linear(in_channels, out_channels, bias=False)
        self.batchnorm = VNBatchNorm(out_channels, dim=dim)
        self.map_to_dir = nn.Linear(in_channels, out_channels, bias=False)


class VNSmall(PackedParameterCache, nn.Module):
    """equivariant_networks.py:79-150: (B,3,N) clouds -> (B,3,3) equivariant vectors (pooling "mean")."""

    def __init__(self, hyperparams: Any):
        super().__init__()
        self.n_knn = hyperparams.n_knn
        self.pooling = hyperparams.pooling
        self.conv_pos = VNLinearLeakyReLU(3, 64 // 3, dim=5, negative_slope=0.0)
        self.conv1 = VNLinearLeakyReLU(64 // 3, 64 // 3, dim=4, negative_slope=0.0)
        self.bn1 = VNBatchNorm(64 // 3, dim=4)
        self.conv2 = VNLinearLeakyReLU(64 // 3, 12 // 3, dim=4, negative_slope=0.0)
        self.dropout = nn.Dropout(p=0.5)
        if self.pooling == "max":
            raise NotImplementedError('pooling "max" (VNMaxPool) is outside the B200 hot path (SURVEY.md section 2 row 11)')
        if self.pooling != "mean":
            raise ValueError(f"Pooling type {self.pooling} not supported")

    def _param_list(self) -> List[torch.Tensor]:
        out: List[torch.Tensor] = []
        for conv in (self.conv_pos, self.conv1):
            bn = conv.batchnorm.bn
            out += [conv.map_to_feat.weight, conv.map_to_dir.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var]
            if conv is self.conv1:
                b1 = self.bn1.bn
                out += [b1.weight, b1.bias, b1.running_mean, b1.running_var]
        bn = self.conv2.batchnorm.bn
        out += [self.conv2.map_to_feat.weight, self.conv2.map_to_dir.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        return out

    def forward(self, point_cloud: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("VNSmall on the B200 path is inference-only: call .eval()")
        plist = self._param_list()
        if not self._packed_current(plist):
            self._flat = torch.cat([p.detach().reshape(-1).float() for p in plist])
        return ops.vnsmall_forward(point_cloud, self._flat, self.n_knn, self.conv_pos.batchnorm.bn.eps)


class VNLeakyReLU(_Holder):
    """nbody custom_group_equivariant_layers.py:55-99."""

    def __init__(self, in_channels: int, share_nonlinearity: bool = False, negative_slope: float = 0.2):
        super().__init__()
        if share_nonlinearity:
            raise NotImplementedError("share_nonlinearity is not used by VNDeepSets and not covered")
        self.map_to_dir = nn.Linear(in_channels, in_channels, bias=False)
        self.negative_slope = negative_slope


class VNSoftplus(VNLeakyReLU):
    """nbody custom_group_equivariant_layers.py:7-52."""

    def __init__(self, in_channels: int, share_nonlinearity: bool = False, negative_slope: float = 0.0):
        super().__init__(in_channels, share_nonlinearity, negative_slope)


class VNDeepSetLayer(_Holder):
    """custom_equivariant_networks.py:175-252."""

    def __init__(self, in_channels: int, out_channels: int, nonlinearity: str, pooling: str = "sum", residual: bool = True,
                 dropout: float = 0.0):
        super().__init__()
        self.in_dim, self.out_dim, self.pooling, self.residual = in_channels, out_channels, pooling, residual
        self.nonlinearity, self.dropout = nonlinearity, dropout
        self.identity_linear = nn.Linear(in_channels, out_channels)
        self.pooling_linear = nn.Linear(in_channels, out_channels)
        self.dropout_layer = nn.Dropout(dropout)
        if nonlinearity == "softplus":
            self.nonlinear_function = VNSoftplus(out_channels, share_nonlinearity=False)
        elif nonlinearity == "relu":
            self.nonlinear_function = VNLeakyReLU(out_channels, share_nonlinearity=False, negative_slope=0.0)
        elif nonlinearity == "leakyrelu":
            self.nonlinear_function = VNLeakyReLU(out_channels, share_nonlinearity=False)
        else:
            raise ValueError(f"unknown nonlinearity {nonlinearity}")


class SequentialMultiple(nn.Sequential):
    """custom_equivariant_networks.py:255-280 (container only here)."""


_NONLIN = {"relu": 0, "leakyrelu": 1, "softplus": 2}


class VNDeepSets(PackedParameterCache, nn.Module):
    """custom_equivariant_networks.py:13-172: 5-particle systems -> (rotation vectors (M,3,3), translation (M,3))."""

    def __init__(self, hyperparams: Any, device: str = "cuda" if torch.cuda.is_available() else "cpu"):
        super().__init__()
        self.device = device
        self.prediction_mode = hyperparams.out_dim == 1
        if self.prediction_mode:
            raise NotImplementedError("prediction mode (out_dim == 1) is not a canonicalization network and not covered")
        self.model = "vndeepsets"
        self.hidden_dim = hyperparams.hidden_dim
        self.layer_pooling = hyperparams.layer_pooling
        self.final_pooling = hyperparams.final_pooling
        self.num_layers = hyperparams.num_layers
        self.nonlinearity = hyperparams.nonlinearity
        self.canon_feature = hyperparams.canon_feature
        self.canon_translation = hyperparams.canon_translation
        self.angular_feature = getattr(hyperparams, "angular_feature", False)
        self.dropout = hyperparams.dropout
        self.out_dim = hyperparams.out_dim
        self.in_dim = len(self.canon_feature)
        if self.canon_feature not in ("p", "pv", "pva", "pvc", "pvac"):
            raise ValueError(f"unknown canon_feature {self.canon_feature}")
        if self.out_dim != 4:
            raise NotImplementedError("out_dim must be 4 (three rotation vectors + one translation vector)")
        for name in (self.layer_pooling, self.final_pooling):
            if name not in ("mean", "sum"):
                raise NotImplementedError(f"pooling {name} is not covered (mean / sum are)")
        self.first_set_layer = VNDeepSetLayer(self.in_dim, self.hidden_dim, self.nonlinearity, self.layer_pooling, False,
                                              dropout=self.dropout)
        self.set_layers = SequentialMultiple(*[
            VNDeepSetLayer(self.hidden_dim, self.hidden_dim, self.nonlinearity, self.layer_pooling, dropout=self.dropout)
            for _ in range(self.num_layers - 1)])
        self.output_layer = nn.Linear(self.hidden_dim, self.out_dim)
        self.batch_size = hyperparams.batch_size
        self.to(device)

    def _param_list(self) -> List[torch.Tensor]:
        out: List[torch.Tensor] = []
        for layer in [self.first_set_layer] + list(self.set_layers):
            out += [layer.identity_linear.weight, layer.identity_linear.bias, layer.pooling_linear.weight,
                    layer.pooling_linear.bias, layer.nonlinear_function.map_to_dir.weight]
        out += [self.output_layer.weight, self.output_layer.bias]
        return out

    def forward(self, nodes: torch.Tensor, loc: torch.Tensor, edges: torch.Tensor, vel: torch.Tensor,
                edge_attr: torch.Tensor, charges: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.training:
            raise NotImplementedError("VNDeepSets on the B200 path is inference-only: call .eval()")
        plist = self._param_list()
        if not self._packed_current(plist):
            self._flat = torch.cat([p.detach().reshape(-1).float() for p in plist])
        if isinstance(edges, (list, tuple)):
            edges = torch.stack(list(edges))
        f = self.canon_feature
        return ops.vndeepsets_forward(loc, vel, charges, edges, self._flat, self.in_dim, self.hidden_dim, self.num_layers,
                                      "v" in f, "a" in f, "c" in f, _NONLIN[self.nonlinearity], self.layer_pooling == "mean",
                                      self.final_pooling == "mean", bool(self.canon_translation))
