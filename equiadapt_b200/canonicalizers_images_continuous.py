"""Continuous-group (SO(2) / O(2)) image canonicalizers on the sm_100a affine warp.

Mirrors equiadapt/images/canonicalization/continuous_group.py:
  ContinuousGroupImageCanonicalization        :15-240   (canonicalize :162-210)
  SteerableImageCanonicalization              :243-308
  OptimizedSteerableImageCanonicalization     :311-497  (group_augment :362-412)
The reference's flip blend -> Pad(edge) -> K.geometry.warp_affine -> CenterCrop chain (and, for the optimized
variant, Pad -> F.affine_grid / F.grid_sample -> CenterCrop) is ONE kernel (eqb_warp_affine) that never materialises
the padded image or the sampling grid.  The 2x2 / 3x3 algebra on the network's output vectors (normalise, Gram-Schmidt,
determinant) is a handful of scalars per sample and stays in torch, as does the consumer network.
Reference quirks kept: the rotation centre is (H_pad // 2, W_pad // 2) - half a pixel off the symmetric centre for even
sizes (:189-193); `canonicalize` flips the sign of the off-diagonal entries of group_element["rotation"] IN PLACE
(:180); `invert_canonicalization` goes through get_action_on_image_features with an empty group_info_dict and
therefore raises KeyError in the reference (:212-239, images/utils.py:54) - here it raises NotImplementedError.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from . import ops
from .canonicalizers_base import ContinuousGroupCanonicalization
from .canonicalizers_images import _center_crop_offset, _resize_output_size


def gram_schmidt_2d(vectors: torch.Tensor) -> torch.Tensor:
    """common/utils.py:22-51 on (B,2,2): rows orthonormalised in order, no eps."""
    v1 = vectors[:, 0]
    v1 = v1 / torch.norm(v1, dim=1, keepdim=True)
    v2 = vectors[:, 1] - torch.sum(vectors[:, 1] * v1, dim=1, keepdim=True) * v1
    v2 = v2 / torch.norm(v2, dim=1, keepdim=True)
    return torch.stack([v1, v2], dim=1)


class ContinuousGroupImageCanonicalization(ContinuousGroupCanonicalization):
    def __init__(self, canonicalization_network: torch.nn.Module, canonicalization_hyperparams: Any, in_shape: tuple):
        super().__init__(canonicalization_network)
        assert len(in_shape) == 3, "Input shape should be in the format (channels, height, width)"
        self.in_shape = tuple(in_shape)
        self.is_grayscale = in_shape[0] == 1
        self.pad_amount = 0 if self.is_grayscale else math.ceil(in_shape[-1] * 0.5)
        self.crop_canonization_size = (math.ceil(in_shape[-2] * canonicalization_hyperparams.input_crop_ratio),
                                       math.ceil(in_shape[-1] * canonicalization_hyperparams.input_crop_ratio))
        self.resize_shape = canonicalization_hyperparams.resize_shape
        self.group_info_dict: Dict[str, Any] = {}

    def get_groupelement(self, x: torch.Tensor) -> dict:
        raise NotImplementedError("get_groupelement method is not implemented")

    def transformations_before_canonicalization_network_forward(self, x: torch.Tensor) -> torch.Tensor:
        """continuous_group.py:106-120: CenterCrop + antialiased Resize (Identity for grayscale)."""
        if self.is_grayscale:
            return x
        h, w = x.shape[-2:]
        ch, cw = self.crop_canonization_size
        oh, ow = _resize_output_size(ch, cw, self.resize_shape)
        return ops.crop_resize_aa(x, _center_crop_offset(h, ch), _center_crop_offset(w, cw), ch, cw, oh, ow)

    def get_group_from_out_vectors(self, out_vectors: torch.Tensor) -> Tuple[dict, torch.Tensor]:
        """continuous_group.py:122-160."""
        group_element_dict = {}
        if self.group_type == "roto-reflection":
            rotoreflection_matrices = gram_schmidt_2d(out_vectors)
            determinant = (rotoreflection_matrices[:, 0, 0] * rotoreflection_matrices[:, 1, 1]
                           - rotoreflection_matrices[:, 0, 1] * rotoreflection_matrices[:, 1, 0])
            group_element_dict["reflection"] = (1 - determinant[:, None, None, None]) / 2
            reflection_indices = determinant < 0
            rotation_matrices = rotoreflection_matrices           # same storage, as in the reference (:147-148)
            rotation_matrices[reflection_indices, :, 1] *= -1
        else:
            rotation_matrices = self.get_rotation_matrix_from_vector(out_vectors[:, 0])
        group_element_dict["rotation"] = rotation_matrices
        return group_element_dict, (rotoreflection_matrices if self.group_type == "roto-reflection" else rotation_matrices)

    def get_rotation_matrix_from_vector(self, vectors: torch.Tensor) -> torch.Tensor:
        """continuous_group.py:264-277."""
        v1 = vectors / torch.norm(vectors, dim=1, keepdim=True)
        v2 = torch.stack([-v1[:, 1], v1[:, 0]], dim=1)
        return torch.stack([v1, v2], dim=1)

    def canonicalize(self, x: torch.Tensor, targets: Optional[List] = None, **kwargs: Any):
        """continuous_group.py:162-210 as one kernel."""
        self.device = x.device
        group_element_dict = self.get_groupelement(x)
        rotation_matrices = group_element_dict["rotation"]
        rotation_matrices[:, [0, 1], [1, 0]] *= -1                 # in place, as the reference (:180)
        refl = group_element_dict["reflection"].reshape(-1) if "reflection" in group_element_dict else None
        h, w = x.shape[-2:]
        p = self.pad_amount
        # centre of the PADDED image by integer division, x <- shape[-2], y <- shape[-1] (:189), in un-padded coordinates
        cx, cy = (h + 2 * p) // 2 - p, (w + 2 * p) // 2 - p
        if torch.is_grad_enabled() and (rotation_matrices.requires_grad or (refl is not None and refl.requires_grad)):
            # training: the sampling map the kernel uses, rebuilt with torch algebra exactly as the reference builds its
            # warp_affine matrix (:186-197), so autograd carries the kernel's d loss / d theta back to the network
            theta = self._sampling_map(rotation_matrices, h, w, p)
            return ops.warp_affine_canonicalize_autograd(x, rotation_matrices, refl, p, float(cx), float(cy), theta)
        return ops.warp_affine(x, rotation_matrices, refl, True, p, float(cx), float(cy))

    @staticmethod
    def _sampling_map(rotation_matrices: torch.Tensor, h: int, w: int, p: int) -> torch.Tensor:
        """(B,2,3) destination -> source map in un-padded pixel coordinates: the inverse of the reference's 2x3 matrix
        [R | ((1-a) cx - b cy, b cx + (1-a) cy)] (continuous_group.py:186-197, padded coordinates), shifted by the pad."""
        alpha, beta = rotation_matrices[:, 0, 0], rotation_matrices[:, 0, 1]
        cxp, cyp = (h + 2 * p) // 2, (w + 2 * p) // 2
        affine_part = torch.stack([(1 - alpha) * cxp - beta * cyp, beta * cxp + (1 - alpha) * cyp], dim=1)
        m = torch.cat([rotation_matrices, affine_part.unsqueeze(-1)], dim=-1)
        bottom = torch.tensor([0.0, 0.0, 1.0], device=m.device, dtype=m.dtype).expand(m.shape[0], 1, 3)
        inv = torch.linalg.inv(torch.cat([m, bottom], dim=1))
        a, s = inv[:, :2, :2], inv[:, :2, 2]
        pvec = torch.full((2,), float(p), device=m.device, dtype=m.dtype)
        return torch.cat([a, (a @ pvec + s - pvec).unsqueeze(-1)], dim=-1)

    def invert_canonicalization(self, x_canonicalized_out: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        raise NotImplementedError(
            "the reference's continuous-group invert_canonicalization is not functional (it reads num_rotations from "
            "an empty group_info_dict: continuous_group.py:212-239, images/utils.py:54) and is not reproduced")


class SteerableImageCanonicalization(ContinuousGroupImageCanonicalization):
    """continuous_group.py:243-308: the network returns (B, n, 2) equivariant vectors."""

    def __init__(self, canonicalization_network: torch.nn.Module, canonicalization_hyperparams: Any, in_shape: tuple):
        super().__init__(canonicalization_network, canonicalization_hyperparams, in_shape)
        self.group_type = canonicalization_network.group_type

    def get_groupelement(self, x: torch.Tensor) -> dict:
        x = self.transformations_before_canonicalization_network_forward(x)
        out_vectors = self.canonicalization_network(x)
        if not hasattr(self, "canonicalization_info_dict"):
            self.canonicalization_info_dict = {}
        group_element_dict, rep = self.get_group_from_out_vectors(out_vectors)
        self.canonicalization_info_dict["group_element_matrix_representation"] = rep
        self.canonicalization_info_dict["group_element"] = group_element_dict  # type: ignore
        return group_element_dict


class OptimizedSteerableImageCanonicalization(ContinuousGroupImageCanonicalization):
    """continuous_group.py:311-497: any network scores the image and one random augmentation of it."""

    def __init__(self, canonicalization_network: torch.nn.Module, canonicalization_hyperparams: Any, in_shape: tuple):
        super().__init__(canonicalization_network, canonicalization_hyperparams, in_shape)
        self.group_type = canonicalization_hyperparams.group_type

    def group_augment(self, x: torch.Tensor, angles: Optional[torch.Tensor] = None,
                      reflect: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """continuous_group.py:362-412.  `angles` (radians) / `reflect` (+-1) default to the reference's random draws;
        passing them makes the call reproducible."""
        batch_size = x.shape[0]
        dev = x.device
        if angles is None:
            angles = torch.rand(batch_size, device=dev) * 2 * torch.pi
        cos_a, sin_a = torch.cos(angles), torch.sin(angles)
        rotation_matrices = torch.zeros(batch_size, 2, 3, device=dev)
        rotation_matrices[:, :2, :2] = torch.stack((cos_a, -sin_a, sin_a, cos_a)).reshape(-1, 2, 2)
        if self.group_type == "roto-reflection":
            if reflect is None:
                reflect = torch.randint(0, 2, (batch_size,), device=dev).float() * 2 - 1
            rotation_matrices[:, 0, 0] *= reflect
        h, w = x.shape[-2:]
        if h != w:
            raise NotImplementedError("group_augment is covered for square images (affine_grid's normalised rotation "
                                      "shears non-square ones)")
        # affine_grid / grid_sample with align_corners=False on the padded square: a rotation about its symmetric centre
        aug = ops.warp_affine(x, rotation_matrices[:, :, :2].contiguous(), None, False, self.pad_amount,
                              0.5 * (h - 1), 0.5 * (w - 1))
        rotation_matrices[:, [0, 1], [1, 0]] *= -1
        return aug, rotation_matrices[:, :, :2]

    def get_groupelement(self, x: torch.Tensor) -> dict:
        self.device = x.device
        batch_size = x.shape[0]
        x_augmented, gt = self.group_augment(x)
        x_all = torch.cat([x, x_augmented], dim=0)
        x_all = self.transformations_before_canonicalization_network_forward(x_all)
        out_vectors_all = self.canonicalization_network(x_all).reshape(2 * batch_size, -1, 2)
        out_vectors, out_vectors_augmented = out_vectors_all.chunk(2, dim=0)
        if not hasattr(self, "canonicalization_info_dict"):
            self.canonicalization_info_dict = {}
        group_element_dict, rep = self.get_group_from_out_vectors(out_vectors)
        self.canonicalization_info_dict["group_element_matrix_representation"] = rep
        self.canonicalization_info_dict["group_element"] = group_element_dict  # type: ignore
        _, rep_aug = self.get_group_from_out_vectors(out_vectors_augmented)
        self.canonicalization_info_dict["group_element_matrix_representation_augmented"] = rep_aug
        self.canonicalization_info_dict["group_element_matrix_representation_augmented_gt"] = gt
        return group_element_dict

    def get_optimization_specific_loss(self) -> torch.Tensor:
        return torch.nn.functional.mse_loss(
            self.canonicalization_info_dict["group_element_matrix_representation_augmented"],
            self.canonicalization_info_dict["group_element_matrix_representation_augmented_gt"])
