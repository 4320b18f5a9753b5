"""Continuous-group frames: SO(3) point clouds and E(3) n-body systems on sm_100a kernels.

Mirrors
  gram_schmidt                                   equiadapt/common/utils.py:22-51
  ContinuousGroupPointcloudCanonicalization      equiadapt/pointcloud/canonicalization/continuous_group.py:14-81
  EquivariantPointcloudCanonicalization          ... :84-134
  EuclideanGroupNBody                            equiadapt/nbody/canonicalization/euclidean_group.py:8-157
The frame-predicting networks (VNSmall, VNDeepSets) are the caller's torch modules in this round
(SURVEY.md 8f row N1); what runs natively is Gram-Schmidt, the rotation/translation apply, its
inverse and the prior statistic.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from . import ops
from .canonicalizers_base import ContinuousGroupCanonicalization


def gram_schmidt(vectors: torch.Tensor) -> torch.Tensor:
    """Classical Gram-Schmidt on the three rows of (B,3,3) (common/utils.py:22-51): no eps, no handedness fix."""
    return ops.gram_schmidt3(vectors, modified=False)


class ContinuousGroupPointcloudCanonicalization(ContinuousGroupCanonicalization):
    def __init__(self, canonicalization_network: torch.nn.Module, canonicalization_hyperparams: Any):
        super().__init__(canonicalization_network)

    def get_groupelement(self, x: torch.Tensor) -> dict:
        raise NotImplementedError("get_groupelement method is not implemented")

    def canonicalize(self, x: torch.Tensor, targets: Optional[List] = None, **kwargs: Any):
        """x (B,3,N) -> R x  (continuous_group.py:51-81: bmm(x^T, R^T)^T)."""
        self.device = x.device
        group_element_dict = self.get_groupelement(x)
        return ops.so3_apply(x, group_element_dict["rotation"])


class EquivariantPointcloudCanonicalization(ContinuousGroupPointcloudCanonicalization):
    def __init__(self, canonicalization_network: torch.nn.Module, canonicalization_hyperparams: Any):
        super().__init__(canonicalization_network, canonicalization_hyperparams)

    def get_groupelement(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """continuous_group.py:107-134."""
        group_element_dict = {}
        out_vectors = self.canonicalization_network(x)
        if not hasattr(self, "canonicalization_info_dict"):
            self.canonicalization_info_dict = {}
        group_element_dict["rotation"] = gram_schmidt(out_vectors)
        self.canonicalization_info_dict["group_element_matrix_representation"] = group_element_dict["rotation"]
        self.canonicalization_info_dict["group_element"] = group_element_dict  # type: ignore
        return group_element_dict


class EuclideanGroupNBody(ContinuousGroupCanonicalization):
    def __init__(self, canonicalization_network: torch.nn.Module) -> None:
        super().__init__(canonicalization_network)

    def forward(self, x: torch.Tensor, targets: Optional[List] = None, **kwargs: Any):
        return self.canonicalize(x, None, **kwargs)

    def get_groupelement(self, nodes, loc, edges, vel, edge_attr, charges) -> Dict[str, torch.Tensor]:
        """euclidean_group.py:43-85; additionally stores R under "group_element_matrix_representation"
        so the inherited prior loss works (the reference raises KeyError there: SURVEY.md 8e)."""
        group_element_dict: Dict[str, torch.Tensor] = {}
        rotation_vectors, translation_vectors = self.canonicalization_network(nodes, loc, edges, vel, edge_attr, charges)
        rotation_matrix = self.modified_gram_schmidt(rotation_vectors)
        if not hasattr(self, "canonicalization_info_dict"):
            self.canonicalization_info_dict = {}
        group_element_dict["rotation_matrix"] = rotation_matrix
        group_element_dict["translation_vectors"] = translation_vectors
        group_element_dict["rotation_matrix_inverse"] = rotation_matrix.transpose(1, 2)
        self.canonicalization_info_dict["group_element"] = group_element_dict
        self.canonicalization_info_dict["group_element_matrix_representation"] = rotation_matrix
        return group_element_dict

    def canonicalize(self, x: torch.Tensor, targets: Optional[List] = None, **kwargs: Any):
        """(loc - t) R^T, vel R^T per particle row (euclidean_group.py:87-124); kwargs are unpacked
        POSITIONALLY as loc, edges, vel, edge_attr, charges like the reference (:104)."""
        self.device = x.device
        loc, edges, vel, edge_attr, charges = kwargs.values()
        group_element_dict = self.get_groupelement(x, loc, edges, vel, edge_attr, charges)
        canonical_loc, canonical_vel = ops.e3_apply(loc, vel, group_element_dict["rotation_matrix"],
                                                    group_element_dict["translation_vectors"])
        if loc.shape[0] == 1:  # the reference's .squeeze() drops the row dimension for a single row (quirk A.4-7)
            return canonical_loc.squeeze(), canonical_vel.squeeze()
        return canonical_loc, canonical_vel

    def invert_canonicalization(self, x_canonicalized_out: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        """x R + t (euclidean_group.py:126-137)."""
        rotation_matrix, translation_vectors, _ = self.canonicalization_info_dict["group_element"].values()
        return ops.e3_invert(x_canonicalized_out, rotation_matrix, translation_vectors)

    def modified_gram_schmidt(self, vectors: torch.Tensor) -> torch.Tensor:
        """euclidean_group.py:139-157."""
        return ops.gram_schmidt3(vectors, modified=True)
