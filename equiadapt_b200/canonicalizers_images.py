"""Discrete-group image canonicalizers on the sm_100a kernels.

Same classes, constructor arguments, methods, info-dict keys and errors as
equiadapt/images/canonicalization/discrete_group.py:
  DiscreteGroupImageCanonicalization (:20-259), GroupEquivariantImageCanonicalization (:262-317),
  OptimizedGroupEquivariantImageCanonicalization (:320-512),
and get_action_on_image_features / roll_by_gather of equiadapt/images/utils.py:8-94.
What the reference does with torchvision Pad/CenterCrop/Resize + kornia rotate/hflip (dozens of
library launches and three full-size temporaries per call) is here one kernel per step:
eqb_crop_resize_aa, eqb_group_pool_select, eqb_warp_canonicalize, eqb_warp_invert,
eqb_orbit_expand, eqb_cosine_group_activations.  Gradients: the two warps are differentiable in their IMAGE argument
(eqb_warp_adjoint; enough to train a prediction network through invert_canonicalization behind a frozen canonicalizer);
the group element and the canonicalization network are not differentiated in this round.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from . import ops
from .canonicalizers_base import DiscreteGroupCanonicalization


def _center_crop_offset(size: int, crop: int) -> int:
    # torchvision center_crop: int(round((size - crop) / 2.0)), Python rounding (half to even)
    return int(round((size - crop) / 2.0))


def _resize_output_size(h: int, w: int, size) -> Tuple[int, int]:
    """torchvision.transforms.Resize(size) on an (h, w) tensor image."""
    if isinstance(size, int):
        short, long = (w, h) if w <= h else (h, w)
        new_short, new_long = size, int(size * long / short)
        return (new_long, new_short) if w <= h else (new_short, new_long)
    size = tuple(int(s) for s in size)
    if len(size) == 1:
        return _resize_output_size(h, w, size[0])
    return size[0], size[1]


def group_element_to_index(group_element_dict: Dict[str, torch.Tensor], num_rotations: int) -> torch.Tensor:
    """rotation (degrees) [+ reflection 0/1] -> int32 group index; the inverse of the angle tables of
    discrete_group.py:110-133.  Used when the element dict did not come from our select kernel
    (e.g. a caller overrides get_groupelement, as the reference's own test fixture does)."""
    rot = group_element_dict["rotation"]
    idx = torch.round(rot / 360.0 * num_rotations).to(torch.int32) % num_rotations
    if "reflection" in group_element_dict:
        idx = idx + num_rotations * torch.round(group_element_dict["reflection"]).to(torch.int32)
    return idx


def get_action_on_image_features(feature_map: torch.Tensor, group_info_dict: dict, group_element_dict: dict,
                                 induced_rep_type: str = "regular") -> torch.Tensor:
    """Forward group action on a (B,C,H,W) feature map: equiadapt/images/utils.py:32-94."""
    num_rotations = group_info_dict["num_rotations"]
    num_group = group_info_dict["num_group"]
    assert len(feature_map.shape) == 4
    if induced_rep_type == "vector":
        raise NotImplementedError("Action for vector representation is not implemented")
    if induced_rep_type not in ("regular", "scalar"):
        raise ValueError("induced_rep_type must be regular, scalar or vector")
    if induced_rep_type == "regular":
        assert feature_map.shape[1] % num_group == 0
    idx = group_element_dict.get("index")
    if idx is None:
        idx = group_element_to_index(group_element_dict, num_rotations)
    reflect = "reflection" in group_element_dict
    return ops.warp_invert_autograd(feature_map, idx, num_rotations, reflect, induced_rep_type == "regular",
                                    rotation=group_element_dict["rotation"],
                                    reflection=group_element_dict["reflection"] if reflect else None)


class _ElementDict(dict):
    """group element dict {"rotation", ["reflection"]} that also remembers the int32 index the select
    kernel produced, without exposing it as a key (callers iterate / test keys of this dict)."""

    index: Optional[torch.Tensor] = None

    def get(self, key, default=None):
        if key == "index":
            return self.index
        return super().get(key, default)


class DiscreteGroupImageCanonicalization(DiscreteGroupCanonicalization):
    """discrete_group.py:20-259."""

    def __init__(self, canonicalization_network: torch.nn.Module, canonicalization_hyperparams: Any, in_shape: tuple):
        super().__init__(canonicalization_network)
        self.beta = canonicalization_hyperparams.beta
        assert len(in_shape) == 3, "Input shape should be in the format (channels, height, width)"
        self.in_shape = tuple(in_shape)
        self.is_grayscale = in_shape[0] == 1
        # geometry of the reference's transforms (discrete_group.py:60-92); grayscale: all Identity
        self.pad_amount = 0 if self.is_grayscale else math.ceil(in_shape[-1] * 0.5)
        self.crop_canonization_size = (
            math.ceil(in_shape[-2] * canonicalization_hyperparams.input_crop_ratio),
            math.ceil(in_shape[-1] * canonicalization_hyperparams.input_crop_ratio),
        )
        self.resize_shape = canonicalization_hyperparams.resize_shape

    # -- a9 -----------------------------------------------------------------------------------------
    def groupactivations_to_groupelement(self, group_activations: torch.Tensor) -> dict:
        """discrete_group.py:94-135: rotation in degrees (+ reflection 0/1) of the arg-max element."""
        sel = self._selection_for(group_activations)
        element = _ElementDict()
        sampled = self.gradient_trick != "straight_through"
        if sampled or self.training:
            # the reference's own expression (sum(onehot * angles), :110-133): in train() it carries the straight-through
            # gradient (and its 90-eps rounding), and with gumbel_softmax the element IS the sampled one-hot, not the
            # arg-max -- so the index that drives the warp is derived from that same one-hot
            onehot = self.groupactivations_to_groupelementonehot(group_activations)
            comp, ident = self._element_tables(onehot.device)
            element["rotation"] = torch.sum(onehot * comp, dim=-1)
            if self.group_type == "roto-reflection":
                element["reflection"] = torch.sum(onehot * ident, dim=-1)
            element.index = onehot.detach().argmax(dim=-1).to(torch.int32) if sampled else sel["idx"]
        else:
            element["rotation"] = sel["rotation"]
            if self.group_type == "roto-reflection":
                element["reflection"] = sel["reflection"]
            element.index = sel["idx"]
        return element

    def _element_tables(self, device: torch.device):
        """The reference's per-call constants (discrete_group.py:110-133): angles of the group elements and the reflection
        indicator, built on the CPU exactly as there (same fp32 bits) and kept on the device -- the reference uploads them on
        every call, a pageable host-to-device copy that also makes the step impossible to capture in a CUDA graph."""
        cache = self.__dict__.setdefault("_eqb_element_tables", {})
        key = (str(device), self.num_rotations, self.group_type)
        if key not in cache:
            angles = torch.linspace(0.0, 360.0, self.num_rotations + 1)[: self.num_rotations]
            comp = torch.cat([angles, angles]) if self.group_type == "roto-reflection" else angles
            ident = torch.cat([torch.zeros(self.num_rotations), torch.ones(self.num_rotations)])
            cache[key] = (comp.to(device), ident.to(device))
        return cache[key]

    def get_group_activations(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(
            "get_group_activations is not implemented for the DiscreteGroupImageCanonicalization class")

    def get_groupelement(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """discrete_group.py:152-172."""
        group_activations = self.get_group_activations(x)
        group_element_dict = self.groupactivations_to_groupelement(group_activations)
        if not hasattr(self, "canonicalization_info_dict"):
            self.canonicalization_info_dict = {}
        self.canonicalization_info_dict["group_element"] = group_element_dict  # type: ignore
        self.canonicalization_info_dict["group_activations"] = group_activations
        return group_element_dict

    # -- a3 -----------------------------------------------------------------------------------------
    def transformations_before_canonicalization_network_forward(self, x: torch.Tensor) -> torch.Tensor:
        """CenterCrop(ceil(H*ratio)) + antialiased Resize (discrete_group.py:174-188); Identity for C == 1."""
        if self.is_grayscale:
            return x
        h, w = x.shape[-2:]
        ch, cw = self.crop_canonization_size
        if ch > h or cw > w:
            raise NotImplementedError("input_crop_ratio > 1 (zero-padding CenterCrop) is not covered")
        oh, ow = _resize_output_size(ch, cw, self.resize_shape)
        # the kernel also leaves max |x_pre[b]| per image on the result for the conv stack's operand scaling
        return ops.crop_resize_aa(x, _center_crop_offset(h, ch), _center_crop_offset(w, cw), ch, cw, oh, ow,
                                  with_absmax=not torch.is_grad_enabled() or not x.requires_grad)

    # -- a10 ----------------------------------------------------------------------------------------
    def _element_index(self, group_element_dict) -> torch.Tensor:
        idx = group_element_dict.get("index") if isinstance(group_element_dict, _ElementDict) else None
        if idx is None:
            idx = group_element_to_index(group_element_dict, self.num_rotations)
        return idx

    def canonicalize(self, x: torch.Tensor, targets: Optional[List] = None, **kwargs: Any):
        """discrete_group.py:190-238: pad(edge) -> flip blend -> rotate(-angle) -> crop, as ONE kernel."""
        self.device = x.device
        if tuple(x.shape[1:]) != self.in_shape:
            # the reference builds Pad / CenterCrop from the constructor's in_shape (:60-92); the kernels derive them from
            # the tensor, so a different size would silently give a different transform than the reference's
            raise ValueError(f"input of shape {tuple(x.shape[1:])} does not match in_shape {self.in_shape}")
        group_element_dict = self.get_groupelement(x)
        if targets:
            raise NotImplementedError(
                "canonicalizing segmentation targets (boxes / masks, discrete_group.py:217-236) is outside the "
                "B200 hot path (SURVEY.md section 2 row 5)")
        idx = self._element_index(group_element_dict)
        reflect = "reflection" in group_element_dict.keys()
        # in training the element is the straight-through one (:103-121): the warp node hands its angle / flip
        # gradients back to it, as autograd through kornia's rotate does in the reference
        return ops.warp_canonicalize_autograd(x, idx, self.num_rotations, reflect,
                                              rotation=group_element_dict["rotation"],
                                              reflection=group_element_dict["reflection"] if reflect else None)

    # -- a11 ----------------------------------------------------------------------------------------
    def invert_canonicalization(self, x_canonicalized_out: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        """discrete_group.py:240-259."""
        induced_rep_type = kwargs.get("induced_rep_type", "regular")
        return get_action_on_image_features(
            feature_map=x_canonicalized_out,
            group_info_dict=self.group_info_dict,
            group_element_dict=self.canonicalization_info_dict["group_element"],  # type: ignore
            induced_rep_type=induced_rep_type,
        )


class GroupEquivariantImageCanonicalization(DiscreteGroupImageCanonicalization):
    """discrete_group.py:262-317: activations come from a group-equivariant network."""

    def __init__(self, canonicalization_network: torch.nn.Module, canonicalization_hyperparams: Any, in_shape: tuple):
        super().__init__(canonicalization_network, canonicalization_hyperparams, in_shape)
        self.group_type = canonicalization_network.group_type
        self.num_rotations = canonicalization_network.num_rotations
        self.num_group = self.num_rotations if self.group_type == "rotation" else 2 * self.num_rotations
        self.group_info_dict = {"num_rotations": self.num_rotations, "num_group": self.num_group}

    def get_group_activations(self, x: torch.Tensor) -> torch.Tensor:
        x = self.transformations_before_canonicalization_network_forward(x)
        return self.canonicalization_network(x)


class OptimizedGroupEquivariantImageCanonicalization(DiscreteGroupImageCanonicalization):
    """discrete_group.py:320-512: any (non-equivariant) network scores the |G|-expanded orbit."""

    def __init__(self, canonicalization_network: torch.nn.Module, canonicalization_hyperparams: Any, in_shape: tuple):
        super().__init__(canonicalization_network, canonicalization_hyperparams, in_shape)
        self.group_type = canonicalization_hyperparams.group_type
        self.num_rotations = canonicalization_hyperparams.num_rotations
        self.artifact_err_wt = canonicalization_hyperparams.artifact_err_wt
        self.num_group = self.num_rotations if self.group_type == "rotation" else 2 * self.num_rotations
        self.out_vector_size = canonicalization_network.out_vector_size
        # the reference treats resize_shape as an int here (quirk A.4-9)
        self.group_augment_in_shape = canonicalization_hyperparams.resize_shape
        self.group_augment_pad = 0 if self.is_grayscale else math.ceil(self.group_augment_in_shape * 0.5)
        self.reference_vector = torch.nn.Parameter(
            torch.randn(1, self.out_vector_size), requires_grad=canonicalization_hyperparams.learn_ref_vec)
        self.group_info_dict = {"num_rotations": self.num_rotations, "num_group": self.num_group}

    def group_augment(self, x: torch.Tensor) -> torch.Tensor:
        """(B,C,h,w) -> (|G|*B,C,r,r), group-major (discrete_group.py:387-427): one kernel."""
        return ops.orbit_expand(x, self.group_augment_pad, self.group_augment_in_shape, self.num_rotations,
                                self.group_type == "roto-reflection")

    def get_group_activations(self, x: torch.Tensor) -> torch.Tensor:
        """discrete_group.py:429-481."""
        x = self.transformations_before_canonicalization_network_forward(x)
        x_augmented = self.group_augment(x)
        vector_out = self.canonicalization_network(x_augmented)
        self.canonicalization_info_dict = {"vector_out": vector_out}
        if self.artifact_err_wt:
            # rotate every orbit member by a random element and back (:448-473); both are the
            # canonicalize warp (pad ceil(r/2), rotate, crop) with index r and N-r
            rotation_indices = torch.randint(0, self.num_rotations, (x_augmented.shape[0],)).to(self.device)
            fwd = rotation_indices.to(torch.int32)
            x_dummy = ops.warp_canonicalize(x_augmented, fwd, self.num_rotations, False)
            x_dummy = ops.warp_canonicalize(x_dummy, (self.num_rotations - fwd) % self.num_rotations,
                                            self.num_rotations, False)
            self.canonicalization_info_dict.update({"vector_out_dummy": self.canonicalization_network(x_dummy)})
        return ops.cosine_group_activations(vector_out, self.reference_vector, self.num_group)

    def get_optimization_specific_loss(self) -> torch.Tensor:
        """discrete_group.py:483-512 (tiny (B,|G|,V) algebra on the network output; stays in torch so the
        consumer network trains through it)."""
        vectors = self.canonicalization_info_dict["vector_out"]
        rotation_artifact_error = 0
        if self.artifact_err_wt:
            vectors_dummy = self.canonicalization_info_dict["vector_out_dummy"]
            rotation_artifact_error = torch.nn.functional.mse_loss(vectors_dummy, vectors)  # type: ignore
        vectors = vectors.reshape(self.num_group, -1, self.out_vector_size).permute((1, 0, 2))
        distances = vectors @ vectors.permute((0, 2, 1))
        mask = 1.0 - torch.eye(self.num_group).to(self.device)
        return torch.abs(distances * mask).mean() + self.artifact_err_wt * rotation_artifact_error
