"""Evaluation-time inference wrappers of the image examples (second half).

Mirror of examples/images/classification/inference_utils.py: `get_inference_method` (:8-26), `VanillaInference`
(:29-77) and `GroupInference` (:80-168) with the same constructor arguments, metric keys and per-element logits
dictionary.  What changes is how the orbit is produced: the reference pads, mirrors, rotates (torchvision, NEAREST)
and crops the batch once per group element (2|G| torchvision calls on the padded 1.8 H x 1.8 W image); here
`eqb_orbit_rotate_nearest` writes all |G| members in one launch, bit-exact with torchvision, and the canonicalizer +
prediction network then run once per member on views of that buffer (no copies), exactly as `forward` does in the
reference.
"""
from __future__ import annotations

import math
from typing import Any, Dict

import torch

from .. import ops


def _get(hp: Any, key: str):
    return hp[key] if isinstance(hp, dict) else getattr(hp, key)


def get_inference_method(canonicalizer: torch.nn.Module, prediction_network: torch.nn.Module, num_classes: int,
                         inference_hyperparams: Any, in_shape: tuple = (3, 32, 32)):
    """inference_utils.py:8-26."""
    method = _get(inference_hyperparams, "method")
    if method == "vanilla":
        return VanillaInference(canonicalizer, prediction_network, num_classes)
    if method == "group":
        return GroupInference(canonicalizer, prediction_network, num_classes, inference_hyperparams, in_shape)
    raise ValueError(f"{method} is not implemented for now.")


class VanillaInference:
    """inference_utils.py:29-77."""

    def __init__(self, canonicalizer: torch.nn.Module, prediction_network: torch.nn.Module, num_classes: int) -> None:
        self.canonicalizer = canonicalizer
        self.prediction_network = prediction_network
        self.num_classes = num_classes

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x_canonicalized = self.canonicalizer(x)
        return self.prediction_network(x_canonicalized)

    def _class_metrics(self, preds: torch.Tensor, y: torch.Tensor) -> Dict[str, Any]:
        acc_per_class = [(preds[y == i] == y[y == i]).float().mean() for i in range(self.num_classes)]
        acc_per_class = [torch.tensor(0.0) if math.isnan(acc) else acc for acc in acc_per_class]
        return {f"test/acc_class_{i}": max(acc, 0.0) for i, acc in enumerate(acc_per_class)}

    def get_inference_metrics(self, x: torch.Tensor, y: torch.Tensor) -> Dict[str, Any]:
        logits = self.forward(x)
        preds = logits.argmax(dim=-1)
        metrics: Dict[str, Any] = {"test/acc": (preds == y).float().mean()}
        metrics.update(self._class_metrics(preds, y))
        return metrics


class GroupInference(VanillaInference):
    """inference_utils.py:80-168."""

    def __init__(self, canonicalizer: torch.nn.Module, prediction_network: torch.nn.Module, num_classes: int,
                 inference_hyperparams: Any, in_shape: tuple = (3, 32, 32)):
        super().__init__(canonicalizer, prediction_network, num_classes)
        self.group_type = _get(inference_hyperparams, "group_type")
        self.num_rotations = _get(inference_hyperparams, "num_rotations")
        self.num_group_elements = self.num_rotations if self.group_type == "rotation" else 2 * self.num_rotations
        self.in_shape = tuple(in_shape)

    def group_orbit(self, x: torch.Tensor) -> torch.Tensor:
        """(|G|, B, C, H, W): every member of the evaluation orbit, rotations first (inference_utils.py:97-120)."""
        if tuple(x.shape[-2:]) != self.in_shape[-2:]:
            # the reference's Pad / CenterCrop are sized from in_shape at construction (:93-94)
            raise ValueError(f"GroupInference was built for images of {self.in_shape[-2:]}, got {tuple(x.shape[-2:])}")
        return ops.orbit_rotate_nearest(x, self.num_rotations, self.group_type == "roto-reflection")

    def get_group_element_wise_logits(self, x: torch.Tensor) -> Dict[int, torch.Tensor]:
        orbit = self.group_orbit(x)
        return {g: self.forward(orbit[g]) for g in range(orbit.shape[0])}

    def get_inference_metrics(self, x: torch.Tensor, y: torch.Tensor) -> Dict[str, Any]:
        logits_dict = self.get_group_element_wise_logits(x)
        acc_per_group_element = torch.tensor(
            [(logits.argmax(dim=-1) == y).float().mean() for logits in logits_dict.values()])
        metrics: Dict[str, Any] = {"test/group_acc": torch.mean(acc_per_group_element)}
        metrics.update({f"test/acc_group_element_{i}": acc_per_group_element[i] for i in range(self.num_group_elements)})
        preds = logits_dict[0].argmax(dim=-1)
        metrics.update({"test/acc": (preds == y).float().mean()})
        metrics.update(self._class_metrics(preds, y))
        return metrics
