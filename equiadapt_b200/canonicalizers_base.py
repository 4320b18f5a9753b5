"""The reference's abstract canonicalization surface, kept verbatim in names and call contract.

Mirrors equiadapt/common/basecanonicalization.py: BaseCanonicalization (:29-93),
IdentityCanonicalization (:96-179), DiscreteGroupCanonicalization (:182-311),
ContinuousGroupCanonicalization (:314-430).  What differs is underneath: the arg-max / one-hot /
soft-max / cross-entropy chain and the MSE-to-identity reduction are single sm_100a kernels
(eqb_group_pool_select, eqb_prior_stats_continuous) that also leave a 3-float statistic, which is
all-reduced once across ranks when torch.distributed is initialised.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from . import distributed as D
from . import ops


class BaseCanonicalization(torch.nn.Module):
    """forward(x, targets=None, **kw) -> canonicalize(...); owns the network and the info dict."""

    def __init__(self, canonicalization_network: torch.nn.Module):
        super().__init__()
        self.canonicalization_network = canonicalization_network
        self.canonicalization_info_dict: Dict[str, torch.Tensor] = {}
        # Opt-in: all-reduce the prior statistic across ranks -- ONE 3-float collective per forward,
        # shared by get_prior_regularization_loss() and get_identity_metric(), which then report the value the
        # single-process reference computes on the un-sharded batch.  OFF by default, as in the reference, where both
        # readers are rank-local: with it on they are COLLECTIVE calls (every rank must read them the same number of
        # times -- a rank-0-only logging read would hang).
        self.sync_prior_across_ranks = False
        # issue that collective asynchronously right after the kernel that produces the statistic, so it runs
        # beside the warp kernels instead of on the critical path of the loss read.  Off by default: it makes
        # every forward a collective call (all ranks must then call forward the same number of times).
        self.prefetch_prior_allreduce = False

    def _multi_rank(self) -> bool:
        return bool(self.sync_prior_across_ranks and D.world()[1] > 1)

    def _start_stats_allreduce(self, stats: torch.Tensor):
        """-> (tensor, work | None): the 3-float statistic summed over ranks, possibly still in flight."""
        if not self._multi_rank():
            return stats, None
        return D.allreduce_stats_async(stats[:3])

    @staticmethod
    def _finish_stats_allreduce(pending) -> torch.Tensor:
        tensor, work = pending
        if work is not None:
            work.wait()          # the current CUDA stream waits; the host does not block
        return tensor

    def _grad_scale(self, local_count: int, global_count: torch.Tensor):
        """Factor that turns the rank-local mean's gradient into this rank's share of the GLOBAL mean's gradient under
        DDP's average over ranks: (local_B / global_B) * world.  1 when the statistic is rank-local."""
        if not self._multi_rank():
            return None
        return (float(local_count) * D.world()[1]) / global_count

    def capture_step(self, x_example: torch.Tensor, fn=None, induced_rep_type: str = "scalar", with_prior: bool = True,
                     warmup: int = 3):
        """Record canonicalize -> fn -> invert_canonicalization [-> prior loss, identity metric] for inputs shaped like
        `x_example` as ONE CUDA graph (equiadapt_b200.graphed); -> callable(x) -> (y, z[, loss, metric])."""
        from . import graphed
        return graphed.capture_image_step(self, x_example, fn, induced_rep_type, with_prior, warmup)

    def forward(self, x: torch.Tensor, targets: Optional[List] = None, **kwargs: Any):
        return self.canonicalize(x, targets, **kwargs)

    def canonicalize(self, x: torch.Tensor, targets: Optional[List] = None, **kwargs: Any):
        raise NotImplementedError()

    def invert_canonicalization(self, x_canonicalized_out: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        raise NotImplementedError()


class IdentityCanonicalization(BaseCanonicalization):
    """No-op canonicalizer (basecanonicalization.py:96-179)."""

    def __init__(self, canonicalization_network: torch.nn.Module = torch.nn.Identity()):
        super().__init__(canonicalization_network)

    def canonicalize(self, x: torch.Tensor, targets: Optional[List] = None, **kwargs: Any):
        if targets:
            return x, targets
        return x

    def invert_canonicalization(self, x_canonicalized_out: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        return x_canonicalized_out

    def get_prior_regularization_loss(self) -> torch.Tensor:
        return torch.tensor(0.0)

    def get_identity_metric(self) -> torch.Tensor:
        return torch.tensor(1.0)


class _PriorCrossEntropy(torch.autograd.Function):
    """mean_b CE(act_b, class 0) whose VALUE comes from the select kernel's statistic; the backward is the closed form
    (softmax(act) - e_0) / B on the (B,|G|) activations -- what torch autograd gives for CrossEntropyLoss in the
    reference (basecanonicalization.py:290-301).  Rank-local statistic: the gradient of THIS rank's mean (the
    reference's loss; DDP averages the parameter gradients).  Synchronised statistic: `scale` = local_B / global_B *
    world makes DDP's average of these gradients the gradient of the global mean the value reports."""

    @staticmethod
    def forward(ctx, group_activations, value, scale=None):
        ctx.save_for_backward(group_activations, scale)
        return value.detach().clone()

    @staticmethod
    def backward(ctx, grad):
        act, scale = ctx.saved_tensors
        g = torch.softmax(act.detach().float(), dim=-1)
        g[:, 0] -= 1.0
        f = grad / act.shape[0]
        if scale is not None:
            f = f * scale
        return g * f, None, None


class _PriorMSE(torch.autograd.Function):
    """MSE(R, I) whose VALUE comes from the statistic kernel; backward 2 (R - I) / (B d d) on the (B,d,d) matrices, what
    torch autograd gives for MSELoss in the reference (basecanonicalization.py:390-408).  `scale`: see _PriorCrossEntropy."""

    @staticmethod
    def forward(ctx, rep, value, scale=None):
        ctx.save_for_backward(rep, scale)
        return value.detach().clone()

    @staticmethod
    def backward(ctx, grad):
        rep, scale = ctx.saved_tensors
        r = rep.detach().float()
        eye = torch.eye(r.shape[-1], device=r.device, dtype=r.dtype)
        f = 2.0 * grad / r.numel()
        if scale is not None:
            f = f * scale
        return (r - eye) * f, None, None


class DiscreteGroupCanonicalization(BaseCanonicalization):
    """Discrete groups: activations (B,|G|) -> one-hot group element, CE prior, identity metric."""

    def __init__(self, canonicalization_network: torch.nn.Module, beta: float = 1.0,
                 gradient_trick: str = "straight_through"):
        super().__init__(canonicalization_network)
        self.beta = beta
        self.gradient_trick = gradient_trick

    # filled by subclasses
    num_group: int
    num_rotations: int
    group_type: str

    def _select(self, group_activations: torch.Tensor):
        """One kernel: arg-max index, angle, reflection flag, one-hot and the prior statistic."""
        reflect = self.group_type == "roto-reflection"
        fused = getattr(group_activations, "_eqb_selection", None)
        if fused is not None and fused[0] == self.num_rotations and fused[1] == reflect:
            idx, rot, refl, onehot, stats = fused[2:]          # selected inside the network's finish kernel (same bits)
        else:
            idx, rot, refl, onehot, stats = ops.group_pool_select(group_activations, self.num_rotations, reflect)
        self._selected = {"activations": group_activations, "idx": idx, "rotation": rot, "reflection": refl,
                          "onehot": onehot, "stats": stats, "global": None}
        if self.prefetch_prior_allreduce:
            self._selected["global"] = self._start_stats_allreduce(stats)
        return self._selected

    def _selection_for(self, group_activations: torch.Tensor):
        sel = getattr(self, "_selected", None)
        if sel is None or sel["activations"] is not group_activations:
            sel = self._select(group_activations)
        return sel

    def groupactivations_to_groupelementonehot(self, group_activations: torch.Tensor) -> torch.Tensor:
        """basecanonicalization.py:221-256.  eval: hard one-hot from the kernel; train: the
        straight-through expression evaluates to the same values (soft - soft.detach() == 0)."""
        if self.gradient_trick == "straight_through":
            onehot = self._selection_for(group_activations)["onehot"]
            if self.training:
                soft = torch.nn.functional.softmax(self.beta * group_activations, dim=-1)
                return onehot + soft - soft.detach()
            return onehot
        if self.gradient_trick == "gumbel_softmax":
            return torch.nn.functional.gumbel_softmax(group_activations, tau=1, hard=True)
        raise ValueError(f"Gradient trick {self.gradient_trick} not implemented")

    def _discrete_stats(self) -> torch.Tensor:
        """This rank's [sum CE, sum identity, B]."""
        act = self.canonicalization_info_dict["group_activations"]
        return self._selection_for(act)["stats"]

    def _global_discrete_stats(self) -> torch.Tensor:
        """The statistic summed over ranks: one all-reduce per forward, cached for both readers."""
        act = self.canonicalization_info_dict["group_activations"]
        sel = self._selection_for(act)
        if sel["global"] is None:
            sel["global"] = self._start_stats_allreduce(sel["stats"])
        if isinstance(sel["global"], tuple):
            sel["global"] = self._finish_stats_allreduce(sel["global"])
        return sel["global"]

    def get_prior_regularization_loss(self) -> torch.Tensor:
        """mean_b CE(act_b, class 0) (basecanonicalization.py:290-301), from [sum CE, sum id, B]."""
        s = self._global_discrete_stats()
        value = s[0] / s[2] if self._multi_rank() else s[3]   # one rank: the kernel already divided
        act = self.canonicalization_info_dict["group_activations"]
        if torch.is_grad_enabled() and act.requires_grad:
            return _PriorCrossEntropy.apply(act, value, self._grad_scale(act.shape[0], s[2]))
        return value

    def get_identity_metric(self) -> torch.Tensor:
        """mean_b [argmax == 0] (basecanonicalization.py:303-311)."""
        s = self._global_discrete_stats()
        return s[1] / s[2] if self._multi_rank() else s[4]


class ContinuousGroupCanonicalization(BaseCanonicalization):
    """Continuous groups: MSE-to-identity prior on the (B,d,d) group-element representation."""

    def __init__(self, canonicalization_network: torch.nn.Module, beta: float = 1.0):
        super().__init__(canonicalization_network)
        self.beta = beta

    def canonicalizationnetworkout_to_groupelement(self, group_activations: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError()

    def _continuous_stats(self) -> torch.Tensor:
        """[sum (R - I)^2, B*d*d, 0] summed over ranks: one kernel + one all-reduce per forward, cached."""
        rep = self.canonicalization_info_dict["group_element_matrix_representation"]
        cached = getattr(self, "_rep_stats", None)
        if cached is None or cached[0] is not rep:
            local = ops.prior_stats_continuous(rep)
            cached = (rep, self._finish_stats_allreduce(self._start_stats_allreduce(local)))
            self._rep_stats = cached
        return cached[1]

    def get_prior_regularization_loss(self) -> torch.Tensor:
        """MSE(R, I) over B*d*d entries (basecanonicalization.py:390-408)."""
        s = self._continuous_stats()
        value = s[0] / s[1] if self._multi_rank() else s[3]
        rep = self.canonicalization_info_dict["group_element_matrix_representation"]
        if torch.is_grad_enabled() and rep.requires_grad:
            return _PriorMSE.apply(rep, value, self._grad_scale(rep.numel(), s[1]))
        return value

    def get_identity_metric(self) -> torch.Tensor:
        """1 - MSE(R, I) (basecanonicalization.py:410-430)."""
        return 1.0 - self.get_prior_regularization_loss()
