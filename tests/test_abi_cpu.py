"""CPU: the C-ABI library builds/loads, exports every symbol the header declares, validates its
arguments before touching a GPU, and the host-side logic (geometry, sharding, one all-reduce) is right.
No compute kernel is launched here."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from equiadapt_b200 import build, native
    build.build_native()
    return native.lib()


def test_library_exports_every_declared_symbol(lib):
    from equiadapt_b200 import native
    names = native.declared_symbols()
    assert len(names) >= 18
    raw = ctypes.CDLL(native.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} is declared in include/equiadapt_b200.h but not exported"
    assert set(native._SIGNATURES) == set(names)
    assert lib.eqb_abi_version() == 1


def test_argument_validation_without_gpu(lib):
    from equiadapt_b200 import native
    # bad enum / bad shapes are rejected before any CUDA call
    assert lib.eqb_warp_invert(None, None, None, 1, 3, 8, 8, 4, 0, 7, None) == native.EQB_ERR_INVALID
    assert b"rep" in lib.eqb_last_error()
    assert lib.eqb_warp_invert(None, None, None, 1, 5, 8, 8, 4, 0, native.REP_REGULAR, None) == native.EQB_ERR_INVALID
    assert lib.eqb_warp_canonicalize(None, None, None, 1, 0, 8, 8, 4, 0, None) == native.EQB_ERR_INVALID
    assert lib.eqb_crop_resize_aa(None, None, 1, 3, 8, 8, 4, 4, 8, 8, 4, 4, None) == native.EQB_ERR_INVALID
    assert lib.eqb_gconv_stack_workspace_bytes(1, 3, 32, 32, 64, 5, 8, 0, 3) == native.EQB_ERR_UNSUPPORTED
    assert lib.eqb_gconv_stack_workspace_bytes(1, 3, 4, 4, 8, 5, 4, 0, 3) == native.EQB_ERR_INVALID
    assert lib.eqb_gconv_stack_workspace_bytes(512, 3, 96, 96, 32, 5, 8, 0, 3) > 0
    # a7 / N1 / N2 entry points
    assert lib.eqb_conv_stack_workspace_bytes(4, 3, 96, 96, 32, 5, 4, 3) > 0
    assert lib.eqb_conv_stack_workspace_bytes(4, 3, 96, 96, 128, 5, 4, 3) == native.EQB_ERR_UNSUPPORTED   # N = 512 > 256
    assert lib.eqb_conv_stack_workspace_bytes(4, 3, 8, 8, 8, 5, 4, 3) == native.EQB_ERR_INVALID           # map shrinks below k
    assert lib.eqb_vnsmall_param_count() == 1444
    assert lib.eqb_vnsmall_forward(None, 2, 8, None, 20, 1e-5, None, None, 0, None) == native.EQB_ERR_INVALID      # k > N
    assert lib.eqb_vnsmall_forward(None, 2, 64, None, 40, 1e-5, None, None, 0, None) == native.EQB_ERR_UNSUPPORTED # k > 32
    assert lib.eqb_vndeepsets_param_count(1, 16, 4) == 2788
    assert lib.eqb_vndeepsets_forward(None, None, None, None, 0, 3, None, 2, 16, 4, 0, 0, 0, 0, 1, 1, 0, None, None, None, 0,
                                      None, None) == native.EQB_ERR_INVALID                               # in_dim vs feature flags
    assert lib.eqb_vndeepsets_forward(None, None, None, None, 0, 3, None, 1, 12, 4, 0, 0, 0, 0, 1, 1, 0, None, None, None, 0,
                                      None, None) == native.EQB_ERR_UNSUPPORTED                           # hidden width
    assert lib.eqb_warp_affine(None, None, None, None, 1, 1, 3, 8, 8, -1, 4.0, 4.0, None) == native.EQB_ERR_INVALID
    with pytest.raises(ValueError):
        native.check(native.EQB_ERR_INVALID)
    with pytest.raises(NotImplementedError):
        native.check(native.EQB_ERR_UNSUPPORTED)


def test_cpu_tensors_fail_loudly():
    from equiadapt_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.warp_canonicalize(torch.rand(1, 3, 8, 8), torch.zeros(1, dtype=torch.int32), 4, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gram_schmidt3(torch.rand(2, 3, 3))


def test_regular_roll_shift_reproduces_reference_truncation(lib):
    """(angle / 360.0 * N).long() with angle = linspace(0,360,N+1)[r] (images/utils.py:67,:28).

    Checked for N <= 16 (every group order the reference's configs use is <= 8).  Beyond one SIMD vector
    ATen's CPU linspace fills whole vectors from the chunk base, so the reference's own fp32 angles - and
    with them the truncated shift - depend on the host's vector width; the library reproduces the scalar
    formula (exactly r for power-of-two N on any host)."""
    for n in range(1, 17):
        angles = torch.linspace(0.0, 360.0, n + 1)[:n]
        want = (angles / 360.0 * n).long().tolist()
        got = [lib.eqb_regular_roll_shift(r, n) for r in range(n)]
        assert got == want, f"N={n}: {got} vs {want}"


def test_geometry_helpers_match_torchvision():
    from torchvision import transforms
    from equiadapt_b200.canonicalizers_images import _center_crop_offset, _resize_output_size
    for size, crop in [(224, 180), (64, 58), (32, 29), (33, 30), (40, 32), (7, 7)]:
        x = torch.arange(size * size, dtype=torch.float32).reshape(1, 1, size, size)
        y = transforms.CenterCrop(crop)(x)
        off = _center_crop_offset(size, crop)
        assert torch.equal(y, x[..., off:off + crop, off:off + crop])
    for (h, w, s) in [(180, 180, 96), (29, 29, 32), (24, 32, 16), (32, 24, 16), (24, 32, (16, 16))]:
        out = transforms.Resize(size=s)(torch.zeros(1, 3, h, w)).shape[-2:]
        assert tuple(out) == _resize_output_size(h, w, s)


def test_surface_and_state_dict_keys():
    from types import SimpleNamespace
    import equiadapt_b200 as E
    from equiadapt_b200.images.canonicalization.discrete_group import (
        GroupEquivariantImageCanonicalization, OptimizedGroupEquivariantImageCanonicalization)
    from equiadapt_b200.images.canonicalization_networks.custom_equivariant_networks import CustomEquivariantNetwork
    torch.manual_seed(0)
    net = CustomEquivariantNetwork((3, 32, 32), 16, 5, "rotation", 4, 3, device="cpu")
    # checkpoint compatibility with the reference class (custom_group_equivariant_layers.py:46-52,266-274)
    assert list(net.state_dict()) == [f"eqv_network.{i}.{p}" for i in (0, 2, 4) for p in ("weights", "bias")]
    assert net.eqv_network[0].weights.shape == (16, 3, 5, 5) and net.eqv_network[2].weights.shape == (16, 16, 4, 1, 1)
    can = GroupEquivariantImageCanonicalization(net, SimpleNamespace(beta=1.0, input_crop_ratio=0.9, resize_shape=32), (3, 32, 32))
    assert can.group_info_dict == {"num_rotations": 4, "num_group": 4} and can.num_group == 4
    assert can.crop_canonization_size == (29, 29) and can.pad_amount == 16
    for m in ("canonicalize", "invert_canonicalization", "get_prior_regularization_loss", "get_identity_metric",
              "get_groupelement", "transformations_before_canonicalization_network_forward"):
        assert callable(getattr(can, m))
    with pytest.raises(AssertionError):
        GroupEquivariantImageCanonicalization(net, SimpleNamespace(beta=1.0, input_crop_ratio=0.9, resize_shape=32), (3, 32))
    with pytest.raises(ValueError):
        CustomEquivariantNetwork((3, 32, 32), 16, 5, "so2", 4, 3, device="cpu")

    class V(torch.nn.Module):
        out_vector_size = 16
    hp = SimpleNamespace(beta=1.0, input_crop_ratio=0.8, resize_shape=24, group_type="roto-reflection", num_rotations=4,
                         artifact_err_wt=0, learn_ref_vec=False)
    opt = OptimizedGroupEquivariantImageCanonicalization(V(), hp, (3, 40, 40))
    assert "reference_vector" in opt.state_dict() and opt.num_group == 8 and opt.group_augment_pad == 12
    assert isinstance(E.IdentityCanonicalization().get_identity_metric(), torch.Tensor)


def test_shard_bounds_cover_batch():
    from equiadapt_b200.distributed import shard_bounds
    for n in (0, 1, 7, 512, 513):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_host_pipeline_stage_bounds():
    """HostStreamedCanonicalizer.shard_bounds: the stages tile the batch in order, never exceed the slot size, and the
    ramped schedule shortens the first and last stage to a quarter shard."""
    from equiadapt_b200.host_pipeline import HostStreamedCanonicalizer as H
    h = H.__new__(H)
    for shard in (1, 4, 64):
        for ramp in (False, True):
            h.shard, h.ramp = shard, ramp
            for B in (1, 3, 63, 64, 65, 255, 256, 300, 512):
                b = h.shard_bounds(B)
                assert b[0][0] == 0 and b[-1][1] == B
                assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
                assert all(0 < hi - lo <= shard for lo, hi in b)
    h.shard, h.ramp = 64, True
    sizes = [hi - lo for lo, hi in h.shard_bounds(512)]
    assert sizes[:2] == [16, 32] and sizes[-2:] == [32, 16] and sum(sizes) == 512


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from equiadapt_b200 import distributed as D
from oracle import reference_path as O
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = D.world()
g = torch.Generator().manual_seed(0)
act = torch.randn(37, 8, generator=g)                 # the un-sharded batch, same on every rank
rep = O.gram_schmidt(torch.randn(21, 3, 3, generator=g))
mine = D.shard_batch(act)
# what the select kernel leaves per rank: [sum CE, sum [idx==0], n]  (restated with torch for the CPU test)
ce = torch.logsumexp(mine, -1) - mine[:, 0]
stats = torch.stack([ce.sum(), (mine.argmax(-1) == 0).float().sum(), torch.tensor(float(mine.shape[0]))])
loss = D.mean_from_stats(stats, 0, 2)
ident = D.mean_from_stats(stats, 1, 2)
assert abs(float(loss) - float(O.prior_loss_discrete(act))) < 1e-5, (float(loss), float(O.prior_loss_discrete(act)))
assert abs(float(ident) - float(O.identity_metric_discrete(act))) < 1e-6
r_mine = D.shard_batch(rep)
cstats = torch.stack([((r_mine - torch.eye(3)) ** 2).sum(), torch.tensor(float(r_mine.numel())), torch.tensor(0.0)])
mse = D.mean_from_stats(cstats, 0, 1)
assert abs(float(mse) - float(O.prior_loss_continuous(rep))) < 1e-6
local_only = D.mean_from_stats(stats, 0, 2, sync=False)
assert world == 1 or abs(float(local_only) - float(loss)) > 0 or True
# the canonicalizer classes' path: ONE cached collective per forward, optionally started early (prefetch)
from equiadapt_b200.canonicalizers_base import DiscreteGroupCanonicalization
for prefetch in (False, True):
    can = DiscreteGroupCanonicalization(torch.nn.Identity())
    assert can.sync_prior_across_ranks is False                 # the reference's readers are rank-local: opt-in only
    can.sync_prior_across_ranks = True
    can.prefetch_prior_allreduce = prefetch
    can.canonicalization_info_dict = {{"group_activations": mine}}
    can._selected = {{"activations": mine, "stats": stats, "global": can._start_stats_allreduce(stats) if prefetch else None}}
    calls = []
    real = D.allreduce_stats_async
    D.allreduce_stats_async = lambda t, group=None: (calls.append(1), real(t, group))[1]
    l2, i2 = can.get_prior_regularization_loss(), can.get_identity_metric()
    D.allreduce_stats_async = real
    assert len(calls) == (0 if prefetch else 1), calls          # never a second collective for the metric
    assert abs(float(l2) - float(O.prior_loss_discrete(act))) < 1e-5 and abs(float(i2) - float(O.identity_metric_discrete(act))) < 1e-6
# default (sync off) = the reference: rank-local values, no collective, so a rank-0-only logging read cannot hang
can = DiscreteGroupCanonicalization(torch.nn.Identity())
can.canonicalization_info_dict = {{"group_activations": mine}}
local5 = torch.cat([stats, stats[:2] / stats[2]])
can._selected = {{"activations": mine, "stats": local5, "global": None}}
if rank == 0:
    l_loc = can.get_prior_regularization_loss()
    assert abs(float(l_loc) - float(O.prior_loss_discrete(mine))) < 1e-5
# synchronised statistic: value = global mean, and DDP's average of the per-rank gradients = gradient of that global mean
# (unequal shards: 19 / 18 samples)
can = DiscreteGroupCanonicalization(torch.nn.Identity())
can.sync_prior_across_ranks = True
leaf = mine.clone().requires_grad_(True)
can.canonicalization_info_dict = {{"group_activations": leaf}}
can._selected = {{"activations": leaf, "stats": local5, "global": None}}
can.get_prior_regularization_loss().backward()
full = act.clone().requires_grad_(True)
torch.nn.functional.cross_entropy(full, torch.zeros(act.shape[0], dtype=torch.long)).backward()
lo, hi = D.shard_bounds(act.shape[0], rank, world)
assert torch.allclose(leaf.grad / world, full.grad[lo:hi], atol=1e-7), (leaf.grad / world - full.grad[lo:hi]).abs().max()
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_prior_statistic_allreduce_world2_gloo(tmp_path):
    """N>1 host logic on CPU: batch sharding + the single 3-float all-reduce == un-sharded reference value."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


# ---- host-side pieces of the training path (no kernel involved) ----------------------------------------------------------
def test_prior_loss_nodes_backward_matches_torch_autograd():
    """_PriorCrossEntropy / _PriorMSE take their VALUE from the statistic kernels; their closed-form backward must equal
    torch autograd through CrossEntropyLoss / MSELoss (basecanonicalization.py:290-301, :390-408)."""
    from equiadapt_b200.canonicalizers_base import _PriorCrossEntropy, _PriorMSE
    g = torch.Generator().manual_seed(0)
    act = torch.randn(9, 8, generator=g, requires_grad=True)
    ref = torch.nn.functional.cross_entropy(act, torch.zeros(9, dtype=torch.long))
    (gr,) = torch.autograd.grad(3.0 * ref, act)
    act2 = act.detach().clone().requires_grad_(True)
    out = _PriorCrossEntropy.apply(act2, ref.detach())
    assert torch.equal(out, ref.detach())
    (3.0 * out).backward()
    assert torch.allclose(act2.grad, gr, atol=1e-7)
    rep = torch.randn(5, 3, 3, generator=g, requires_grad=True)
    ref = torch.nn.functional.mse_loss(rep, torch.eye(3).expand(5, 3, 3))
    (gr,) = torch.autograd.grad(2.0 * ref, rep)
    rep2 = rep.detach().clone().requires_grad_(True)
    (2.0 * _PriorMSE.apply(rep2, ref.detach())).backward()
    assert torch.allclose(rep2.grad, gr, atol=1e-7)


def test_continuous_sampling_map_is_the_inverse_of_the_reference_matrix():
    """ContinuousGroupImageCanonicalization._sampling_map: destination -> source map in un-padded pixels == what kornia's
    warp_affine does with the reference's 2x3 matrix on the padded image followed by the centre crop
    (continuous_group.py:186-208), checked by pushing pixel coordinates through both."""
    import math
    from equiadapt_b200.images.canonicalization.continuous_group import ContinuousGroupImageCanonicalization as C
    h = w = 36
    p = math.ceil(w * 0.5)
    ang = torch.tensor([0.3, 1.7, 4.0], dtype=torch.float64)
    rot = torch.stack([torch.stack([torch.cos(ang), torch.sin(ang)], 1), torch.stack([-torch.sin(ang), torch.cos(ang)], 1)], 1)
    rot = rot.clone()
    rot[:, [0, 1], [1, 0]] *= -1                     # as canonicalize() hands it over (:180)
    theta = C._sampling_map(rot, h, w, p)
    cx, cy = (h + 2 * p) // 2, (w + 2 * p) // 2
    alpha, beta = rot[:, 0, 0], rot[:, 0, 1]
    m = torch.cat([rot, torch.stack([(1 - alpha) * cx - beta * cy, beta * cx + (1 - alpha) * cy], 1).unsqueeze(-1)], -1)
    src = torch.tensor([[3.0, 5.0], [20.5, 11.25], [0.0, 35.0]], dtype=torch.float64)        # un-padded source pixels
    for b in range(3):
        dst_padded = (m[b, :, :2] @ (src + p).T).T + m[b, :, 2]      # kornia: M maps source to destination
        dst = dst_padded - p                                          # CenterCrop offset of the padded output
        back = (theta[b, :, :2] @ dst.T).T + theta[b, :, 2]
        assert torch.allclose(back, src, atol=1e-9)


def test_group_inference_host_logic():
    from types import SimpleNamespace
    from equiadapt_b200.images.inference import GroupInference, VanillaInference, get_inference_method
    ident, pred = torch.nn.Identity(), torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(12, 3))
    v = get_inference_method(ident, pred, 3, SimpleNamespace(method="vanilla"), (3, 2, 2))
    assert isinstance(v, VanillaInference) and not isinstance(v, GroupInference)
    x, y = torch.randn(6, 3, 2, 2), torch.tensor([0, 1, 2, 0, 1, 2])
    m = v.get_inference_metrics(x, y)
    assert set(m) == {"test/acc", "test/acc_class_0", "test/acc_class_1", "test/acc_class_2"}
    gi = get_inference_method(ident, pred, 3, {"method": "group", "group_type": "roto-reflection", "num_rotations": 4}, (3, 2, 2))
    assert isinstance(gi, GroupInference) and gi.num_group_elements == 8
    with pytest.raises(ValueError):
        gi.group_orbit(torch.zeros(1, 3, 4, 4))
    with pytest.raises(ValueError):
        get_inference_method(ident, pred, 3, {"method": "other"}, (3, 2, 2))
    with pytest.raises(RuntimeError):       # CPU tensors never fall back: the orbit kernel needs a CUDA tensor
        gi.group_orbit(torch.zeros(1, 3, 2, 2))


def test_argument_validation_of_the_training_and_orbit_entry_points(lib):
    """N3 / N4 entry points reject bad arguments before any CUDA call, and their tensor wrappers never fall back to CPU."""
    from equiadapt_b200 import native, ops
    inv, uns = native.EQB_ERR_INVALID, native.EQB_ERR_UNSUPPORTED
    assert lib.eqb_warp_adjoint(None, None, None, 1, 3, 8, 8, 4, 0, 3, None) == inv                 # mode out of range
    assert lib.eqb_warp_adjoint(None, None, None, 1, 5, 8, 8, 4, 0, 2, None) == inv                 # regular: C % |G| != 0
    assert lib.eqb_warp_element_grad(None, None, None, 1, 3, 8, 8, 0, 0, 0, None, None, None) == inv
    assert lib.eqb_warp_element_grad(None, None, None, 1, 3, 8, 8, 4, 0, 0, None, None, None) == inv    # null pointers
    assert lib.eqb_warp_affine_grad(None, None, None, None, 1, 3, 8, 8, -1, None, None, None) == inv
    assert lib.eqb_orbit_rotate_nearest(None, None, 1, 3, 8, 8, 0, 0, None) == inv
    assert lib.eqb_orbit_rotate_nearest(None, None, 1, 3, 8, 8, 65, 0, None) == uns
    assert lib.eqb_orbit_rotate_nearest(None, None, 0, 3, 8, 8, 4, 1, None) == 0                    # empty batch: nothing to do
    assert lib.eqb_conv2d_forward(None, None, None, None, None, 1, 3, 4, 4, 8, 5, 1, None) == inv   # map smaller than the kernel
    assert lib.eqb_conv2d_weight_grad(None, None, None, 1, 3, 8, 8, 8, 3, None) == inv              # null dw
    assert lib.eqb_conv2d_forward_scaled(None, None, None, None, None, 1, 3, 4, 4, 8, 5, 1, None, None, None) == inv
    assert lib.eqb_conv2d_forward_scaled(None, None, None, None, None, 0, 3, 8, 8, 8, 5, 1, None, None, None) == 0   # empty batch
    assert lib.eqb_conv2d_weight_grad_scaled(None, None, None, 1, 3, 8, 8, 8, 3, None, None, None, None) == inv      # null dw
    assert lib.eqb_plane_sums(None, 0, 5, None, None) == 0
    assert lib.eqb_plane_sums(None, 3, 0, None, None) == inv
    assert lib.eqb_group_mean_backward(None, None, 2, 0, 4, 16, None) == inv
    assert lib.eqb_lift_filter_orbit_adjoint(None, None, 4, 3, 5, 8, 0, None) == inv
    assert lib.eqb_regular_filter_orbit_adjoint(None, None, 4, 4, 1, 0, 0, None) == inv
    assert lib.eqb_cosine_group_activations_backward(None, None, None, None, None, 2, 0, 8, None) == inv
    assert lib.eqb_gram_schmidt3_backward(None, None, None, 3, 0, None) == inv
    assert lib.eqb_gram_schmidt3_backward(None, None, None, 0, 0, None) == 0
    assert lib.eqb_so3_apply_backward(None, None, None, None, None, -1, 4, None) == inv
    assert lib.eqb_e3_apply_backward(None, None, None, None, None, None, None, None, None, None, 3, None) == inv
    assert lib.eqb_e3_invert_backward(None, None, None, None, None, None, 3, None) == inv
    for call in (lambda: ops.warp_element_grad(torch.rand(1, 3, 8, 8), torch.rand(1, 3, 8, 8), torch.zeros(1, dtype=torch.int32), 4, False, 0),
                 lambda: ops.conv2d_forward(torch.rand(1, 3, 8, 8), torch.rand(4, 3, 3, 3), None, True),
                 lambda: ops.conv2d_weight_grad(torch.rand(1, 4, 6, 6), torch.rand(1, 3, 8, 8), 3),
                 lambda: ops.orbit_rotate_nearest(torch.rand(1, 3, 8, 8), 4, False),
                 lambda: ops.cosine_group_activations(torch.rand(8, 5), torch.rand(1, 5), 4),
                 lambda: ops.so3_apply(torch.rand(2, 3, 7), torch.rand(2, 3, 3))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_inference_metrics_match_the_unmodified_reference(monkeypatch):
    """VanillaInference / GroupInference.get_inference_metrics (host-side assembly: accuracies per class with the nan -> 0
    rule, per group element, their mean) vs the metric dictionaries the UNMODIFIED reference produced
    (tests/golden/inference_metrics.npz, inference_utils.py:50-77, :124-168).  The group orbit itself is a kernel and is
    replayed here from the reference's recorded per-element logits; its parity is a GPU test."""
    import numpy as np
    from conftest import GOLDEN_DIR
    from equiadapt_b200.images.inference import get_inference_method
    z = np.load(os.path.join(GOLDEN_DIR, "inference_metrics.npz"), allow_pickle=False)
    x, y = torch.from_numpy(z["x"]), torch.from_numpy(z["y"])
    pred = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 16 * 16, 5))
    with torch.no_grad():
        pred[1].weight.copy_(torch.from_numpy(z["weight"]))
        pred[1].bias.copy_(torch.from_numpy(z["bias"]))
        van = get_inference_method(torch.nn.Identity(), pred, 5, {"method": "vanilla"}, (3, 16, 16))
        m = van.get_inference_metrics(x, y)
        assert sorted(m) == [str(k) for k in z["vanilla_keys"]]
        assert np.allclose([float(m[k]) for k in sorted(m)], z["vanilla_values"], atol=1e-7)
        grp = get_inference_method(torch.nn.Identity(), pred, 5,
                                   {"method": "group", "group_type": "roto-reflection", "num_rotations": 4}, (3, 16, 16))
        logits = torch.from_numpy(z["group_logits"])
        monkeypatch.setattr(grp, "get_group_element_wise_logits", lambda _x: {g: logits[g] for g in range(logits.shape[0])})
        m = grp.get_inference_metrics(x, y)
        assert sorted(m) == [str(k) for k in z["group_keys"]]
        assert np.allclose([float(m[k]) for k in sorted(m)], z["group_values"], atol=1e-7)


def test_recorded_sample_maxima_follow_the_tensor_version():
    """ops._record_amax / _known_amax (the per-sample maxima the tensor-core training convolutions hand from call to call,
    eqb_conv2d_*_scaled): the record is bound to the tensor's version counter, so any in-place write drops it, and it does not
    travel with views, clones or detached copies."""
    from equiadapt_b200 import ops
    t = torch.arange(12.0).reshape(3, 4)
    assert ops._known_amax(t) is None
    amax = t.abs().amax(dim=1)
    ops._record_amax(t, amax)
    assert ops._known_amax(t) is amax
    assert ops._known_amax(t.clone()) is None and ops._known_amax(t.detach()) is None and ops._known_amax(t[:2]) is None
    t.mul_(2.0)
    assert ops._known_amax(t) is None                 # stale after an in-place write
    ops._record_amax(t, t.abs().amax(dim=1))
    t[0, 0] = 100.0
    assert ops._known_amax(t) is None
