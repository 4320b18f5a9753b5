"""ORACLE tooling: generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the dev container only (needs /root/reference):

    python -m oracle.make_golden            # from /root/repo

The reference package is imported from /root/reference through the stand-in modules in
oracle/shims/ (kornia -> oracle/kornia_restated.py, omegaconf, e2cnn, torch_scatter: none of
them is installed in this image).  Every array written is an input to, or an output of, a
reference class/function, named below with its file:line.  The fixtures are small on purpose
(tests/golden/ is a few MB in total) and are what the `-m "not gpu"` tests pin the oracle against and
what the `-m gpu` tests compare the CUDA path with on the GPU box, where /root/reference does
not exist.
"""
from __future__ import annotations

import os
import sys
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("EQUIADAPT_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), ROOT, REF]
    import equiadapt  # noqa: F401

    return equiadapt


def _np(d):
    out = {}
    for k, v in d.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def _save(name, d):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **_np(d))
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, keys={sorted(d)}")


class _HP(dict):
    __getattr__ = dict.__getitem__


def smooth_images(b, c, h, w, seed):
    """Smooth non-zero-mean synthetic images (wide argmax margins)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(b, c, max(h // 8, 2), max(w // 8, 2), generator=g)
    return torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=False) + 0.5


def golden_gram_schmidt():
    from equiadapt.common.utils import gram_schmidt  # common/utils.py:22-51

    torch.manual_seed(0)
    v = torch.randn(1, 3, 3)
    out = gram_schmidt(v)
    g = torch.Generator().manual_seed(7)
    vb = torch.randn(32, 3, 3, generator=g)
    _save("gram_schmidt", {"kat_in": v, "kat_out": out, "batch_in": vb, "batch_out": gram_schmidt(vb)})


def _image_case(name, group_type, num_rotations, in_shape, batch, out_channels, kernel_size, num_layers,
                crop_ratio, resize, beta, seed, regular_fields=1, fshape=(16, 16)):
    from equiadapt.images.canonicalization.discrete_group import GroupEquivariantImageCanonicalization
    from equiadapt.images.canonicalization_networks.custom_equivariant_networks import CustomEquivariantNetwork

    torch.manual_seed(seed)
    net = CustomEquivariantNetwork((in_shape[0], resize, resize), out_channels, kernel_size, group_type,
                                   num_rotations, num_layers, device="cpu")
    # quirk A.4-1: the reference class omits these two attributes the wrapper reads
    net.group_type, net.num_rotations = group_type, num_rotations
    with torch.no_grad():  # non-zero biases so the bias path is pinned too
        for m in net.eqv_network:
            if hasattr(m, "bias") and m.bias is not None:
                m.bias.uniform_(-0.05, 0.05)
    hp = _HP(beta=beta, input_crop_ratio=crop_ratio, resize_shape=resize)
    can = GroupEquivariantImageCanonicalization(net, hp, in_shape).eval()
    x = smooth_images(batch, *in_shape, seed=seed + 1)
    num_group = can.num_group
    gen = torch.Generator().manual_seed(seed + 2)
    d = {"x": x, "group_type": group_type, "num_rotations": num_rotations, "in_shape": np.array(in_shape),
         "crop_ratio": crop_ratio, "resize": resize, "beta": beta, "num_layers": num_layers}
    for i, m in enumerate(net.eqv_network):
        if hasattr(m, "weights"):
            d[f"w{i}"] = m.weights
            d[f"b{i}"] = m.bias
    with torch.no_grad():
        # discrete_group.py:174-188
        d["x_pre"] = can.transformations_before_canonicalization_network_forward(x)
        # custom_group_equivariant_layers.py:62-90 / :298-334 (orbit of the first two layers)
        lift = net.eqv_network[0]
        d["orbit_lift"] = (lift.get_rotated_weights(lift.weights, num_rotations) if group_type == "rotation"
                           else lift.get_rotoreflected_weights(lift.weights, num_rotations))
        if num_layers > 1:
            reg = net.eqv_network[2]
            d["orbit_reg"] = (reg.get_rotated_permuted_weights(reg.weights, num_rotations) if group_type == "rotation"
                              else reg.get_rotoreflected_permuted_weights(reg.weights, num_rotations))
        # full forward: discrete_group.py:190-238
        y = can(x)
        d["act"] = can.canonicalization_info_dict["group_activations"]
        d["rotation"] = can.canonicalization_info_dict["group_element"]["rotation"]
        if group_type == "roto-reflection":
            d["reflection"] = can.canonicalization_info_dict["group_element"]["reflection"]
        d["x_canon"] = y
        # basecanonicalization.py:290-311
        d["prior_loss"] = can.get_prior_regularization_loss()
        d["identity_metric"] = can.get_identity_metric()
        # images/utils.py:32-94 through discrete_group.py:240-259
        f_reg = torch.randn(batch, regular_fields * num_group, *fshape, generator=gen)
        f_sca = torch.randn(batch, 3, *fshape, generator=gen)
        d["f_regular"], d["f_scalar"] = f_reg, f_sca
        d["inv_regular"] = can.invert_canonicalization(f_reg, induced_rep_type="regular")
        d["inv_scalar"] = can.invert_canonicalization(f_sca, induced_rep_type="scalar")

        # every group element, by patching get_groupelement the way the reference's own
        # fixture does (tests/images/canonicalization/test_continuous_group.py:94-121)
        idx = torch.arange(batch) % num_group
        angles = torch.linspace(0.0, 360.0, num_rotations + 1)[:num_rotations]
        forced = {"rotation": angles[idx % num_rotations]}
        if group_type == "roto-reflection":
            forced["reflection"] = (idx >= num_rotations).float()
        with mock.patch.object(can, "get_groupelement", return_value=forced):
            d["forced_idx"] = idx
            d["forced_canon"] = can.canonicalize(x)
        can.canonicalization_info_dict["group_element"] = forced
        d["forced_inv_regular"] = can.invert_canonicalization(f_reg, induced_rep_type="regular")
        d["forced_inv_scalar"] = can.invert_canonicalization(f_sca, induced_rep_type="scalar")
    _save(name, d)


def golden_optimized(name, group_type, num_rotations, in_shape, batch, resize, crop_ratio, seed):
    from equiadapt.images.canonicalization.discrete_group import OptimizedGroupEquivariantImageCanonicalization
    from equiadapt.images.canonicalization_networks.custom_nonequivariant_networks import ConvNetwork

    torch.manual_seed(seed)
    net = ConvNetwork((in_shape[0], resize, resize), out_channels=8, kernel_size=3, num_layers=2, out_vector_size=16)
    hp = _HP(beta=1.0, input_crop_ratio=crop_ratio, resize_shape=resize, group_type=group_type,
             num_rotations=num_rotations, artifact_err_wt=0, learn_ref_vec=False)
    can = OptimizedGroupEquivariantImageCanonicalization(net, hp, in_shape).eval()
    x = smooth_images(batch, *in_shape, seed=seed + 1)
    d = {"x": x, "group_type": group_type, "num_rotations": num_rotations, "in_shape": np.array(in_shape),
         "crop_ratio": crop_ratio, "resize": resize, "reference_vector": can.reference_vector}
    with torch.no_grad():
        can.device = x.device
        xp = can.transformations_before_canonicalization_network_forward(x)
        d["x_pre"] = xp
        d["x_orbit"] = can.group_augment(xp)  # discrete_group.py:411-427
        y = can(x)
        d["vector_out"] = can.canonicalization_info_dict["vector_out"]
        d["act"] = can.canonicalization_info_dict["group_activations"]  # :475-481
        d["rotation"] = can.canonicalization_info_dict["group_element"]["rotation"]
        if group_type == "roto-reflection":
            d["reflection"] = can.canonicalization_info_dict["group_element"]["reflection"]
        d["x_canon"] = y
        d["opt_loss"] = can.get_optimization_specific_loss()  # :483-512
        d["prior_loss"] = can.get_prior_regularization_loss()
    _save(name, d)


def golden_pointcloud():
    from equiadapt.pointcloud.canonicalization.continuous_group import EquivariantPointcloudCanonicalization

    g = torch.Generator().manual_seed(11)
    x = torch.randn(6, 3, 64, generator=g)
    vecs = torch.randn(6, 3, 3, generator=g)

    class Net(torch.nn.Module):
        def forward(self, _x):
            return vecs

    can = EquivariantPointcloudCanonicalization(Net(), _HP()).eval()
    with torch.no_grad():
        y = can(x)  # pointcloud/canonicalization/continuous_group.py:51-81, :107-134
        d = {"x": x, "vectors": vecs, "x_canon": y,
             "rotation": can.canonicalization_info_dict["group_element_matrix_representation"],
             "prior_loss": can.get_prior_regularization_loss(),  # basecanonicalization.py:390-408
             "identity_metric": can.get_identity_metric()}
    _save("pointcloud_so3", d)


def golden_nbody():
    from equiadapt.nbody.canonicalization.euclidean_group import EuclideanGroupNBody

    g = torch.Generator().manual_seed(13)
    systems, particles = 7, 5
    m = systems * particles
    loc = torch.randn(m, 3, generator=g)
    vel = torch.randn(m, 3, generator=g)
    rv = torch.randn(systems, 3, 3, generator=g).repeat_interleave(particles, dim=0)
    t = torch.randn(systems, 3, generator=g).repeat_interleave(particles, dim=0)

    class Net(torch.nn.Module):
        def forward(self, *a):
            return rv, t

    can = EuclideanGroupNBody(Net()).eval()
    nodes = torch.sqrt(torch.sum(vel ** 2, dim=1)).unsqueeze(1)
    with torch.no_grad():
        # nbody/canonicalization/euclidean_group.py:87-124 (kwarg ORDER matters, :104)
        cl, cv = can(nodes, None, loc=loc, edges=None, vel=vel, edge_attr=None, charges=None)
        pred = torch.randn(m, 3, generator=g)
        inv = can.invert_canonicalization(pred)  # :126-137
        d = {"loc": loc, "vel": vel, "rot_vectors": rv, "translation": t, "canon_loc": cl, "canon_vel": cv,
             "rotation": can.canonicalization_info_dict["group_element"]["rotation_matrix"],
             "pred": pred, "inverted": inv}
    _save("nbody_e3", d)


def golden_frames_training():
    """N3: gradients of the UNMODIFIED reference's point-cloud / n-body canonicalizers in train() mode with respect to the
    frame network's OUTPUT (the (B,3,3) vectors, the per-row rotation vectors and translations): torch autograd through
    gram_schmidt + bmm + the MSE prior (pointcloud continuous_group.py:51-134, basecanonicalization.py:390-408) and through
    modified Gram-Schmidt + the row products + the inverse map (nbody euclidean_group.py:43-157)."""
    from equiadapt.nbody.canonicalization.euclidean_group import EuclideanGroupNBody
    from equiadapt.pointcloud.canonicalization.continuous_group import EquivariantPointcloudCanonicalization

    g = torch.Generator().manual_seed(101)
    x = torch.randn(6, 3, 80, generator=g)
    w = torch.randn(6, 3, 80, generator=g)
    vecs = torch.randn(6, 3, 3, generator=g).requires_grad_(True)

    class PNet(torch.nn.Module):
        def forward(self, _x):
            return vecs * 1.0

    can = EquivariantPointcloudCanonicalization(PNet(), _HP()).train()
    loss = (can(x) * w).sum() + 5.0 * can.get_prior_regularization_loss()
    loss.backward()
    _save("pointcloud_train", {"x": x, "w": w, "vectors": vecs, "loss": loss, "g_vectors": vecs.grad})

    g = torch.Generator().manual_seed(103)
    systems, particles = 6, 5
    m = systems * particles
    loc, vel = torch.randn(m, 3, generator=g), torch.randn(m, 3, generator=g)
    rv = torch.randn(systems, 3, 3, generator=g).repeat_interleave(particles, dim=0).requires_grad_(True)
    t = torch.randn(systems, 3, generator=g).repeat_interleave(particles, dim=0).requires_grad_(True)
    wl, wv, wi = (torch.randn(m, 3, generator=g) for _ in range(3))
    pred = torch.randn(m, 3, generator=g)

    class NNet(torch.nn.Module):
        def forward(self, *a):
            return rv * 1.0, t * 1.0

    can = EuclideanGroupNBody(NNet()).train()
    nodes = torch.sqrt(torch.sum(vel ** 2, dim=1)).unsqueeze(1)
    cl, cv = can(nodes, None, loc=loc, edges=None, vel=vel, edge_attr=None, charges=None)
    inv = can.invert_canonicalization(pred)
    loss = (cl * wl).sum() + (cv * wv).sum() + (inv * wi).sum()
    loss.backward()
    _save("nbody_train", {"loc": loc, "vel": vel, "rot_vectors": rv, "translation": t, "wl": wl, "wv": wv, "wi": wi,
                          "pred": pred, "loss": loss, "g_rot_vectors": rv.grad, "g_translation": t.grad})


def _randomise_bn(net, g):
    """eval-mode batch norms with non-trivial running statistics and affine parameters"""
    for mod in net.modules():
        if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            n = mod.num_features
            mod.weight.data = torch.rand(n, generator=g) * 0.8 + 0.6
            mod.bias.data = torch.randn(n, generator=g) * 0.2
            mod.running_mean.data = torch.rand(n, generator=g) * 0.5
            mod.running_var.data = torch.rand(n, generator=g) * 1.5 + 0.5


def golden_vnsmall():
    from equiadapt.pointcloud.canonicalization_networks.equivariant_networks import VNSmall

    g = torch.Generator().manual_seed(17)
    torch.manual_seed(17)
    net = VNSmall(_HP(n_knn=20, pooling="mean")).eval()   # equivariant_networks.py:79-150
    _randomise_bn(net, g)
    x = torch.randn(3, 3, 160, generator=g)
    with torch.no_grad():
        out = net(x)
    d = {"x": x, "out": out, "n_knn": 20}
    d.update({"sd." + k: v for k, v in net.state_dict().items() if "num_batches_tracked" not in k})
    _save("vnsmall", d)


def golden_vndeepsets():
    from equiadapt.nbody.canonicalization_networks.custom_equivariant_networks import VNDeepSets

    for tag, nonlin, feat, pool, trans, seed in (("relu_p", "relu", "p", "mean", False, 19),
                                                 ("softplus_pvac", "softplus", "pvac", "sum", True, 23)):
        g = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        systems = 9
        from types import SimpleNamespace
        hp = SimpleNamespace(out_dim=4, hidden_dim=16, layer_pooling=pool, final_pooling="mean", num_layers=4, nonlinearity=nonlin,
                 canon_feature=feat, canon_translation=trans, angular_feature=0, dropout=0.5, batch_size=systems)
        net = VNDeepSets(hp, device="cpu").eval()          # nbody custom_equivariant_networks.py:13-172
        m = systems * 5
        loc, vel = torch.randn(m, 3, generator=g), torch.randn(m, 3, generator=g)
        charges = (torch.randint(0, 2, (m, 1), generator=g) * 2 - 1).float()
        rows, cols = [], []                                # K5 per system, examples/nbody/model_utils.py:60-89
        for s in range(systems):
            for i in range(5):
                for j in range(5):
                    if i != j:
                        rows.append(s * 5 + i)
                        cols.append(s * 5 + j)
        edges = torch.tensor([rows, cols], dtype=torch.long)
        with torch.no_grad():
            rv, t = net(None, loc, edges, vel, None, charges)
        d = {"loc": loc, "vel": vel, "charges": charges, "edges": edges, "rot_vectors": rv, "translation": t}
        d.update({"sd." + k: v for k, v in net.state_dict().items()})
        _save("vndeepsets_" + tag, d)


def golden_continuous_images():
    from equiadapt.images.canonicalization.continuous_group import (ContinuousGroupImageCanonicalization,
                                                                    OptimizedSteerableImageCanonicalization)

    for tag, shape, with_reflection, seed in (("rot", (3, 40, 40), False, 31), ("refl", (3, 36, 36), True, 33),
                                              ("gray", (1, 28, 28), False, 35)):
        g = torch.Generator().manual_seed(seed)
        b = 6
        x = smooth_images(b, *shape, seed=seed)
        ang = torch.rand(b, generator=g) * 2 * torch.pi
        rot = torch.stack([torch.stack([torch.cos(ang), torch.sin(ang)], 1), torch.stack([-torch.sin(ang), torch.cos(ang)], 1)], 1)
        element = {"rotation": rot.clone()}
        if with_reflection:
            element["reflection"] = torch.randint(0, 2, (b, 1, 1, 1), generator=g).float()
        hp = _HP(input_crop_ratio=0.9, resize_shape=(16, 16))
        can = ContinuousGroupImageCanonicalization(torch.nn.Identity(), hp, shape)
        # the reference's own test fixture mocks get_groupelement the same way (tests/.../test_continuous_group.py:94-121)
        with mock.patch.object(can, "get_groupelement", return_value=element), torch.no_grad():
            y = can.canonicalize(x)           # images/canonicalization/continuous_group.py:162-210
        d = {"x": x, "rotation": rot, "y": y, "rotation_after": element["rotation"]}
        if with_reflection:
            d["reflection"] = element["reflection"]
        _save("image_cont_" + tag, d)
    for tag, group_type, seed in (("rot", "rotation", 37), ("refl", "roto-reflection", 39)):
        shape, b = (3, 32, 32), 5
        x = smooth_images(b, *shape, seed=seed)
        hp = _HP(input_crop_ratio=0.9, resize_shape=(16, 16), group_type=group_type)
        can = OptimizedSteerableImageCanonicalization(torch.nn.Identity(), hp, shape)
        can.device = "cpu"
        torch.manual_seed(seed)
        with torch.no_grad():
            aug, mats = can.group_augment(x)  # continuous_group.py:362-412
        torch.manual_seed(seed)               # replay the reference's draws (:375, :386-388)
        angles = torch.rand(b) * 2 * torch.pi
        d = {"x": x, "angles": angles, "aug": aug, "mats": mats}
        if group_type == "roto-reflection":
            d["reflect"] = torch.randint(0, 2, (b,)).float() * 2 - 1
        _save("image_cont_augment_" + tag, d)


def golden_training_step():
    """N3: parameter gradients of the UNMODIFIED reference in train() mode (torch autograd on CPU).
    Loss = prior regularisation (basecanonicalization.py:290-301) + a linear functional of the group activations, i.e.
    everything of the training step that is differentiable without ambiguity: the gradient through kornia's rotate at
    quarter turns is one-sided and decided by float32 noise, so the task-loss path is pinned against the
    oracle chain instead (tests/test_gpu_parity.py)."""
    from equiadapt.images.canonicalization.discrete_group import GroupEquivariantImageCanonicalization
    from equiadapt.images.canonicalization_networks.custom_equivariant_networks import CustomEquivariantNetwork

    for tag, group_type, n, layers, seed in (("c8", "rotation", 8, 3, 91), ("d4", "roto-reflection", 4, 2, 93)):
        torch.manual_seed(seed)
        net = CustomEquivariantNetwork((3, 24, 24), 6, 5, group_type, n, layers, device="cpu")
        net.group_type, net.num_rotations = group_type, n       # quirk A.4-1
        with torch.no_grad():
            for m in net.eqv_network:
                if hasattr(m, "bias") and m.bias is not None:
                    m.bias.uniform_(-0.05, 0.05)
        hp = _HP(beta=1.0, input_crop_ratio=0.9, resize_shape=24)
        can = GroupEquivariantImageCanonicalization(net, hp, (3, 32, 32)).train()
        x = smooth_images(6, 3, 32, 32, seed=seed + 1)
        G = can.num_group
        wact = torch.randn(6, G, generator=torch.Generator().manual_seed(seed + 2))
        can(x)                                                   # canonicalize: fills canonicalization_info_dict
        act = can.canonicalization_info_dict["group_activations"]
        prior = can.get_prior_regularization_loss()
        loss = 100.0 * prior + (act * wact).sum()
        loss.backward()
        d = {"x": x, "wact": wact, "group_type": group_type, "num_rotations": n, "num_layers": layers, "act": act,
             "prior": prior, "loss": loss, "in_shape": np.array((3, 32, 32)), "crop_ratio": 0.9, "resize": 24, "beta": 1.0}
        for i, m in enumerate(net.eqv_network):
            if hasattr(m, "weights"):
                d[f"w{i}"], d[f"b{i}"] = m.weights, m.bias
                d[f"gw{i}"], d[f"gb{i}"] = m.weights.grad, m.bias.grad
        _save("train_step_" + tag, d)


def golden_optimized_training():
    """N3: the optimisation-based variant of the UNMODIFIED reference in train() mode: gradients of
    loss = 10 * prior + 3 * optimisation-specific loss with respect to the consumer network's OUTPUT (the (|G| B, V) vectors,
    so the fixture does not depend on that network's architecture) and to the learnable reference vector
    (discrete_group.py:429-512, basecanonicalization.py:290-301)."""
    from equiadapt.images.canonicalization.discrete_group import OptimizedGroupEquivariantImageCanonicalization
    from equiadapt.images.canonicalization_networks.custom_nonequivariant_networks import ConvNetwork

    for tag, group_type, n, seed in (("c8", "rotation", 8, 95), ("d4", "roto-reflection", 4, 97)):
        torch.manual_seed(seed)
        net = ConvNetwork((3, 24, 24), out_channels=8, kernel_size=3, num_layers=2, out_vector_size=16)
        hp = _HP(beta=1.0, input_crop_ratio=0.9, resize_shape=24, group_type=group_type, num_rotations=n,
                 artifact_err_wt=0, learn_ref_vec=True)
        can = OptimizedGroupEquivariantImageCanonicalization(net, hp, (3, 32, 32)).train()
        x = smooth_images(5, 3, 32, 32, seed=seed + 1)
        can(x)
        vec = can.canonicalization_info_dict["vector_out"]
        vec.retain_grad()
        prior, opt = can.get_prior_regularization_loss(), can.get_optimization_specific_loss()
        loss = 10.0 * prior + 3.0 * opt
        loss.backward()
        _save("opt_train_step_" + tag, {
            "x": x, "group_type": group_type, "num_rotations": n, "in_shape": np.array((3, 32, 32)), "crop_ratio": 0.9,
            "resize": 24, "vector_out": vec, "reference_vector": can.reference_vector, "prior": prior, "opt_loss": opt,
            "loss": loss, "act": can.canonicalization_info_dict["group_activations"],
            "g_vector_out": vec.grad, "g_reference_vector": can.reference_vector.grad})


def golden_group_inference():
    """The evaluation orbit of the UNMODIFIED reference: GroupInference.get_group_element_wise_logits
    (examples/images/classification/inference_utils.py:97-122) called with an identity canonicalizer and an identity
    prediction network, so the "logits" of element g ARE the padded / mirrored / rotated / cropped batch.  The example
    module imports through the omegaconf stand-in of oracle/shims; Pad / hflip / rotate / CenterCrop are torchvision's."""
    sys.path.insert(0, os.path.join(REF, "examples", "images", "classification"))
    import inference_utils  # the reference's example module

    for tag, n, reflect, shape, seed in (("c4", 4, False, (2, 3, 32, 32), 41), ("d8", 8, True, (2, 3, 32, 32), 43),
                                         ("c6_gray", 6, False, (3, 1, 28, 28), 45), ("d5_rect", 5, True, (1, 3, 30, 37), 47)):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(*shape, generator=g)
        hp = _HP(method="group", group_type="roto-reflection" if reflect else "rotation", num_rotations=n)
        inf = inference_utils.get_inference_method(torch.nn.Identity(), torch.nn.Identity(), 10, hp, tuple(shape[1:]))
        with torch.no_grad():
            logits = inf.get_group_element_wise_logits(x)
        members = [logits[k] for k in sorted(logits)]
        _save("group_inference_orbit_" + tag, {"x": x, "orbit": torch.stack(members), "num_rotations": n,
                                               "reflect": int(reflect)})


def golden_inference_metrics():
    """Metric dictionaries of the UNMODIFIED reference's VanillaInference / GroupInference.get_inference_metrics
    (inference_utils.py:50-77, :124-168) for a fixed linear prediction network and an identity canonicalizer; the
    per-element logits are stored too, so the host-side metric assembly can be checked without a GPU."""
    sys.path.insert(0, os.path.join(REF, "examples", "images", "classification"))
    import inference_utils

    torch.manual_seed(99)
    pred = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 16 * 16, 5))
    x = torch.randn(24, 3, 16, 16)
    y = torch.randint(0, 4, (24,))            # class 4 never occurs: exercises the nan -> 0 branch (:63-66)
    d = {"x": x, "y": y, "weight": pred[1].weight, "bias": pred[1].bias}
    with torch.no_grad():
        for name, hp in (("vanilla", _HP(method="vanilla")),
                         ("group", _HP(method="group", group_type="roto-reflection", num_rotations=4))):
            inf = inference_utils.get_inference_method(torch.nn.Identity(), pred, 5, hp, (3, 16, 16))
            m = inf.get_inference_metrics(x, y)
            d[name + "_keys"] = np.array(sorted(m))
            d[name + "_values"] = torch.stack([torch.as_tensor(float(m[k])) for k in sorted(m)])
            if name == "group":
                logits = inf.get_group_element_wise_logits(x)
                d["group_logits"] = torch.stack([logits[k] for k in sorted(logits)])
    _save("inference_metrics", d)


def main():
    _import_reference()
    if "--only-orbit" in sys.argv:
        golden_group_inference()
        golden_inference_metrics()
        return
    if "--only-train" in sys.argv:
        golden_training_step()
        golden_optimized_training()
        golden_frames_training()
        return
    if "--only-cont" in sys.argv:
        golden_continuous_images()
        return
    if "--only-vn" in sys.argv:
        golden_vnsmall()
        golden_vndeepsets()
        return
    golden_gram_schmidt()
    # cfg1 of BASELINE.json (C4, 3x32x32, oc16/k5/L3, crop .9 -> 29 (offset 2), resize 32), batch cut to 4
    _image_case("image_c4_cfg1", "rotation", 4, (3, 32, 32), 4, 16, 5, 3, 0.9, 32, 1.0, seed=0)
    # cfg2 in miniature: C8, crop .8, down-sampling antialiased resize, k5, L3
    _image_case("image_c8_small", "rotation", 8, (3, 40, 40), 8, 4, 5, 3, 0.8, 20, 1.0, seed=10)
    # D4 (roto-reflection), 2 regular fields in the inverted feature map
    _image_case("image_d4_small", "roto-reflection", 4, (3, 36, 36), 8, 4, 3, 2, 0.9, 24, 0.5, seed=20,
                regular_fields=2, fshape=(12, 12))
    # D8 exercises reflected 45-degree elements; single layer (lift only)
    _image_case("image_d8_small", "roto-reflection", 8, (3, 24, 24), 16, 3, 3, 1, 1.0, 16, 1.0, seed=30, fshape=(10, 10))
    # grayscale: no pad / crop / resize, zero-fill rotate (discrete_group.py:60-71)
    _image_case("image_c4_gray", "rotation", 4, (1, 28, 28), 4, 4, 3, 2, 0.9, 28, 1.0, seed=40)
    # non-square input
    _image_case("image_c8_rect", "rotation", 8, (3, 24, 32), 8, 2, 3, 2, 1.0, (16, 16), 1.0, seed=50, fshape=(10, 14))
    golden_optimized("image_opt_d4", "roto-reflection", 4, (3, 40, 40), 4, 24, 0.8, seed=60)
    golden_optimized("image_opt_c8", "rotation", 8, (3, 32, 32), 3, 16, 0.9, seed=70)
    golden_pointcloud()
    golden_nbody()
    golden_vnsmall()
    golden_vndeepsets()
    golden_continuous_images()
    golden_group_inference()
    golden_inference_metrics()
    golden_training_step()
    golden_optimized_training()
    golden_frames_training()


if __name__ == "__main__":
    main()
