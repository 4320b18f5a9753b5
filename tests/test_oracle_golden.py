"""CPU: pin the oracle (oracle/reference_path.py) against the reference's own outputs.

The golden files were produced by running the unmodified reference (oracle/make_golden.py);
the one numeric known-answer test of the reference (tests/common/test_utils.py:6-12) is here too.
"""
import pytest
import torch

from conftest import IMAGE_CASES, golden_layers, load_golden, rel_err, resize_arg
from oracle import reference_path as O

TOL = 1e-6  # same ops in the same order: differences are thread-count dependent summation order only


def test_gram_schmidt_known_answer():
    # /root/reference/tests/common/test_utils.py:6-12
    torch.manual_seed(0)
    out = O.gram_schmidt(torch.randn(1, 3, 3))
    assert torch.allclose(out[0][0][0], torch.tensor(0.5740), atol=1e-4)


def test_gram_schmidt_golden():
    g = load_golden("gram_schmidt")
    assert torch.allclose(g["kat_out"][0, 0, 0], torch.tensor(0.5740), atol=1e-4)
    assert torch.equal(O.gram_schmidt(g["kat_in"]), g["kat_out"])
    assert rel_err(O.gram_schmidt(g["batch_in"]), g["batch_out"]) < TOL


@pytest.mark.parametrize("case", IMAGE_CASES)
def test_image_path_matches_reference(case):
    g = load_golden(case)
    reflect = g["group_type"] == "roto-reflection"
    n = g["num_rotations"]
    num_group = n * (2 if reflect else 1)
    in_shape = tuple(int(v) for v in g["in_shape"])
    layers = golden_layers(g)

    x_pre = O.pre_network_transform(g["x"], in_shape, g["crop_ratio"], resize_arg(g))
    assert rel_err(x_pre, g["x_pre"]) < TOL
    assert rel_err(O.lift_filter_orbit(layers[0][0], n, reflect), g["orbit_lift"]) < TOL
    if len(layers) > 1:
        assert rel_err(O.regular_filter_orbit(layers[1][0], n, reflect), g["orbit_reg"]) < TOL
    act = O.custom_equivariant_network(x_pre, layers, n, reflect)
    assert rel_err(act, g["act"]) < 1e-5
    el = O.activations_to_group_element(g["act"], n, reflect, g["beta"])
    assert torch.equal(el["rotation"], g["rotation"])
    if reflect:
        assert torch.equal(el["reflection"], g["reflection"])
    y = O.canonicalize_image(g["x"], g["rotation"], g.get("reflection"))
    assert rel_err(y, g["x_canon"]) < TOL
    assert abs(float(O.prior_loss_discrete(g["act"]) - g["prior_loss"])) < 1e-6
    assert float(O.identity_metric_discrete(g["act"])) == float(g["identity_metric"])
    for rep in ("regular", "scalar"):
        inv = O.invert_image_features(g[f"f_{rep}"], g["rotation"], g.get("reflection"), n, num_group, rep)
        assert rel_err(inv, g[f"inv_{rep}"]) < TOL

    # every group element (forced)
    idx = g["forced_idx"]
    ang = torch.linspace(0.0, 360.0, n + 1)[:n][idx % n]
    refl = (idx >= n).float() if reflect else None
    assert rel_err(O.canonicalize_image(g["x"], ang, refl), g["forced_canon"]) < TOL
    for rep in ("regular", "scalar"):
        inv = O.invert_image_features(g[f"f_{rep}"], ang, refl, n, num_group, rep)
        assert rel_err(inv, g[f"forced_inv_{rep}"]) < TOL


@pytest.mark.parametrize("case", ["image_opt_d4", "image_opt_c8"])
def test_optimized_path_matches_reference(case):
    g = load_golden(case)
    reflect = g["group_type"] == "roto-reflection"
    n = g["num_rotations"]
    num_group = n * (2 if reflect else 1)
    in_shape = tuple(int(v) for v in g["in_shape"])
    x_pre = O.pre_network_transform(g["x"], in_shape, g["crop_ratio"], g["resize"])
    assert rel_err(x_pre, g["x_pre"]) < TOL
    assert rel_err(O.group_augment(x_pre, n, reflect, g["resize"]), g["x_orbit"]) < TOL
    act = O.cosine_group_activations(g["vector_out"], g["reference_vector"], num_group)
    assert rel_err(act, g["act"]) < TOL
    el = O.activations_to_group_element(g["act"], n, reflect)
    assert torch.equal(el["rotation"], g["rotation"])
    assert rel_err(O.canonicalize_image(g["x"], g["rotation"], g.get("reflection")), g["x_canon"]) < TOL
    v = g["vector_out"]
    assert abs(float(O.optimization_specific_loss(v, num_group, v.shape[1]) - g["opt_loss"])) < 1e-6
    assert abs(float(O.prior_loss_discrete(g["act"]) - g["prior_loss"])) < 1e-6


def test_pointcloud_matches_reference():
    g = load_golden("pointcloud_so3")
    r = O.gram_schmidt(g["vectors"])
    assert rel_err(r, g["rotation"]) < TOL
    assert rel_err(O.so3_canonicalize(g["x"], r), g["x_canon"]) < TOL
    assert abs(float(O.prior_loss_continuous(r) - g["prior_loss"])) < 1e-6
    assert abs(float(O.identity_metric_continuous(r) - g["identity_metric"])) < 1e-6


def test_nbody_matches_reference():
    g = load_golden("nbody_e3")
    r = O.modified_gram_schmidt(g["rot_vectors"])
    assert rel_err(r, g["rotation"]) < TOL
    cl, cv = O.e3_canonicalize(g["loc"], g["vel"], r, g["translation"])
    assert rel_err(cl, g["canon_loc"]) < TOL
    assert rel_err(cv, g["canon_vel"]) < TOL
    assert rel_err(O.e3_invert(g["pred"], r, g["translation"]), g["inverted"]) < TOL


def test_aa_resize_restatement_matches_torch():
    """The closed-form separable filter the CUDA kernel implements == ATen _upsample_bilinear2d_aa."""
    g = torch.Generator().manual_seed(3)
    for (h, w, oh, ow) in [(180, 180, 96, 96), (29, 29, 32, 32), (32, 32, 20, 20), (24, 32, 16, 16), (17, 40, 33, 9)]:
        x = torch.rand(2, 3, h, w, generator=g)
        ref = torch.nn.functional.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False, antialias=True)
        assert rel_err(O.aa_resize_restated(x, oh, ow), ref) < 2e-6


def test_rotate_closed_form_bounds_reference_noise():
    """kornia-style fp32 rotate vs the exact map (SURVEY.md section 7 hard part 2)."""
    from oracle import kornia_restated as K

    g = torch.Generator().manual_seed(5)
    x = torch.rand(8, 3, 32, 32, generator=g)
    ang = torch.linspace(0.0, 360.0, 9)[:8]
    ref = K.rotate(x, ang)
    exact = O.rotate_closed_form(x, ang, clamp=False)
    assert rel_err(ref, exact) < 1e-4
    assert torch.equal(exact[2].float(), torch.rot90(x[2], 1, (1, 2)))  # +90 deg == rot90(k=1)
    y = O.canonicalize_image(x, ang, None)
    exact_c = O.rotate_closed_form(x, -ang, clamp=True)
    assert rel_err(y, exact_c) < 1e-4


def test_expanded_conv_restatement_agrees_with_pinned_custom_network():
    """a7's oracle (dense ops on expanded filters) reproduces the golden-pinned a6 oracle when it is handed the
    custom network's own filter orbits: ties the un-pinnable e2cnn path to arithmetic that IS pinned."""
    torch.manual_seed(5)
    n, cout, k = 4, 6, 3
    w0, b0 = torch.randn(cout, 3, k, k) * 0.2, torch.randn(cout) * 0.1
    w1, b1 = torch.randn(cout, cout, n, 1, 1) * 0.2, torch.randn(cout) * 0.1
    x = torch.rand(3, 3, 20, 20)
    ref = O.custom_equivariant_network(x, [(w0, b0), (w1, b1)], n, False)
    filt = [O.lift_filter_orbit(w0, n, False), O.regular_filter_orbit(w1, n, False)]
    bias = [b0.repeat_interleave(n), b1.repeat_interleave(n)]
    got = O.expanded_conv_network(x, filt, bias, [None, None], [None, None], n)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-7)


def _sd(g):
    return {k[3:]: v for k, v in g.items() if k.startswith("sd.")}


def test_vnsmall_matches_reference():
    """N1: oracle restatement of VNSmall vs the unmodified reference module (oracle/make_golden.py::golden_vnsmall)."""
    g = load_golden("vnsmall")
    out = O.vnsmall_forward(g["x"], _sd(g), int(g["n_knn"]))
    assert out.shape == (3, 3, 3)
    assert rel_err(out, g["out"]) < 1e-5


@pytest.mark.parametrize("tag,nonlin,feat,pool,trans", [("relu_p", "relu", "p", "mean", False),
                                                        ("softplus_pvac", "softplus", "pvac", "sum", True)])
def test_vndeepsets_matches_reference(tag, nonlin, feat, pool, trans):
    g = load_golden("vndeepsets_" + tag)
    rv, t = O.vndeepsets_forward(g["loc"], g["vel"], g["charges"], g["edges"].long(), _sd(g), 4, nonlin, feat, pool, "mean", trans)
    assert rel_err(rv, g["rot_vectors"]) < 1e-5
    assert rel_err(t, g["translation"]) < 1e-5


@pytest.mark.parametrize("tag", ["rot", "refl", "gray"])
def test_continuous_canonicalize_matches_reference(tag):
    """N2: restated canonicalize (continuous_group.py:162-210) vs the unmodified reference class."""
    g = load_golden("image_cont_" + tag)
    y = O.canonicalize_image_continuous(g["x"], g["rotation"], g.get("reflection"))
    assert rel_err(y, g["y"]) < TOL
    neg = g["rotation"].clone()
    neg[:, [0, 1], [1, 0]] *= -1
    assert torch.equal(neg, g["rotation_after"])      # the reference's in-place sign flip


@pytest.mark.parametrize("tag", ["rot", "refl"])
def test_continuous_group_augment_matches_reference(tag):
    g = load_golden("image_cont_augment_" + tag)
    aug, mats = O.group_augment_continuous(g["x"], g["angles"], g.get("reflect"))
    assert rel_err(aug, g["aug"]) < TOL and rel_err(mats, g["mats"]) < TOL


ORBIT_CASES = ["c4", "d8", "c6_gray", "d5_rect"]


@pytest.mark.parametrize("tag", ORBIT_CASES)
def test_group_inference_orbit_matches_torchvision(tag):
    """N4: restated evaluation orbit vs the UNMODIFIED reference's GroupInference.get_group_element_wise_logits
    (examples/images/classification/inference_utils.py:97-122; identity canonicalizer and prediction network, so the
    per-element logits are the torchvision Pad / hflip / rotate(NEAREST) / CenterCrop outputs): bit-exact, ties included."""
    g = load_golden("group_inference_orbit_" + tag)
    orbit, margin = O.group_inference_orbit(g["x"], int(g["num_rotations"]), bool(int(g["reflect"])), return_margin=True)
    assert orbit.shape == g["orbit"].shape and margin.shape == orbit.shape[:1] + orbit.shape[-2:]
    assert torch.equal(orbit, g["orbit"])
    assert torch.equal(orbit[0], g["x"])                  # element 0 is the identity


def test_linspace_degrees_restatement():
    """torch.linspace's scalar formula; when 360/n is not a float32 the vectorised CPU kernel (lane base + lane*step,
    so the value depends on the SIMD width of the machine) may differ from it in the last ulp."""
    for n in (1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 16, 64):
        assert O.linspace_degrees(n) == [d.item() for d in torch.linspace(0, 360, n + 1)[:-1]]
    for n in (7, 11, 13, 31):
        for a, b in zip(O.linspace_degrees(n), torch.linspace(0, 360, n + 1)[:-1]):
            assert abs(a - b.item()) <= 3.1e-5


@pytest.mark.parametrize("tag", ["c8", "d4"])
def test_training_gradients_of_the_oracle_match_the_unmodified_reference(tag):
    """N3 pin: torch autograd through the oracle's restated network (filter orbits -> conv2d -> ReLU -> mean) gives the
    parameter gradients the UNMODIFIED reference produced in train() mode for loss = 100 * prior + <act, w>
    (tests/golden/train_step_*.npz, written by oracle/make_golden.py::golden_training_step)."""
    from conftest import golden_layers
    g = load_golden("train_step_" + tag)
    reflect = g["group_type"] == "roto-reflection"
    n = g["num_rotations"]
    layers = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in golden_layers(g)]
    x = O.pre_network_transform(g["x"], (3, 32, 32), g["crop_ratio"], int(g["resize"]))
    act = O.custom_equivariant_network(x, layers, n, reflect)
    assert rel_err(act.detach(), g["act"]) < TOL
    loss = 100.0 * O.prior_loss_discrete(act) + (act * g["wact"]).sum()
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    loss.backward()
    for i, (w, b) in enumerate(layers):
        assert rel_err(w.grad, g[f"gw{2 * i}"]) < 1e-4
        assert rel_err(b.grad, g[f"gb{2 * i}"]) < 1e-4


@pytest.mark.parametrize("tag", ["c8", "d4"])
def test_optimized_variant_training_gradients_match_the_unmodified_reference(tag):
    """N3 pin: gradients of 10 * prior + 3 * optimisation-specific loss w.r.t. the consumer network's output vectors and the
    reference vector, oracle chain vs the UNMODIFIED reference class in train() (tests/golden/opt_train_step_*.npz)."""
    g = load_golden("opt_train_step_" + tag)
    G = g["num_rotations"] * (2 if g["group_type"] == "roto-reflection" else 1)
    vec = g["vector_out"].clone().requires_grad_(True)
    rv = g["reference_vector"].clone().requires_grad_(True)
    act = O.cosine_group_activations(vec, rv, G)
    assert rel_err(act.detach(), g["act"]) < TOL
    loss = 10.0 * O.prior_loss_discrete(act) + 3.0 * O.optimization_specific_loss(vec, G, vec.shape[1])
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    assert rel_err(vec.grad, g["g_vector_out"]) < 1e-5 and rel_err(rv.grad, g["g_reference_vector"]) < 1e-5


def test_frame_path_training_gradients_match_the_unmodified_reference():
    """N3 pin: torch autograd through the oracle's frame functions equals the gradients the UNMODIFIED reference classes
    produced in train() mode w.r.t. the frame network's outputs (tests/golden/pointcloud_train.npz, nbody_train.npz)."""
    g = load_golden("pointcloud_train")
    v = g["vectors"].clone().requires_grad_(True)
    R = O.gram_schmidt(v)
    loss = (O.so3_canonicalize(g["x"], R) * g["w"]).sum() + 5.0 * O.prior_loss_continuous(R)
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    assert rel_err(v.grad, g["g_vectors"]) < 1e-5
    g = load_golden("nbody_train")
    rv, t = g["rot_vectors"].clone().requires_grad_(True), g["translation"].clone().requires_grad_(True)
    R = O.modified_gram_schmidt(rv)
    cl, cv = O.e3_canonicalize(g["loc"], g["vel"], R, t)
    loss = (cl * g["wl"]).sum() + (cv * g["wv"]).sum() + (O.e3_invert(g["pred"], R, t) * g["wi"]).sum()
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    assert rel_err(rv.grad, g["g_rot_vectors"]) < 1e-5 and rel_err(t.grad, g["g_translation"]) < 1e-5
