"""GPU parity: the CUDA path (through the C ABI / wrapper classes) vs golden vectors of the unmodified
reference and vs the oracle on seeded inputs.

Tolerance (BASELINE.json north_star): 1e-4 relative fp32, measured as ||ours-ref||_inf / ||ref||_inf;
bit-exact for the discrete group-element index wherever the fp64 oracle margin exceeds the combined
fp32 noise of both implementations (hard part 1).
"""
from types import SimpleNamespace

import pytest
import torch

from conftest import IMAGE_CASES, golden_layers, load_golden, rel_err, resize_arg
from oracle import reference_path as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _mods():
    from equiadapt_b200 import ops
    from equiadapt_b200.images.canonicalization.discrete_group import (
        GroupEquivariantImageCanonicalization, OptimizedGroupEquivariantImageCanonicalization)
    from equiadapt_b200.images.canonicalization_networks.custom_equivariant_networks import CustomEquivariantNetwork
    return ops, GroupEquivariantImageCanonicalization, OptimizedGroupEquivariantImageCanonicalization, CustomEquivariantNetwork


def build_canonicalizer(g, device):
    _, GEIC, _, Net = _mods()
    layers = golden_layers(g)
    in_shape = tuple(int(v) for v in g["in_shape"])
    cout, cin, k, _ = layers[0][0].shape
    r = resize_arg(g)
    rr = r if isinstance(r, int) else r[0]
    net = Net((cin, rr, rr), cout, k, g["group_type"], g["num_rotations"], len(layers), device=str(device))
    mods = [m for m in net.eqv_network if hasattr(m, "weights")]
    with torch.no_grad():
        for m, (w, b) in zip(mods, layers):
            m.weights.copy_(w.to(device))
            m.bias.copy_(b.to(device))
    hp = SimpleNamespace(beta=g["beta"], input_crop_ratio=g["crop_ratio"], resize_shape=r)
    return GEIC(net, hp, in_shape).eval()


def assert_index_parity(act_ours, act_ref32, act_64, idx_ours):
    """Exact index wherever the fp64 margin clears 4x the combined fp32 noise; report the rest."""
    top2 = torch.topk(act_64, 2, dim=-1).values
    margin = (top2[:, 0] - top2[:, 1])
    noise = (act_ours.double() - act_64).abs().max() + (act_ref32.double() - act_64).abs().max()
    decided = margin > 4 * noise
    want = act_64.argmax(-1)
    assert torch.equal(idx_ours.long()[decided], want[decided]), "group index differs outside the tie band"
    return int(decided.sum()), int((~decided).sum())


@pytest.mark.parametrize("case", IMAGE_CASES)
def test_image_path_vs_reference_golden(case, cuda_device):
    ops = _mods()[0]
    g = load_golden(case)
    dev = cuda_device
    reflect = g["group_type"] == "roto-reflection"
    n = g["num_rotations"]
    num_group = n * (2 if reflect else 1)
    can = build_canonicalizer(g, dev)
    x = g["x"].to(dev)

    with torch.no_grad():
        # a3
        x_pre = can.transformations_before_canonicalization_network_forward(x)
        assert rel_err(x_pre.cpu(), g["x_pre"]) < RTOL
        # a4 / a5 filter orbits
        mods = [m for m in can.canonicalization_network.eqv_network if hasattr(m, "weights")]
        assert rel_err(mods[0].filter_orbit().cpu(), g["orbit_lift"]) < RTOL
        if len(mods) > 1:
            assert rel_err(mods[1].filter_orbit().cpu(), g["orbit_reg"]) < RTOL
        # a6 fused stack on the REFERENCE's pre-transformed input and on ours
        act_on_ref_pre = can.canonicalization_network(g["x_pre"].to(dev))
        assert rel_err(act_on_ref_pre.cpu(), g["act"]) < RTOL
        # full forward
        y = can(x)
        info = can.canonicalization_info_dict
        act = info["group_activations"].cpu()
        assert rel_err(act, g["act"]) < RTOL
        act64 = O.custom_equivariant_network(
            O.pre_network_transform(g["x"].double(), tuple(int(v) for v in g["in_shape"]), g["crop_ratio"], resize_arg(g)),
            [(w.double(), b.double()) for w, b in golden_layers(g)], n, reflect)
        idx = info["group_element"].index.cpu()
        decided, in_band = assert_index_parity(act, g["act"], act64, idx)
        assert decided >= 1
        ref_idx = torch.round(g["rotation"] / 360.0 * n).long() % n
        if reflect:
            ref_idx = ref_idx + n * g["reflection"].long()
        # samples whose index equals the reference's are compared element-wise; a differing index is only legal inside
        # the tie band (and never skips the comparison of the others)
        same = idx.long() == ref_idx
        assert int((~same).sum()) <= in_band, "group index differs from the reference outside the tie band"
        assert int(same.sum()) >= 1
        assert torch.equal(info["group_element"]["rotation"].cpu()[same], g["rotation"][same])
        assert rel_err(y.cpu()[same], g["x_canon"][same]) < RTOL
        if reflect:
            assert torch.equal(info["group_element"]["reflection"].cpu()[same], g["reflection"][same])
        for rep in ("regular", "scalar"):
            inv = can.invert_canonicalization(g[f"f_{rep}"].to(dev), induced_rep_type=rep)
            assert rel_err(inv.cpu()[same], g[f"inv_{rep}"][same]) < RTOL
        # a13
        assert abs(float(can.get_prior_regularization_loss()) - float(g["prior_loss"])) < 1e-5
        assert float(can.get_identity_metric()) == pytest.approx(float((act.argmax(-1) == 0).float().mean()))

        # every group element, forced through the C-ABI warps (a10, a11)
        fidx = g["forced_idx"].to(dev).to(torch.int32)
        yc = ops.warp_canonicalize(x, fidx, n, reflect)
        assert rel_err(yc.cpu(), g["forced_canon"]) < RTOL
        for rep in ("regular", "scalar"):
            inv = ops.warp_invert(g[f"f_{rep}"].to(dev), fidx, n, reflect, rep == "regular")
            assert rel_err(inv.cpu(), g[f"forced_inv_{rep}"]) < RTOL


@pytest.mark.parametrize("case", ["image_opt_d4", "image_opt_c8"])
def test_optimized_path_vs_reference_golden(case, cuda_device):
    ops, _, OGEIC, _ = _mods()
    g = load_golden(case)
    dev = cuda_device
    reflect = g["group_type"] == "roto-reflection"
    n = g["num_rotations"]
    num_group = n * (2 if reflect else 1)
    vec = g["vector_out"].to(dev)

    class Net(torch.nn.Module):  # the consumer CNN is the caller's torch module: replay the reference's output
        out_vector_size = vec.shape[1]

        def forward(self, xa):
            self.seen = xa
            return vec

    hp = SimpleNamespace(beta=1.0, input_crop_ratio=g["crop_ratio"], resize_shape=g["resize"],
                         group_type=g["group_type"], num_rotations=n, artifact_err_wt=0, learn_ref_vec=False)
    net = Net()
    can = OGEIC(net, hp, tuple(int(v) for v in g["in_shape"])).to(dev).eval()
    with torch.no_grad():
        can.reference_vector.copy_(g["reference_vector"].to(dev))
        y = can(g["x"].to(dev))
    assert rel_err(net.seen.cpu(), g["x_orbit"]) < RTOL           # a3 + a12 orbit expand
    act = can.canonicalization_info_dict["group_activations"].cpu()
    assert rel_err(act, g["act"]) < RTOL                            # cosine activations
    assert torch.equal(can.canonicalization_info_dict["group_element"]["rotation"].cpu(), g["rotation"])
    assert rel_err(y.cpu(), g["x_canon"]) < RTOL
    assert abs(float(can.get_optimization_specific_loss()) - float(g["opt_loss"])) < 1e-5
    assert abs(float(can.get_prior_regularization_loss()) - float(g["prior_loss"])) < 1e-5


def test_pointcloud_vs_reference_golden(cuda_device):
    from equiadapt_b200.common.utils import gram_schmidt
    from equiadapt_b200.pointcloud.canonicalization.continuous_group import EquivariantPointcloudCanonicalization

    g = load_golden("pointcloud_so3")
    gs = load_golden("gram_schmidt")
    dev = cuda_device
    out = gram_schmidt(gs["kat_in"].to(dev)).cpu()
    assert torch.allclose(out[0][0][0], torch.tensor(0.5740), atol=1e-4)  # the reference's own KAT
    assert rel_err(out, gs["kat_out"]) < 1e-6
    assert rel_err(gram_schmidt(gs["batch_in"].to(dev)).cpu(), gs["batch_out"]) < 1e-5
    vecs = g["vectors"].to(dev)

    class Net(torch.nn.Module):
        def forward(self, _x):
            return vecs

    can = EquivariantPointcloudCanonicalization(Net(), SimpleNamespace()).eval()
    y = can(g["x"].to(dev))
    assert rel_err(can.canonicalization_info_dict["group_element_matrix_representation"].cpu(), g["rotation"]) < 1e-5
    assert rel_err(y.cpu(), g["x_canon"]) < 1e-5
    assert abs(float(can.get_prior_regularization_loss()) - float(g["prior_loss"])) < 1e-5
    assert abs(float(can.get_identity_metric()) - float(g["identity_metric"])) < 1e-5


def test_nbody_vs_reference_golden(cuda_device):
    from equiadapt_b200.nbody.canonicalization.euclidean_group import EuclideanGroupNBody

    g = load_golden("nbody_e3")
    dev = cuda_device
    rv, t = g["rot_vectors"].to(dev), g["translation"].to(dev)

    class Net(torch.nn.Module):
        def forward(self, *a):
            return rv, t

    can = EuclideanGroupNBody(Net()).eval()
    loc, vel = g["loc"].to(dev), g["vel"].to(dev)
    nodes = torch.sqrt(torch.sum(vel ** 2, dim=1)).unsqueeze(1)
    cl, cv = can(nodes, None, loc=loc, edges=None, vel=vel, edge_attr=None, charges=None)
    assert rel_err(can.canonicalization_info_dict["group_element"]["rotation_matrix"].cpu(), g["rotation"]) < 1e-5
    assert rel_err(cl.cpu(), g["canon_loc"]) < 1e-5
    assert rel_err(cv.cpu(), g["canon_vel"]) < 1e-5
    assert rel_err(can.invert_canonicalization(g["pred"].to(dev)).cpu(), g["inverted"]) < 1e-5
    # extension: the prior loss works for n-body (the reference raises KeyError), oracle = a14 restated
    r = O.modified_gram_schmidt(g["rot_vectors"])
    assert abs(float(can.get_prior_regularization_loss()) - float(O.prior_loss_continuous(r))) < 1e-5


# ---------------------------------------------------------------------------------------------------
# seeded inputs vs the oracle, sizes the oracle finishes in seconds
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("group_type,n,shape,cout,k,layers,crop,resize", [
    ("rotation", 8, (3, 64, 64), 8, 5, 3, 0.8, 32),
    ("rotation", 4, (3, 33, 47), 5, 3, 2, 0.9, (21, 21)),
    ("roto-reflection", 4, (3, 48, 48), 6, 5, 3, 0.8, 24),
    ("rotation", 8, (3, 224, 224), 32, 5, 3, 0.8, 96),       # cfg2 network, small batch
    ("rotation", 6, (3, 40, 40), 4, 3, 2, 1.0, 20),           # N not a power of two
])
def test_full_pipeline_vs_oracle(group_type, n, shape, cout, k, layers, crop, resize, cuda_device):
    ops, GEIC, _, Net = _mods()
    dev = cuda_device
    reflect = group_type == "roto-reflection"
    num_group = n * (2 if reflect else 1)
    torch.manual_seed(0)
    rr = resize if isinstance(resize, int) else resize[0]
    net = Net((shape[0], rr, rr), cout, k, group_type, n, layers, device="cpu")
    with torch.no_grad():
        for m in net.eqv_network:
            if hasattr(m, "bias"):
                m.bias.uniform_(-0.05, 0.05)
    lay = [(m.weights.detach().clone(), m.bias.detach().clone()) for m in net.eqv_network if hasattr(m, "weights")]
    net = net.to(dev)
    can = GEIC(net, SimpleNamespace(beta=1.0, input_crop_ratio=crop, resize_shape=resize), shape).eval()
    g = torch.Generator().manual_seed(1)
    b = 4 if shape[-1] >= 224 else 12
    x = torch.rand(b, *shape, generator=g)
    with torch.no_grad():
        y = can(x.to(dev))
    info = can.canonicalization_info_dict
    x_pre = O.pre_network_transform(x, shape, crop, resize)
    act32 = O.custom_equivariant_network(x_pre, lay, n, reflect)
    act64 = O.custom_equivariant_network(
        O.pre_network_transform(x.double(), shape, crop, resize), [(w.double(), bb.double()) for w, bb in lay], n, reflect)
    act = info["group_activations"].cpu()
    assert rel_err(act, act32) < RTOL
    idx = info["group_element"].index.cpu()
    assert_index_parity(act, act32, act64, idx)
    # warp parity given OUR index (so near-ties cannot mask a warp bug)
    ang = torch.linspace(0.0, 360.0, n + 1)[:n][idx.long() % n]
    refl = (idx >= n).float() if reflect else None
    assert rel_err(y.cpu(), O.canonicalize_image(x, ang, refl)) < RTOL
    for rep, ch in (("regular", 2 * num_group), ("scalar", 3)):
        f = torch.randn(b, ch, 30, 26, generator=g)
        inv = can.invert_canonicalization(f.to(dev), induced_rep_type=rep)
        assert rel_err(inv.cpu(), O.invert_image_features(f, ang, refl, n, num_group, rep)) < RTOL
    assert abs(float(can.get_prior_regularization_loss()) - float(O.prior_loss_discrete(act))) < 1e-5
    assert float(can.get_identity_metric()) == pytest.approx(float(O.identity_metric_discrete(act)))


def test_group_pool_select_vs_oracle(cuda_device):
    ops = _mods()[0]
    g = torch.Generator().manual_seed(2)
    for n, reflect, b in [(4, False, 1), (8, False, 513), (4, True, 1000), (8, True, 70000), (3, False, 17)]:
        gsize = n * (2 if reflect else 1)
        act = torch.randn(b, gsize, generator=g)
        act[0, :] = 0.25  # exact tie -> first index, as torch.argmax
        idx, rot, refl, onehot, stats = ops.group_pool_select(act.to(cuda_device), n, reflect)
        assert torch.equal(idx.cpu().long(), act.argmax(-1))
        el = O.activations_to_group_element(act, n, reflect)
        assert torch.allclose(rot.cpu(), el["rotation"], rtol=0, atol=1e-4)
        if reflect:
            assert torch.equal(refl.cpu(), el["reflection"])
        assert torch.equal(onehot.cpu(), O.activations_to_onehot(act, gsize, 1.0))
        s = stats.cpu()
        assert float(s[2]) == b
        assert float(s[0] / s[2]) == pytest.approx(float(O.prior_loss_discrete(act)), rel=1e-5)
        assert float(s[1] / s[2]) == pytest.approx(float(O.identity_metric_discrete(act)), rel=1e-6)
        assert float(s[3]) == pytest.approx(float(O.prior_loss_discrete(act)), rel=1e-5)      # the kernel's own means
        assert float(s[4]) == pytest.approx(float(O.identity_metric_discrete(act)), rel=1e-6)


def test_frames_vs_oracle(cuda_device):
    ops = _mods()[0]
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    v = torch.randn(1000, 3, 3, generator=g)
    # Gram-Schmidt has no eps: nearly collinear triples amplify fp32 rounding by 1/(r2*r3), r_i = the
    # fraction of |v_i| left after projecting out the earlier vectors; the bar scales with that conditioning
    v64 = v.double()
    e1 = v64[:, 0] / v64[:, 0].norm(dim=1, keepdim=True)
    u2 = v64[:, 1] - (v64[:, 1] * e1).sum(1, keepdim=True) * e1
    e2 = u2 / u2.norm(dim=1, keepdim=True)
    u3 = v64[:, 2] - (v64[:, 2] * e1).sum(1, keepdim=True) * e1 - (v64[:, 2] * e2).sum(1, keepdim=True) * e2
    amp = 1.0 / ((u2.norm(dim=1) / v64[:, 1].norm(dim=1)) * (u3.norm(dim=1) / v64[:, 2].norm(dim=1)))
    for modified, fn in ((False, O.gram_schmidt), (True, O.modified_gram_schmidt)):
        ours = ops.gram_schmidt3(v.to(dev), modified=modified).cpu().double()
        ref32, ref64 = fn(v).double(), fn(v64)
        e_ours = (ours - ref64).abs().amax(dim=(1, 2))
        e_ref = (ref32 - ref64).abs().amax(dim=(1, 2))
        assert bool((e_ours <= 2e-6 * amp).all()), float((e_ours / amp).max())
        assert bool((e_ref <= 2e-6 * amp).all())       # the fp32 oracle obeys the same bound
        well = amp < 4
        assert int(well.sum()) > 300 and rel_err(ours[well], ref32[well]) < 1e-5
    r = O.gram_schmidt(v[:128])
    x = torch.randn(128, 3, 1024, generator=g)   # BASELINE cfg4 shape
    assert rel_err(ops.so3_apply(x.to(dev), r.to(dev)).cpu(), O.so3_canonicalize(x, r)) < 1e-5
    x = torch.randn(3, 3, 1, generator=g)
    assert rel_err(ops.so3_apply(x.to(dev), r[:3].to(dev)).cpu(), O.so3_canonicalize(x, r[:3])) < 1e-5
    m = 5000
    rm = O.modified_gram_schmidt(torch.randn(m, 3, 3, generator=g))
    loc, vel, t = (torch.randn(m, 3, generator=g) for _ in range(3))
    cl, cv = ops.e3_apply(loc.to(dev), vel.to(dev), rm.to(dev), t.to(dev))
    ol, ov = O.e3_canonicalize(loc, vel, rm, t)
    assert rel_err(cl.cpu(), ol) < 1e-5 and rel_err(cv.cpu(), ov) < 1e-5
    assert rel_err(ops.e3_invert(ol.to(dev), rm.to(dev), t.to(dev)).cpu(), O.e3_invert(ol, rm, t)) < 1e-5
    s = ops.prior_stats_continuous(rm.to(dev)).cpu()
    assert float(s[0] / s[1]) == pytest.approx(float(O.prior_loss_continuous(rm)), rel=1e-5)
    assert float(s[3]) == pytest.approx(float(O.prior_loss_continuous(rm)), rel=1e-5)
    assert float(s[4]) == pytest.approx(float(O.identity_metric_continuous(rm)), rel=1e-5)


def test_cosine_activations_vs_oracle(cuda_device):
    ops = _mods()[0]
    g = torch.Generator().manual_seed(4)
    for b, gsize, v in [(1, 8, 128), (32, 8, 128), (5, 4, 33)]:
        vec = torch.randn(gsize * b, v, generator=g)
        ref = torch.randn(1, v, generator=g)
        act = ops.cosine_group_activations(vec.to(cuda_device), ref.to(cuda_device), gsize)
        assert rel_err(act.cpu(), O.cosine_group_activations(vec, ref, gsize)) < 1e-5


def test_edge_cases(cuda_device):
    ops = _mods()[0]
    dev = cuda_device
    # empty batch
    x0 = torch.empty(0, 3, 16, 16, device=dev)
    assert ops.warp_canonicalize(x0, torch.empty(0, dtype=torch.int32, device=dev), 4, False).shape == (0, 3, 16, 16)
    assert ops.so3_apply(torch.empty(0, 3, 8, device=dev), torch.empty(0, 3, 3, device=dev)).shape == (0, 3, 8)
    # single pixel / two-row images, sizes that are not multiples of the 32x32 tile, H != W (where the
    # reference's pad of ceil(W/2) is too small and zero fill shows up).  Single-ROW images are left out:
    # kornia's pixel normalisation degenerates there (1e-14 denominator) and the reference output is noise.
    g = torch.Generator().manual_seed(5)
    for shape in [(2, 3, 1, 1), (3, 2, 2, 37), (2, 1, 33, 65), (8, 3, 45, 45), (8, 4, 100, 31)]:
        x = torch.rand(*shape, generator=g)
        idx = torch.arange(shape[0]) % 8
        ang = torch.linspace(0.0, 360.0, 9)[:8][idx]
        y = ops.warp_canonicalize(x.to(dev), idx.to(dev).int(), 8, False)
        assert rel_err(y.cpu(), O.canonicalize_image(x, ang, None)) < RTOL, shape
        yi = ops.warp_invert(x.to(dev), idx.to(dev).int(), 8, False, False)
        assert rel_err(yi.cpu(), O.invert_image_features(x, ang, None, 8, 8, "scalar")) < RTOL, shape
    # errors surface as the reference's exception types
    with pytest.raises(ValueError):
        ops.warp_invert(torch.rand(1, 5, 8, 8, device=dev), torch.zeros(1, dtype=torch.int32, device=dev), 4, False, True)
    with pytest.raises(RuntimeError):
        ops.warp_canonicalize(torch.rand(1, 3, 8, 8), torch.zeros(1, dtype=torch.int32), 4, False)


# ---------------------------------------------------------------------------------------------------
# TMA-staged warp kernels (resample_tma.cu): every element of C8 / D4 / D8 on TMA-eligible shapes, against the
# oracle and against the generic kernel (EQB_NO_TMA=1), plus bit-exactness of the quarter-turn permutations
# ---------------------------------------------------------------------------------------------------
def _no_tma(fn):
    import os
    os.environ["EQB_NO_TMA"] = "1"
    try:
        return fn()
    finally:
        del os.environ["EQB_NO_TMA"]


@pytest.mark.parametrize("n,reflect,shape", [
    (8, False, (16, 3, 224, 224)),     # BASELINE cfg2 image shape
    (4, True, (8, 3, 224, 224)),       # D4 (cfg3 group)
    (8, True, (16, 3, 96, 64)),        # D8, H != W: zero fill reachable, quarter turns are not permutations of the grid
    (8, False, (8, 2, 52, 60)),
    (6, False, (6, 4, 64, 64)),        # N not a power of two
])
def test_warp_tma_vs_oracle_and_generic(n, reflect, shape, cuda_device):
    ops = _mods()[0]
    dev = cuda_device
    num_group = n * (2 if reflect else 1)
    g = torch.Generator().manual_seed(11)
    b = shape[0]
    x = torch.rand(*shape, generator=g)
    idx = torch.arange(b) % num_group
    ang = torch.linspace(0.0, 360.0, n + 1)[:n][idx % n]
    refl = (idx >= n).float() if reflect else None
    xd, idxd = x.to(dev), idx.to(dev).int()
    y = ops.warp_canonicalize(xd, idxd, n, reflect)
    assert rel_err(y.cpu(), O.canonicalize_image(x, ang, refl)) < RTOL
    y_gen = _no_tma(lambda: ops.warp_canonicalize(xd, idxd, n, reflect))
    assert rel_err(y.cpu(), y_gen.cpu()) < 2e-5
    yi = ops.warp_invert(xd, idxd, n, reflect, False)
    assert rel_err(yi.cpu(), O.invert_image_features(x, ang, refl, n, num_group, "scalar")) < RTOL
    f = torch.randn(b, 2 * num_group, shape[2], shape[3], generator=g)
    fr = ops.warp_invert(f.to(dev), idxd, n, reflect, True)
    assert rel_err(fr.cpu(), O.invert_image_features(f, ang, refl, n, num_group, "regular")) < RTOL
    fr_gen = _no_tma(lambda: ops.warp_invert(f.to(dev), idxd, n, reflect, True))
    assert rel_err(fr.cpu(), fr_gen.cpu()) < 2e-5
    if shape[2] == shape[3] and n % 4 == 0:
        # quarter turns of a square image are exact permutations: canonicalize rotates by -angle
        for i in range(b):
            r = int(idx[i]) % n
            if (4 * r) % n == 0 and not (reflect and int(idx[i]) >= n):
                assert torch.equal(y[i].cpu(), torch.rot90(x[i], -(4 * r // n), (1, 2))), (i, r)


def test_orbit_expand_tma_vs_oracle(cuda_device):
    ops = _mods()[0]
    g = torch.Generator().manual_seed(12)
    for n, reflect, r in [(4, True, 96), (8, False, 64)]:
        x = torch.rand(5, 3, r, r, generator=g)
        out = ops.orbit_expand(x.to(cuda_device), (r + 1) // 2, r, n, reflect)
        ref = O.group_augment(x, n, reflect, r)
        assert out.shape == ref.shape
        assert rel_err(out.cpu(), ref) < RTOL
        out_gen = _no_tma(lambda: ops.orbit_expand(x.to(cuda_device), (r + 1) // 2, r, n, reflect))
        assert rel_err(out.cpu(), out_gen.cpu()) < 2e-5


def test_full_size_round_trip_properties(cuda_device):
    """BASELINE cfg2 size (512 x 3 x 224 x 224): size-independent properties instead of an oracle run."""
    ops = _mods()[0]
    dev = cuda_device
    n = 8
    x = torch.rand(512, 3, 224, 224, generator=torch.Generator().manual_seed(13)).to(dev)
    idx = (torch.arange(512) % n).to(dev).int()
    y = ops.warp_canonicalize(x, idx, n, False)
    back = ops.warp_invert(y, idx, n, False, False)
    quarter = (idx % 2 == 0)
    # quarter turns: canonicalize is a permutation (== rot90 by -angle) and invert undoes it bit for bit
    for r in range(0, n, 2):
        sel = idx == r
        assert torch.equal(y[sel], torch.rot90(x[sel], -(r // 2), (2, 3)))
    assert torch.equal(back[quarter], x[quarter])
    # 45-degree elements: two bilinear resamplings of iid noise do not return x, but they commute with the group:
    # rotating by g then by h equals rotating by g+h up to interpolation, and means are preserved inside the disc
    odd = ~quarter
    assert float((y[odd].mean() - x[odd].mean()).abs()) < 5e-3
    # linearity: warp(a*x1 + x2) == a*warp(x1) + warp(x2) (same taps, fp32 rounding only)
    x2 = torch.rand(512, 3, 224, 224, generator=torch.Generator().manual_seed(14)).to(dev)
    lhs = ops.warp_canonicalize(0.5 * x + x2, idx, n, False)
    rhs = 0.5 * y + ops.warp_canonicalize(x2, idx, n, False)
    assert float((lhs - rhs).abs().max()) < 1e-6
    # regular representation: channel roll by the group index composes with the inverse roll to the identity
    f = torch.randn(64, 16, 224, 224, generator=torch.Generator().manual_seed(15)).to(dev)
    i4 = ((torch.arange(64) % 4) * 2).to(dev).int()                      # quarter turns of C8
    fwd = ops.warp_invert(f, i4, n, False, True)
    inv = ops.warp_invert(fwd, (n - i4) % n, n, False, True)
    assert torch.equal(inv, f)


# ---------------------------------------------------------------------------------------------------
# tcgen05 conv stack (gconv_stack_tc.cu) against the fp32 SIMT kernel (EQB_NO_TC=1) and the fp64 oracle
# ---------------------------------------------------------------------------------------------------
def _stack_act(net_cpu, x, dev, no_tc, no_pair=False):
    import copy
    import os
    if no_tc:
        os.environ["EQB_NO_TC"] = "1"
    if no_pair:
        os.environ["EQB_TC_PAIR"] = "0"
    try:
        net = copy.deepcopy(net_cpu).to(dev)   # a fresh module packs its operands under the current setting
        with torch.no_grad():
            out = net(x.to(dev)).cpu()
        torch.cuda.synchronize()
        return out
    finally:
        os.environ.pop("EQB_NO_TC", None)
        os.environ.pop("EQB_TC_PAIR", None)


@pytest.mark.parametrize("group_type,n,cout,k,res,b", [
    ("rotation", 8, 32, 5, 96, 6),          # cfg2 network: N = 256, lift K = 75
    ("rotation", 4, 16, 5, 32, 9),          # cfg1 network: N = 64
    ("roto-reflection", 4, 8, 3, 40, 5),    # D4: N = 64, lift K = 27
    ("rotation", 8, 12, 7, 50, 3),          # N = 96 (Npad 128), lift K = 147 (10 slabs), partial last tile
    ("roto-reflection", 8, 16, 1, 20, 4),   # D8: N = 256, 1x1 lift (K = 3)
])
def test_tcgen05_stack_vs_simt_and_oracle(group_type, n, cout, k, res, b, cuda_device):
    _, _, _, Net = _mods()
    reflect = group_type == "roto-reflection"
    torch.manual_seed(20)
    net = Net((3, res, res), cout, k, group_type, n, 3, device="cpu")
    with torch.no_grad():
        for m in net.eqv_network:
            if hasattr(m, "bias"):
                m.bias.uniform_(-0.1, 0.1)
    lay = [(m.weights.detach().clone(), m.bias.detach().clone()) for m in net.eqv_network if hasattr(m, "weights")]
    x = torch.rand(b, 3, res, res, generator=torch.Generator().manual_seed(21))
    act_tc = _stack_act(net, x, cuda_device, no_tc=False)
    act_simt = _stack_act(net, x, cuda_device, no_tc=True)
    act64 = O.custom_equivariant_network(x.double(), [(w.double(), bb.double()) for w, bb in lay], n, reflect)
    act32 = O.custom_equivariant_network(x, lay, n, reflect)
    assert rel_err(act_simt, act64) < 1e-5
    # 3xTF32 products are fp32-grade; the tensor core's accumulator rounds differently from an fp32 FMA chain
    assert rel_err(act_tc, act64) < 2e-5
    assert rel_err(act_tc, act32) < RTOL
    assert_index_parity(act_tc, act32, act64, act_tc.argmax(-1))
    if cout * n * (2 if reflect else 1) == 256 and 3 * k * k <= 80:
        # this shape runs the CTA-pair kernel (cta_group::2) by default: pin it against the single-CTA kernel too
        act_single = _stack_act(net, x, cuda_device, no_tc=False, no_pair=True)
        assert rel_err(act_single, act64) < 2e-5
        assert rel_err(act_tc, act_single) < 2e-6


def test_tcgen05_stack_full_batch_properties(cuda_device):
    """cfg2 batch (512 x 3 x 96 x 96 after the pre-network transform): the stack is per-sample independent and
    rotation-equivariant, which pins it at a size the oracle cannot reach in seconds."""
    _, _, _, Net = _mods()
    dev = cuda_device
    torch.manual_seed(22)
    net = Net((3, 96, 96), 32, 5, "rotation", 8, 3, device="cpu").to(dev)
    x = torch.rand(512, 3, 96, 96, generator=torch.Generator().manual_seed(23)).to(dev)
    with torch.no_grad():
        act = net(x)
        # per-sample independence: any sub-batch gives the same rows (different work-item partition)
        sub = net(x[37:45])
        assert torch.equal(act[37:45], sub)
        # C4 subgroup equivariance (exact for quarter turns of the input): activations roll by 2 positions of C8
        act_r = net(torch.rot90(x[:64], 1, (2, 3)))
    scale = float(act.abs().max())
    for shift in (2, -2):
        err = float((act_r - torch.roll(act[:64], shift, dims=1)).abs().max())
        if err < 2e-5 * scale:
            break
    else:
        raise AssertionError("activations of a 90-degree rotated batch are not a roll by two group positions")


# ---------------------------------------------------------------------------------------------------
# a7: e2cnn-style expanded-filter conv stack (conv_stack.cu) vs the oracle's dense restatement
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("group_type,n,cout,k,layers,res,b", [
    ("rotation", 4, 32, 5, 3, 44, 3),        # the example config's network (N = 128, inner K = 3200): tcgen05 inner layer
    ("rotation", 4, 8, 3, 4, 36, 4),         # 4 layers: two tcgen05 layers, fp16-pair hand-over between them
    ("rotation", 8, 32, 3, 3, 30, 2),        # N = 256 on the tensor path
    ("rotation", 4, 32, 3, 2, 32, 3),        # the reference's own test fixture (tests/.../test_discrete_group.py:22-40)
    ("rotation", 4, 8, 5, 3, 40, 5),         # N = 32, inner K = 800
    ("roto-reflection", 4, 6, 3, 3, 29, 4),  # D4: N = 48 (Npad 64), odd sizes, partial tiles
    ("rotation", 8, 32, 3, 2, 24, 2),        # N = 256
    ("rotation", 4, 4, 3, 1, 20, 3),         # single layer: conv + pool only
])
def test_escnn_expanded_stack_vs_oracle(group_type, n, cout, k, layers, res, b, cuda_device):
    from equiadapt_b200.images.canonicalization_networks.escnn_networks import ESCNNEquivariantNetwork
    dev = cuda_device
    g = n * (2 if group_type == "roto-reflection" else 1)
    torch.manual_seed(30)
    net = ESCNNEquivariantNetwork((3, res, res), cout, k, group_type, n, layers, device=str(dev)).eval()
    with torch.no_grad():
        for l in range(layers):
            net.biases[l].copy_(torch.empty(cout).uniform_(-0.2, 0.2).repeat_interleave(g))
            if l < layers - 1:
                net.bn_weight[l].uniform_(0.5, 1.5)
                net.bn_bias[l].uniform_(-0.3, 0.3)
                getattr(net, f"bn_running_mean_{l}").uniform_(-0.2, 0.2)
                getattr(net, f"bn_running_var_{l}").uniform_(0.5, 2.0)
    x = torch.rand(b, 3, res, res, generator=torch.Generator().manual_seed(31))
    with torch.no_grad():
        act = net(x.to(dev)).cpu()
        scales, shifts = net.folded_affine()
    torch.cuda.synchronize()
    filt = [f.detach().cpu() for f in net.filters]
    bias = [t.detach().cpu() for t in net.biases]
    sc, sh = [t.cpu() for t in scales] + [None], [t.cpu() for t in shifts] + [None]
    act32 = O.expanded_conv_network(x, filt, bias, sc, sh, g)
    dd = lambda ts: [None if t is None else t.double() for t in ts]
    act64 = O.expanded_conv_network(x.double(), dd(filt), dd(bias), dd(sc), dd(sh), g)
    assert act.shape == (b, g)
    # inner layers on the tensor cores (L >= 3): TMEM accumulation truncates, the error grows with K = Cin*k*k
    # (1.4e-5 at K = 3200); SIMT fp32 FMA chains stay below 1e-5
    assert rel_err(act, act64) < (3e-5 if layers >= 2 else 1e-5)
    assert rel_err(act, act32) < RTOL
    assert_index_parity(act, act32, act64, act.argmax(-1))
    # the all-SIMT path (no tcgen05 inner layers, no fold) computes the same activations
    import os
    os.environ["EQB_CONV_NO_TC"] = "1"
    os.environ["EQB_CONV_NO_FOLD"] = "1"
    try:
        with torch.no_grad():
            act_simt = net(x.to(dev)).cpu()
        torch.cuda.synchronize()
    finally:
        del os.environ["EQB_CONV_NO_TC"], os.environ["EQB_CONV_NO_FOLD"]
    assert rel_err(act_simt, act64) < 1e-5 and rel_err(act, act_simt) < 3e-5


def test_escnn_network_drops_into_canonicalizer_and_is_equivariant(cuda_device):
    """The reference's smoke test (tests/images/canonicalization/test_discrete_group.py:13-69) with its fixture
    hyper-parameters, plus what it does not assert: C4 equivariance of the group activations."""
    from equiadapt_b200.images.canonicalization_networks.escnn_networks import ESCNNEquivariantNetwork
    _, GEIC, _, _ = _mods()
    dev = cuda_device
    torch.manual_seed(32)
    net = ESCNNEquivariantNetwork((3, 32, 32), 32, 3, "rotation", 4, 2, device=str(dev))
    can = GEIC(net, SimpleNamespace(beta=0.1, input_crop_ratio=0.9, resize_shape=(32, 32)), (3, 64, 64)).eval()
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(33)).to(dev)
    with torch.no_grad():
        y = can(x)
        assert y.shape == x.shape
        for rep, c in (("regular", 12), ("scalar", 3)):
            f = torch.randn(1, c, 64, 64, generator=torch.Generator().manual_seed(34)).to(dev)
            assert can.invert_canonicalization(f, induced_rep_type=rep).shape == f.shape
        xs = torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(35)).to(dev)
        a0 = net(xs)
        a1 = net(torch.rot90(xs, 1, (2, 3)))
    scale = float(a0.abs().max())
    assert min(float((a1 - torch.roll(a0, s, dims=1)).abs().max()) for s in (1, -1)) < 2e-5 * scale


# ---------------------------------------------------------------------------------------------------
# N1: frame-predicting networks (vn_networks.cu) vs golden vectors of the unmodified reference modules
# ---------------------------------------------------------------------------------------------------
def _sd(g):
    return {k[3:]: v for k, v in g.items() if k.startswith("sd.")}


def test_vnsmall_vs_reference_golden_and_end_to_end(cuda_device):
    from equiadapt_b200.pointcloud.canonicalization.continuous_group import EquivariantPointcloudCanonicalization
    from equiadapt_b200.pointcloud.canonicalization_networks.equivariant_networks import VNSmall
    g = load_golden("vnsmall")
    dev = cuda_device
    net = VNSmall(SimpleNamespace(n_knn=int(g["n_knn"]), pooling="mean"))
    missing = net.load_state_dict(_sd(g), strict=False)          # the reference's own state-dict keys
    assert not missing.unexpected_keys and all("num_batches_tracked" in k for k in missing.missing_keys)
    net = net.to(dev).eval()
    with torch.no_grad():
        out = net(g["x"].to(dev)).cpu()
    assert rel_err(out, g["out"]) < RTOL
    # BASELINE configs[3] end to end: clouds -> VNSmall -> Gram-Schmidt -> R x, against the oracle chain
    x = torch.randn(5, 3, 1024, generator=torch.Generator().manual_seed(40))
    can = EquivariantPointcloudCanonicalization(net, SimpleNamespace()).eval()
    with torch.no_grad():
        y = can(x.to(dev)).cpu()
    vec = O.vnsmall_forward(x, _sd(g), int(g["n_knn"]))
    rot = O.gram_schmidt(vec)
    assert rel_err(can.canonicalization_info_dict["group_element_matrix_representation"].cpu(), rot) < 5e-4
    assert rel_err(y, O.so3_canonicalize(x, rot)) < 5e-4
    # SO(3) equivariance of the predicted frame: rotating the cloud rotates the canonical cloud not at all
    q = O.gram_schmidt(torch.randn(1, 3, 3, generator=torch.Generator().manual_seed(41)))[0]
    if torch.det(q) < 0:
        q[2] = -q[2]
    with torch.no_grad():
        y_rot = can(torch.einsum("ij,bjn->bin", q, x).to(dev)).cpu()
    assert rel_err(y_rot, y) < 5e-4


@pytest.mark.parametrize("tag,nonlin,feat,pool,trans", [("relu_p", "relu", "p", "mean", False),
                                                        ("softplus_pvac", "softplus", "pvac", "sum", True)])
def test_vndeepsets_vs_reference_golden_and_end_to_end(tag, nonlin, feat, pool, trans, cuda_device):
    from equiadapt_b200.nbody.canonicalization.euclidean_group import EuclideanGroupNBody
    from equiadapt_b200.nbody.canonicalization_networks.custom_equivariant_networks import VNDeepSets
    g = load_golden("vndeepsets_" + tag)
    dev = cuda_device
    hp = SimpleNamespace(out_dim=4, hidden_dim=16, layer_pooling=pool, final_pooling="mean", num_layers=4, nonlinearity=nonlin,
                         canon_feature=feat, canon_translation=trans, angular_feature=0, dropout=0.5, batch_size=9)
    net = VNDeepSets(hp, device=str(dev))
    net.load_state_dict(_sd(g))
    net.eval()
    loc, vel, ch, edges = g["loc"].to(dev), g["vel"].to(dev), g["charges"].to(dev), g["edges"].long().to(dev)
    with torch.no_grad():
        rv, t = net(None, loc, edges, vel, None, ch)
    assert rel_err(rv.cpu(), g["rot_vectors"]) < RTOL
    assert rel_err(t.cpu(), g["translation"]) < RTOL
    # BASELINE configs[4] end to end through the canonicalizer (kwarg order as the reference's caller)
    can = EuclideanGroupNBody(net).eval()
    nodes = torch.sqrt(torch.sum(vel ** 2, dim=1)).unsqueeze(1)
    with torch.no_grad():
        cl, cv = can(nodes, None, loc=loc, edges=edges, vel=vel, edge_attr=None, charges=ch)
    rot = O.modified_gram_schmidt(g["rot_vectors"])
    ol, ov = O.e3_canonicalize(g["loc"], g["vel"], rot, g["translation"])
    assert rel_err(cl.cpu(), ol) < 5e-4 and rel_err(cv.cpu(), ov) < 5e-4
    assert torch.isfinite(can.get_prior_regularization_loss()).item()


# ---------------------------------------------------------------------------------------------------
# N2: continuous-group image canonicalization (eqb_warp_affine) vs goldens of the unmodified reference
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["rot", "refl", "gray"])
def test_continuous_canonicalize_vs_reference_golden(tag, cuda_device):
    from unittest import mock
    from equiadapt_b200.images.canonicalization.continuous_group import ContinuousGroupImageCanonicalization
    g = load_golden("image_cont_" + tag)
    dev = cuda_device
    x = g["x"]
    can = ContinuousGroupImageCanonicalization(torch.nn.Identity(),
                                               SimpleNamespace(input_crop_ratio=0.9, resize_shape=(16, 16)), tuple(x.shape[1:]))
    element = {"rotation": g["rotation"].clone().to(dev)}
    if "reflection" in g:
        element["reflection"] = g["reflection"].to(dev)
    with mock.patch.object(can, "get_groupelement", return_value=element), torch.no_grad():
        y = can.canonicalize(x.to(dev))
    assert rel_err(y.cpu(), g["y"]) < RTOL
    assert torch.equal(element["rotation"].cpu(), g["rotation_after"])     # same in-place side effect as the reference
    # the generic (non-TMA) kernel gives the same image
    element2 = {k: (v.clone() if k != "rotation" else g["rotation"].clone().to(dev)) for k, v in element.items()}
    with mock.patch.object(can, "get_groupelement", return_value=element2), torch.no_grad():
        y2 = _no_tma(lambda: can.canonicalize(x.to(dev)))
    assert rel_err(y2.cpu(), y.cpu()) < 1e-5
    # pre-network transform of the class (reference test: tests/.../test_continuous_group.py:72-91)
    if x.shape[1] != 1:
        pre = can.transformations_before_canonicalization_network_forward(x.to(dev))
        assert tuple(pre.shape[-2:]) == (16, 16)


@pytest.mark.parametrize("tag,group_type", [("rot", "rotation"), ("refl", "roto-reflection")])
def test_continuous_group_augment_vs_reference_golden(tag, group_type, cuda_device):
    from equiadapt_b200.images.canonicalization.continuous_group import OptimizedSteerableImageCanonicalization
    g = load_golden("image_cont_augment_" + tag)
    dev = cuda_device
    can = OptimizedSteerableImageCanonicalization(
        torch.nn.Identity(), SimpleNamespace(input_crop_ratio=0.9, resize_shape=(16, 16), group_type=group_type), (3, 32, 32))
    refl = g["reflect"].to(dev) if "reflect" in g else None
    with torch.no_grad():
        aug, mats = can.group_augment(g["x"].to(dev), angles=g["angles"].to(dev), reflect=refl)
    assert rel_err(aug.cpu(), g["aug"]) < RTOL
    assert rel_err(mats.cpu(), g["mats"]) < 1e-6


def test_steerable_canonicalizer_end_to_end(cuda_device):
    """SteerableImageCanonicalization with a stand-in network (the reference's e2cnn steerable network is out of scope):
    canonicalize == oracle chain, prior loss / identity metric == a14, and the full-size batch is rotation-consistent."""
    from equiadapt_b200.images.canonicalization.continuous_group import SteerableImageCanonicalization
    dev = cuda_device

    class Net(torch.nn.Module):
        group_type = "rotation"

        def forward(self, x):   # two 2-vectors per sample from image moments (any deterministic function will do)
            m = x.mean(dim=(1,))
            gx = (m[:, :, 1:] - m[:, :, :-1]).mean(dim=(1, 2))
            gy = (m[:, 1:, :] - m[:, :-1, :]).mean(dim=(1, 2))
            v = torch.stack([gx + 0.3, gy - 0.2], dim=1)
            return torch.stack([v, v.flip(1)], dim=1)

    can = SteerableImageCanonicalization(Net(), SimpleNamespace(input_crop_ratio=0.8, resize_shape=(32, 32)), (3, 64, 64)).eval()
    x = torch.rand(7, 3, 64, 64, generator=torch.Generator().manual_seed(50))
    with torch.no_grad():
        y = can(x.to(dev))
        rep = can.canonicalization_info_dict["group_element_matrix_representation"].cpu()
        loss, ident = float(can.get_prior_regularization_loss()), float(can.get_identity_metric())
    vec = Net()(O.pre_network_transform(x, (3, 64, 64), 0.8, (32, 32)))[:, 0]
    v1 = vec / vec.norm(dim=1, keepdim=True)
    rot = torch.stack([v1, torch.stack([-v1[:, 1], v1[:, 0]], 1)], 1)
    assert rel_err(y.cpu(), O.canonicalize_image_continuous(x, rot, None)) < 5e-4   # the stand-in net runs in torch on both sides
    neg = rot.clone()
    neg[:, [0, 1], [1, 0]] *= -1
    assert abs(loss - float(O.prior_loss_continuous(neg))) < 1e-4 and abs(ident - (1 - loss)) < 1e-6
    with pytest.raises(NotImplementedError):
        can.invert_canonicalization(y)


# ---------------------------------------------------------------------------------------------------
# stream / graph behaviour of the public path
# ---------------------------------------------------------------------------------------------------
def test_step_is_cuda_graph_capturable(cuda_device):
    """canonicalize + invert + prior statistic enqueue kernels only (no host synchronisation, no allocation outside
    torch's allocator), so a step can be captured once and replayed on new data in the same buffers."""
    _, GEIC, _, Net = _mods()
    dev = cuda_device
    torch.manual_seed(60)
    net = Net((3, 32, 32), 8, 5, "rotation", 8, 3, device="cpu").to(dev)
    can = GEIC(net, SimpleNamespace(beta=1.0, input_crop_ratio=0.8, resize_shape=32), (3, 64, 64)).eval()
    gen = torch.Generator().manual_seed(61)
    x0, x1 = torch.rand(16, 3, 64, 64, generator=gen), torch.rand(16, 3, 64, 64, generator=gen)
    xs = x0.to(dev).clone()

    def step():
        y = can(xs)
        z = can.invert_canonicalization(y, induced_rep_type="scalar")
        return y, z, can.get_prior_regularization_loss(), can.get_identity_metric()

    with torch.no_grad():
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                step()                                   # warm-up: packs the parameters, sizes the allocator pools
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            y, z, loss, ident = step()
        xs.copy_(x1.to(dev))
        graph.replay()
        torch.cuda.synchronize()
        got = (y.clone(), z.clone(), float(loss), float(ident))
        ref = step()
        torch.cuda.synchronize()
    assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
    assert got[2] == float(ref[2]) and got[3] == float(ref[3])


def test_empty_batches_of_the_network_entry_points(cuda_device):
    from equiadapt_b200.images.canonicalization_networks.escnn_networks import ESCNNEquivariantNetwork
    from equiadapt_b200.pointcloud.canonicalization_networks.equivariant_networks import VNSmall
    ops = _mods()[0]
    dev = cuda_device
    net = ESCNNEquivariantNetwork((3, 20, 20), 4, 3, "rotation", 4, 2, device=str(dev)).eval()
    vn = VNSmall(SimpleNamespace(n_knn=20, pooling="mean")).to(dev).eval()
    with torch.no_grad():
        assert net(torch.empty(0, 3, 20, 20, device=dev)).shape == (0, 4)
        assert vn(torch.empty(0, 3, 64, device=dev)).shape == (0, 3, 3)
        assert ops.warp_affine(torch.empty(0, 3, 56, 56, device=dev), torch.empty(0, 2, 2, device=dev), None, True, 28, 28.0, 28.0).shape == (0, 3, 56, 56)
    with pytest.raises(ValueError):
        vn(torch.rand(2, 3, 8, device=dev))              # fewer points than n_knn
    with pytest.raises(NotImplementedError):
        net.train()(torch.rand(1, 3, 20, 20, device=dev))


def test_cfg2_index_parity_on_smooth_images(cuda_device):
    """SURVEY.md section 7 (hard part 1): the headline parity run - BASELINE configs[1] network and image size, 32 smooth
    non-zero-mean images (wide arg-max margins).  The discrete index must equal the fp64 oracle's on every sample
    whose margin exceeds the combined fp32 noise, and our agreement with fp64 must not be worse than the fp32
    reference's own."""
    _, GEIC, _, Net = _mods()
    dev = cuda_device
    torch.manual_seed(0)
    net = Net((3, 96, 96), 32, 5, "rotation", 8, 3, device="cpu")
    lay = [(m.weights.detach().clone(), m.bias.detach().clone()) for m in net.eqv_network if hasattr(m, "weights")]
    can = GEIC(net.to(dev), SimpleNamespace(beta=1.0, input_crop_ratio=0.8, resize_shape=96), (3, 224, 224)).eval()
    g = torch.Generator().manual_seed(7)
    low = torch.randn(32, 3, 14, 14, generator=g)
    x = torch.nn.functional.interpolate(low, size=(224, 224), mode="bicubic", align_corners=False) + 0.5
    with torch.no_grad():
        y = can(x.to(dev))
    act = can.canonicalization_info_dict["group_activations"].cpu()
    idx = can.canonicalization_info_dict["group_element"].index.cpu().long()
    act32 = O.custom_equivariant_network(O.pre_network_transform(x, (3, 224, 224), 0.8, 96), lay, 8, False)
    act64 = O.custom_equivariant_network(O.pre_network_transform(x.double(), (3, 224, 224), 0.8, 96),
                                         [(w.double(), b.double()) for w, b in lay], 8, False)
    assert rel_err(act, act64) < 2e-5
    assert_index_parity(act, act32, act64, idx)
    agree_ours = int((idx == act64.argmax(-1)).sum())
    agree_ref = int((act32.argmax(-1) == act64.argmax(-1)).sum())
    assert agree_ours >= agree_ref
    ang = torch.linspace(0.0, 360.0, 9)[:8][idx]
    assert rel_err(y.cpu(), O.canonicalize_image(x, ang, None)) < RTOL


def test_host_streamed_pipeline_equals_direct_call(cuda_device):
    """HostStreamedCanonicalizer (pinned host in / out, sharded H2D / compute / D2H streams) gives bit-identical outputs
    and the same prior statistic as one un-sharded call on device tensors, including a ragged last shard."""
    from equiadapt_b200.host_pipeline import HostStreamedCanonicalizer
    _, GEIC, _, Net = _mods()
    dev = cuda_device
    torch.manual_seed(70)
    net = Net((3, 32, 32), 8, 5, "rotation", 8, 3, device="cpu").to(dev)
    can = GEIC(net, SimpleNamespace(beta=1.0, input_crop_ratio=0.8, resize_shape=32), (3, 64, 64)).eval()
    x = torch.rand(77, 3, 64, 64, generator=torch.Generator().manual_seed(71))
    with torch.no_grad():
        y = can(x.to(dev))
        z = can.invert_canonicalization(y, induced_rep_type="scalar")
        loss, ident = float(can.get_prior_regularization_loss()), float(can.get_identity_metric())
        x_host, out_host = x.pin_memory(), torch.empty_like(x).pin_memory()
        pipe = HostStreamedCanonicalizer(can, None, "scalar", shard=16, slots=3, device=dev)
        for _ in range(2):                      # second call reuses the slots
            out_host.zero_()
            l2, i2 = pipe(x_host, out_host)
            torch.cuda.synchronize()
            assert torch.equal(out_host, z.cpu())
            assert abs(float(l2) - loss) < 1e-6 and abs(float(i2) - ident) < 1e-6
    with pytest.raises(ValueError):
        pipe(x, out_host)                       # unpinned input


@pytest.mark.parametrize("b", [1, 2, 3, 17, 150])
def test_stack_kernels_agree_for_any_batch_size(b, cuda_device):
    """Work partition of the persistent kernels (items = image x chunk of tiles, clusters of two CTAs) for batch sizes
    below, at and above the SM count: CTA-pair kernel == single-CTA tcgen05 kernel == SIMT kernel on every row."""
    _, _, _, Net = _mods()
    torch.manual_seed(80)
    net = Net((3, 40, 40), 32, 5, "rotation", 8, 3, device="cpu")        # N = 256, K = 75: pair kernel by default
    x = torch.rand(b, 3, 40, 40, generator=torch.Generator().manual_seed(81))
    act_pair = _stack_act(net, x, cuda_device, no_tc=False)
    act_single = _stack_act(net, x, cuda_device, no_tc=False, no_pair=True)
    act_simt = _stack_act(net, x, cuda_device, no_tc=True)
    assert act_pair.shape == (b, 8)
    assert rel_err(act_pair, act_simt) < 2e-5 and rel_err(act_single, act_simt) < 2e-5
    assert rel_err(act_pair, act_single) < 2e-6


@pytest.mark.parametrize("scale", [1e-6, 1e-3, 1.0, 255.0, 1e5])
def test_stack_operand_scaling_covers_the_input_range(scale, cuda_device):
    """The fp16 hi/lo operands are pre-scaled by powers of two derived from max|x| per call: images in any range
    (normalised, 0..255, tiny) keep fp32-grade activations; an all-zero batch and one with a single huge outlier
    pixel stay finite and match the SIMT fp32 kernel."""
    _, _, _, Net = _mods()
    torch.manual_seed(90)
    net = Net((3, 36, 36), 32, 5, "rotation", 8, 3, device="cpu")
    with torch.no_grad():
        for m in net.eqv_network:
            if hasattr(m, "bias"):
                m.bias.uniform_(-0.05, 0.05)
    g = torch.Generator().manual_seed(91)
    x = (torch.rand(4, 3, 36, 36, generator=g) - 0.3) * scale
    x[1] = 0.0                                   # an all-zero image inside the batch
    x[2, 0, 5, 7] = 50.0 * scale                 # one outlier pixel sets the batch maximum
    lay = [(m.weights.detach().clone(), m.bias.detach().clone()) for m in net.eqv_network if hasattr(m, "weights")]
    act64 = O.custom_equivariant_network(x.double(), [(w.double(), b.double()) for w, b in lay], 8, False)
    act_tc = _stack_act(net, x, cuda_device, no_tc=False)
    act_simt = _stack_act(net, x, cuda_device, no_tc=True)
    assert torch.isfinite(act_tc).all()
    assert rel_err(act_simt, act64) < 1e-5
    assert rel_err(act_tc, act64) < 2e-5
    zero = _stack_act(net, torch.zeros(2, 3, 36, 36), cuda_device, no_tc=False)   # max|x| = 0: scale falls back to 1
    assert torch.isfinite(zero).all() and rel_err(zero, _stack_act(net, torch.zeros(2, 3, 36, 36), cuda_device, no_tc=True)) < 1e-5


# ---------------------------------------------------------------------------------------------------
# N3 (partial): gradients of the warps with respect to their image argument vs torch autograd through the oracle
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("group_type,n", [("rotation", 4), ("rotation", 8), ("roto-reflection", 4), ("roto-reflection", 8)])
def test_warp_gradients_vs_oracle_autograd(group_type, n, cuda_device):
    ops = _mods()[0]
    dev = cuda_device
    reflect = group_type == "roto-reflection"
    G = n * (2 if reflect else 1)
    g = torch.Generator().manual_seed(100 + n + reflect)
    b, h, w = 2 * G, 40, 40
    idx = torch.arange(b) % G
    ang = torch.linspace(0.0, 360.0, n + 1)[:n][idx % n]
    refl = (idx >= n).float() if reflect else None
    # canonicalize: d <y, r> / d x
    x = torch.rand(b, 3, h, w, generator=g)
    r = torch.randn(b, 3, h, w, generator=g)
    xo = x.clone().requires_grad_(True)
    (O.canonicalize_image(xo, ang, refl) * r).sum().backward()
    xd = x.to(dev).requires_grad_(True)
    y = ops.warp_canonicalize_autograd(xd, idx.to(dev).int(), n, reflect)
    (y * r.to(dev)).sum().backward()
    assert rel_err(xd.grad.cpu(), xo.grad) < RTOL
    # invert: scalar and regular representations
    for rep, c in (("scalar", 3), ("regular", 2 * G)):
        f = torch.randn(b, c, h, w, generator=g)
        rr = torch.randn(b, c, h, w, generator=g)
        fo = f.clone().requires_grad_(True)
        (O.invert_image_features(fo, ang, refl, n, G, rep) * rr).sum().backward()
        fd = f.to(dev).requires_grad_(True)
        out = ops.warp_invert_autograd(fd, idx.to(dev).int(), n, reflect, rep == "regular")
        (out * rr.to(dev)).sum().backward()
        assert rel_err(fd.grad.cpu(), fo.grad) < RTOL


def test_prediction_network_trains_through_invert_canonicalization(cuda_device):
    """A frozen canonicalizer in front of / behind a trainable prediction network: loss.backward() reaches the
    prediction network's parameters through invert_canonicalization (the reference's segmentation-style use)."""
    _, GEIC, _, Net = _mods()
    dev = cuda_device
    torch.manual_seed(110)
    can = GEIC(Net((3, 32, 32), 8, 5, "rotation", 8, 3, device="cpu").to(dev),
               SimpleNamespace(beta=1.0, input_crop_ratio=0.8, resize_shape=32), (3, 64, 64)).eval()
    for p in can.parameters():
        p.requires_grad_(False)
    pred = torch.nn.Conv2d(3, 8, 3, padding=1).to(dev)
    x = torch.rand(4, 3, 64, 64, generator=torch.Generator().manual_seed(111)).to(dev)
    out = can.invert_canonicalization(pred(can(x)), induced_rep_type="regular")
    out.square().mean().backward()
    assert pred.weight.grad is not None and torch.isfinite(pred.weight.grad).all() and float(pred.weight.grad.abs().sum()) > 0


# ---- N4 (second half): evaluation-time group orbit -------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["c4", "d8", "c6_gray", "d5_rect"])
def test_group_inference_orbit_vs_torchvision_golden(tag, cuda_device):
    """eqb_orbit_rotate_nearest vs torchvision's Pad/hflip/rotate(NEAREST)/CenterCrop (inference_utils.py:97-122):
    bit-exact, including the pixels whose source coordinate sits on a rounding tie."""
    ops = _mods()[0]
    g = load_golden("group_inference_orbit_" + tag)
    out = ops.orbit_rotate_nearest(g["x"].to(cuda_device), int(g["num_rotations"]), bool(int(g["reflect"])))
    assert out.shape == g["orbit"].shape
    assert torch.equal(out.cpu(), g["orbit"])


@pytest.mark.parametrize("n,reflect,shape", [(8, True, (64, 3, 224, 224)), (4, False, (5, 3, 224, 224)), (16, True, (2, 1, 95, 131)),
                                              (7, False, (3, 2, 33, 33)), (64, False, (1, 3, 40, 40))])
def test_group_inference_orbit_vs_oracle_full_size(n, reflect, shape, cuda_device):
    ops = _mods()[0]
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(n))
    out = ops.orbit_rotate_nearest(x.to(cuda_device), n, reflect).cpu()
    # the oracle on one image (it is independent of the batch index), every plane against it
    want, margin = O.group_inference_orbit(x[:1], n, reflect, return_margin=True)
    assert torch.equal(out[:, :1], want)
    # size-independent properties: element 0 is the identity; every output value is either 0 or a value of its plane;
    # the nearest-neighbour source map is the same for every plane
    assert torch.equal(out[0], x)
    ramp = torch.arange(1, shape[-2] * shape[-1] + 1, dtype=torch.float32).reshape(1, 1, *shape[-2:]).to(cuda_device)
    src = ops.orbit_rotate_nearest(ramp, n, reflect).cpu().long()          # 0 = outside, else 1 + flat source index
    flat = torch.cat([torch.zeros(*shape[:2], 1), x.reshape(*shape[:2], -1)], -1)
    for gi in (0, out.shape[0] // 2, out.shape[0] - 1):
        gathered = torch.gather(flat, 2, src[gi].reshape(1, 1, -1).expand(shape[0], shape[1], -1)).reshape(shape)
        assert torch.equal(out[gi], gathered)
    assert int((margin < 1e-4).sum()) >= 0


def test_group_inference_empty_and_errors(cuda_device):
    ops = _mods()[0]
    assert ops.orbit_rotate_nearest(torch.zeros(0, 3, 8, 8, device=cuda_device), 4, True).shape == (8, 0, 3, 8, 8)
    with pytest.raises(NotImplementedError):
        ops.orbit_rotate_nearest(torch.zeros(1, 3, 8, 8, device=cuda_device), 65, False)
    with pytest.raises(ValueError):
        ops.orbit_rotate_nearest(torch.zeros(1, 3, 8, 8, device=cuda_device), 0, False)


@pytest.mark.parametrize("group_type", ["rotation", "roto-reflection"])
def test_group_inference_metrics_match_reference_loop(group_type, cuda_device):
    """GroupInference (mirror of inference_utils.py:80-168) around a real canonicalizer: metric keys and values equal
    the reference's per-element loop run on the oracle's orbit through the same canonicalizer / prediction network."""
    from equiadapt_b200.images.inference import GroupInference, VanillaInference, get_inference_method
    dev = cuda_device
    g = load_golden("image_c4_cfg1")
    can = build_canonicalizer(g, dev)
    torch.manual_seed(3)
    pred = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 32 * 32, 10)).to(dev).eval()
    hp = SimpleNamespace(method="group", group_type=group_type, num_rotations=4)
    inf = get_inference_method(can, pred, 10, hp, (3, 32, 32))
    assert isinstance(inf, GroupInference)
    assert isinstance(get_inference_method(can, pred, 10, {"method": "vanilla"}, (3, 32, 32)), VanillaInference)
    with pytest.raises(ValueError):
        get_inference_method(can, pred, 10, {"method": "nope"}, (3, 32, 32))
    x = torch.randn(16, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    y = torch.randint(0, 10, (16,), generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        metrics = inf.get_inference_metrics(x.to(dev), y.to(dev))
        orbit = O.group_inference_orbit(x, 4, group_type == "roto-reflection")
        accs = []
        for gi in range(orbit.shape[0]):
            logits = pred(can(orbit[gi].to(dev)))
            accs.append((logits.argmax(-1).cpu() == y).float().mean())
    n_el = 4 if group_type == "rotation" else 8
    assert set(metrics) == ({"test/group_acc", "test/acc"} | {f"test/acc_group_element_{i}" for i in range(n_el)}
                            | {f"test/acc_class_{i}" for i in range(10)})
    assert torch.allclose(metrics["test/group_acc"], torch.stack(accs).mean())
    for i in range(n_el):
        assert float(metrics[f"test/acc_group_element_{i}"]) == float(accs[i])
    assert float(metrics["test/acc"]) == float(accs[0])
    with pytest.raises(ValueError):
        inf.get_inference_metrics(torch.zeros(2, 3, 40, 40, device=dev), y[:2].to(dev))


# ---- N3: gradients with respect to the group element (straight-through training of the canonicalization network) ------
def _smooth(b, c, h, w, seed):
    """low-frequency images: the bilinear cell derivative is then close to the derivative of the underlying field"""
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    out = torch.zeros(b, c, h, w)
    for k in range(4):
        f = torch.rand(b, c, 2, generator=g) * 3 + 0.5
        ph = torch.rand(b, c, 2, generator=g) * 6.28
        out += torch.sin(6.28 * f[..., 0, None, None] * xs + ph[..., 0, None, None]) * torch.cos(6.28 * f[..., 1, None, None] * ys + ph[..., 1, None, None])
    return out / 4


@pytest.mark.parametrize("group_type,n", [("rotation", 8), ("roto-reflection", 8), ("rotation", 6), ("roto-reflection", 4),
                                           ("rotation", 4)])
def test_warp_element_gradients_vs_oracle_autograd(group_type, n, cuda_device):
    """eqb_warp_element_grad vs torch autograd through the fp64 oracle (kornia rotate restated + flip blend).
    Elements that are not quarter turns: 2e-4 relative.  Quarter turns sit exactly on the kinks of the bilinear
    interpolant (integral source coordinates): the reference lands on either side at random (its fp32 matrix is +-4e-8
    off the exact permutation) and returns a one-sided derivative; ours returns the symmetric one, checked against the
    mean of the oracle's gradients at angle +- 1e-3 degrees (the two one-sided values)."""
    ops = _mods()[0]
    dev = cuda_device
    reflect = group_type == "roto-reflection"
    G = n * (2 if reflect else 1)
    b, h, w = 2 * G, 40, 40
    idx = torch.arange(b) % G
    angles = torch.linspace(0.0, 360.0, n + 1)[:n][idx % n].double()
    quarter = (angles % 90 == 0)
    gen = torch.Generator().manual_seed(200 + n + reflect)

    def oracle(fn, delta):
        a = (angles + delta).clone().requires_grad_(True)
        r = (idx >= n).double().requires_grad_(True) if reflect else None
        fn(a, r).backward()
        return a.grad, (r.grad if reflect else None)

    def check(ours_rot, ours_ref, fn):
        g0, r0 = oracle(fn, 0.0)
        gp, _ = oracle(fn, 1e-3)
        gm, _ = oracle(fn, -1e-3)
        scale = g0.abs().max()
        assert ((ours_rot.double() - g0).abs()[~quarter] <= 2e-4 * scale).all()
        assert ((ours_rot.double() - 0.5 * (gp + gm)).abs()[quarter] <= 2e-3 * scale).all()
        if reflect:
            assert rel_err(ours_ref.double(), r0) < 2e-4
        else:
            assert ours_ref is None

    x = _smooth(b, 3, h, w, 300 + n)
    go = torch.randn(b, 3, h, w, generator=gen)
    o_rot, o_ref = ops.warp_element_grad(x.to(dev), go.to(dev), idx.to(dev).int(), n, reflect, 0)
    check(o_rot.cpu(), o_ref.cpu() if reflect else None,
          lambda a, r: (O.canonicalize_image(x.double(), a, r) * go.double()).sum())
    for mode, rep, c in ((1, "scalar", 3), (2, "regular", 2 * G)):
        f = _smooth(b, c, h, w, 400 + n + mode)
        gf = torch.randn(b, c, h, w, generator=gen)
        o_rot, o_ref = ops.warp_element_grad(f.to(dev), gf.to(dev), idx.to(dev).int(), n, reflect, mode)

        def inverted(a, r):
            if rep == "scalar":
                return O.invert_image_features(f.double(), a, r, n, G, "scalar")
            # regular: the channel roll is `.long()` of the angle (images/utils.py:28,67) and carries no gradient; it is
            # taken from the exact angles so that the +-1e-3 degree probes do not truncate to a different shift
            out = O.invert_image_features(f.double(), a, r, n, G, "scalar").reshape(b, c // G, G, h, w)
            shift = angles.float() / 360.0 * n
            if reflect:
                out = torch.cat([O.roll_by_gather(out[:, :, :n], shift), O.roll_by_gather(out[:, :, n:], -shift)], dim=2)
            else:
                out = O.roll_by_gather(out, shift)
            return out.reshape(b, c, h, w)

        assert torch.equal(inverted(angles, (idx >= n).double() if reflect else None),
                           O.invert_image_features(f.double(), angles, (idx >= n).double() if reflect else None, n, G, rep))
        check(o_rot.cpu(), o_ref.cpu() if reflect else None, lambda a, r: (inverted(a, r) * gf.double()).sum())


@pytest.mark.parametrize("group_type", ["rotation", "roto-reflection"])
def test_canonicalization_network_trains_through_straight_through_element(group_type, cuda_device):
    """The reference's training step (examples/images/classification/model.py:71-127) with a torch canonicalization
    network: task loss -> canonicalize warp -> straight-through one-hot -> network, plus the prior loss.  The gradient of
    every network parameter equals torch autograd through the oracle chain (same network, fp64)."""
    from equiadapt_b200.images.canonicalization.discrete_group import GroupEquivariantImageCanonicalization
    dev = cuda_device
    n, reflect = 8, group_type == "roto-reflection"
    G = n * (2 if reflect else 1)

    class Scorer(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.group_type, self.num_rotations = group_type, n
            self.conv = torch.nn.Conv2d(3, 6, 5)
            self.fc = torch.nn.Linear(6, G)

        def forward(self, x):
            return self.fc(torch.tanh(self.conv(x)).mean(dim=(2, 3)))

    torch.manual_seed(210 + reflect)
    net = Scorer()
    hp = SimpleNamespace(beta=1.0, input_crop_ratio=1.0, resize_shape=32)
    can = GroupEquivariantImageCanonicalization(net.to(dev), hp, (3, 32, 32)).train()
    x = _smooth(8, 3, 32, 32, 220)
    wtask = torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(221))
    xc = can(x.to(dev))
    assert xc.requires_grad
    loss = (xc * wtask.to(dev)).sum() + 100.0 * can.get_prior_regularization_loss()
    loss.backward()
    idx = can.canonicalization_info_dict["group_element"].index.cpu().long()
    ours = {k: p.grad.detach().cpu().double() for k, p in net.named_parameters()}

    import copy
    ref = copy.deepcopy(net).cpu().double()
    xd = x.double()
    angles = torch.linspace(0.0, 360.0, n + 1)[:n].double()
    comp = torch.cat([angles, angles]) if reflect else angles

    def reference_step(delta):
        """loss and parameter gradients of the reference chain, the rotation probed at +delta degrees"""
        for p in ref.parameters():
            p.grad = None
        act = ref(xd)                   # crop ratio 1, resize 32 -> the pre-network transform is the identity
        assert torch.equal(act.argmax(-1), idx)
        onehot = torch.nn.functional.one_hot(act.argmax(-1), G).double()
        soft = torch.softmax(hp.beta * act, -1)
        st = onehot + soft - soft.detach()                                      # basecanonicalization.py:239-251
        rot = (st * comp).sum(-1) + delta                                       # discrete_group.py:110-133
        refl = (st * torch.cat([torch.zeros(n), torch.ones(n)]).double()).sum(-1) if reflect else None
        yo = O.canonicalize_image(xd, rot, refl)
        lo = (yo * wtask.double()).sum() + 100.0 * torch.nn.functional.cross_entropy(act, torch.zeros(8, dtype=torch.long))
        lo.backward()
        return float(lo.detach()), {k: p.grad.clone() for k, p in ref.named_parameters()}

    lo, _ = reference_step(0.0)
    assert abs(float(loss.detach()) - lo) < 1e-3 * abs(lo)
    # quarter-turn samples sit on the kinks of the bilinear interpolant: the reference's gradient there is one of the
    # two one-sided values (chosen by fp32 noise), ours the symmetric one = the mean of the +-1e-3 degree probes;
    # for the other samples the two probes agree to O(delta)
    _, gp = reference_step(1e-3)
    _, gm = reference_step(-1e-3)
    for k in gp:
        assert rel_err(ours[k], 0.5 * (gp[k] + gm[k])) < 2e-3, k
    # and an optimiser step changes the activations
    opt = torch.optim.SGD(net.parameters(), lr=1e-3)
    opt.step()


# ---- N3: training path of CustomEquivariantNetwork (layer-wise forward + backward kernels) ------------------------------
@pytest.mark.parametrize("b,cin,h,w,n,k,relu", [(3, 3, 20, 24, 32, 5, True), (2, 70, 9, 11, 130, 1, False), (1, 5, 8, 8, 7, 3, True)])
def test_conv2d_forward_and_weight_grad_vs_torch(b, cin, h, w, n, k, relu, cuda_device):
    """eqb_conv2d_forward / eqb_conv2d_weight_grad / eqb_plane_sums vs torch fp64 (F.conv2d and its autograd: what the
    reference runs, custom_group_equivariant_layers.py:104-112).  fp32 tolerance 1e-5 relative."""
    ops = _mods()[0]
    dev = cuda_device
    g = torch.Generator().manual_seed(b * 100 + n)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(n, cin, k, k, generator=g) * 0.2
    bias = torch.randn(n, generator=g)
    y = ops.conv2d_forward(x.to(dev), wt.to(dev), bias.to(dev), relu)
    xd, wd = x.double(), wt.double().requires_grad_(True)
    yo = torch.nn.functional.conv2d(xd, wd, bias.double())
    want = torch.relu(yo) if relu else yo
    assert rel_err(y.cpu().double(), want.detach()) < 1e-5
    mask = torch.randn(y.shape, generator=g)
    ym = ops.conv2d_forward(x.to(dev), wt.to(dev), None, False, mask=mask.to(dev))
    assert rel_err(ym.cpu().double(), (torch.nn.functional.conv2d(xd, wd) * (mask > 0)).detach()) < 1e-5
    dy = torch.randn(yo.shape, generator=g)
    (yo * dy.double()).sum().backward()
    dw = ops.conv2d_weight_grad(dy.to(dev), x.to(dev), k)
    assert rel_err(dw.cpu().double(), wd.grad) < 1e-5
    assert rel_err(ops.plane_sums(dy.to(dev)).cpu().double(), dy.double().sum((2, 3))) < 1e-5


@pytest.mark.parametrize("b,h,w", [(2, 16, 16), (3, 20, 20), (2, 92, 92), (1, 4, 20)])
def test_pointwise_conv_tensor_core_path_vs_torch(b, h, w, cuda_device, monkeypatch):
    """eqb_conv2d_forward with k = 1, 256 -> 256 channels takes the CTA-pair tcgen05 kernel (csrc/gconv_stack_tc.cu,
    namespace pw: fp16 hi/lo operand split, per-image power-of-two scales, fp32 accumulation in TMEM): the regular layers of
    CustomEquivariantNetwork in train() and their data gradient (custom_group_equivariant_layers.py:298-334).  Against torch
    fp64 per image -- images of very different ranges share the batch -- at 4e-6 of the image's max |y| (the tensor core's
    fp32 accumulation truncates: ~1.5e-6 measured; the SIMT kernel gives ~3e-7), and against the SIMT kernel (EQB_TRAIN_TC=0)."""
    ops = _mods()[0]
    dev = cuda_device
    g = torch.Generator().manual_seed(b * 1000 + h)
    x = (torch.randn(b, 256, h, w, generator=g) * torch.logspace(-2, 2, b)[:, None, None, None]).to(dev)
    wt = (torch.randn(256, 256, 1, 1, generator=g) / 16).to(dev)
    bias = torch.randn(256, generator=g).to(dev)
    mask = torch.randn(b, 256, h, w, generator=g).to(dev)

    def check(y, want):
        scale = want.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30)
        return float(((y.double() - want).abs() / scale).max())

    lin = torch.einsum("nc,bchw->bnhw", wt[:, :, 0, 0].double(), x.double())
    want_f = (lin + bias.double()[None, :, None, None]).clamp_min(0)
    want_g = torch.where(mask > 0, lin, torch.zeros_like(lin))
    y_f = ops.conv2d_forward(x, wt, bias, True)
    y_g = ops.conv2d_forward(x, wt, None, False, mask=mask)
    assert check(y_f, want_f) < 4e-6 and check(y_g, want_g) < 4e-6
    monkeypatch.setenv("EQB_TRAIN_TC", "0")
    s_f = ops.conv2d_forward(x, wt, bias, True)
    s_g = ops.conv2d_forward(x, wt, None, False, mask=mask)
    assert check(s_f, want_f) < 1e-6 and check(s_g, want_g) < 1e-6
    assert not torch.equal(s_f, y_f)          # (two different kernels really ran)
    # masked-out and ReLU-clamped positions are exact zeros on both paths
    assert torch.equal(y_g == 0, (mask <= 0) | (y_g == 0)) and bool((y_g[mask <= 0] == 0).all())
    assert bool((y_f >= 0).all())


@pytest.mark.parametrize("b,cin,h,w,k", [(2, 3, 20, 24, 5), (3, 3, 96, 96, 5), (2, 7, 14, 12, 3), (2, 256, 16, 16, 1), (5, 256, 4, 20, 1)])
def test_tensor_core_training_convs_vs_torch(b, cin, h, w, k, cuda_device, monkeypatch):
    """The training GEMMs with 256 output channels on the tensor pipe (csrc/gconv_stack_tc.cu, namespaces pw / wg): forward of
    the 5x5 lift (patch taps gathered by the converters, cin*k*k <= 256) and of the 1x1 layers, and both weight gradients
    (1x1: dy and x by TMA; k x k: patch operand gathered, cin*k*k <= 128), against torch fp64 (F.conv2d and its autograd, what
    the reference runs: custom_group_equivariant_layers.py:104-112, :298-334) at 4e-6 of the result's max (1e-4 at the
    training shape's K = 541 696, tools/check_pw.py), and against the SIMT kernels.  Samples of very different ranges share the
    batch; chained calls hand the per-sample maxima on (eqb_conv2d_*_scaled) and must give the same bits as unchained ones."""
    ops = _mods()[0]
    dev = cuda_device
    g = torch.Generator().manual_seed(b * 1000 + h * 10 + k)
    x = (torch.randn(b, cin, h, w, generator=g) * torch.logspace(-1, 1, b)[:, None, None, None]).to(dev)
    wt = (torch.randn(256, cin, k, k, generator=g) / 8).to(dev)
    bias = torch.randn(256, generator=g).to(dev)
    dy = torch.randn(b, 256, h - k + 1, w - k + 1, generator=g).to(dev)
    wd = wt.double().requires_grad_(True)
    yo = torch.nn.functional.conv2d(x.double(), wd, bias.double())
    (yo * dy.double()).sum().backward()
    want_y = yo.detach().clamp_min(0)

    def err_y(y):
        scale = want_y.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30)
        return float(((y.double() - want_y).abs() / scale).max())

    def err_w(dw):
        return float((dw.double() - wd.grad).abs().max() / wd.grad.abs().max())

    y = ops.conv2d_forward(x, wt, bias, True)
    y0 = y.clone()
    dw = ops.conv2d_weight_grad(dy, x, k)
    assert err_y(y) < 4e-6 and err_w(dw) < 4e-6
    dw2, db = ops.conv2d_weight_grad(dy, x, k, with_bias_grad=True)       # bias gradient from the same pass over dy
    assert torch.allclose(dw2, dw, rtol=0, atol=4e-6 * float(dw.abs().max()))   # (split-K atomics: order varies)
    assert rel_err(db.cpu().double(), dy.double().sum((0, 2, 3)).cpu()) < 1e-5
    assert ops._known_amax(y) is not None and torch.equal(ops._known_amax(y), y.abs().amax(dim=(1, 2, 3)))
    # chained: the maxima recorded on y feed the next layer's operand scale -- same bits as a tensor without the record
    if cin == 256:
        y2a = ops.conv2d_forward(y, wt, None, False)
        y2b = ops.conv2d_forward(y.clone(), wt, None, False)
        assert torch.equal(y2a, y2b)
        y.mul_(2.0)                                   # an in-place write invalidates the record
        assert ops._known_amax(y) is None
    monkeypatch.setenv("EQB_TRAIN_TC", "0")
    ys = ops.conv2d_forward(x, wt, bias, True)
    dws, dbs = ops.conv2d_weight_grad(dy, x, k, with_bias_grad=True)
    assert err_y(ys) < 1e-6 and err_w(dws) < 2e-6
    assert rel_err(dbs.cpu().double(), dy.double().sum((0, 2, 3)).cpu()) < 1e-5
    assert not torch.equal(ys, y0)                    # (two different kernels really ran)


@pytest.mark.parametrize("n,reflect,k", [(4, False, 5), (8, False, 5), (8, True, 3), (6, True, 1)])
def test_filter_orbit_adjoints_vs_oracle_autograd(n, reflect, k, cuda_device):
    ops = _mods()[0]
    dev = cuda_device
    G = n * (2 if reflect else 1)
    g = torch.Generator().manual_seed(n + k)
    w = torch.randn(6, 3, k, k, generator=g).double().requires_grad_(True)
    orbit = O.lift_filter_orbit(w, n, reflect)
    d = torch.randn(orbit.shape, generator=g)
    (orbit * d.double()).sum().backward()
    assert rel_err(ops.lift_filter_orbit_adjoint(d.to(dev), 6, n, reflect).cpu().double(), w.grad) < 1e-5
    wr = torch.randn(4, 5, G, k, k, generator=g).double().requires_grad_(True)
    orbit = O.regular_filter_orbit(wr, n, reflect)
    d = torch.randn(orbit.shape, generator=g)
    (orbit * d.double()).sum().backward()
    assert rel_err(ops.regular_filter_orbit_adjoint(d.to(dev), 4, n, reflect).cpu().double(), wr.grad) < 1e-5


@pytest.mark.parametrize("group_type,n,layers,bias", [("rotation", 8, 3, True), ("roto-reflection", 4, 2, True),
                                                       ("rotation", 4, 1, False)])
def test_custom_network_training_gradients_vs_oracle_autograd(group_type, n, layers, bias, cuda_device):
    """CustomEquivariantNetwork in train(): activations and the gradient of every parameter equal torch autograd
    through the fp64 oracle network (filter orbits -> conv2d -> ReLU -> mean, custom_equivariant_networks.py:49-93)."""
    Net = _mods()[3]
    dev = cuda_device
    reflect = group_type == "roto-reflection"
    G = n * (2 if reflect else 1)
    torch.manual_seed(500 + n + layers)
    net = Net((3, 28, 28), 6, 5, group_type, n, layers, device="cpu").to(dev).train()
    mods = [m for m in net.eqv_network if hasattr(m, "weights")]
    with torch.no_grad():
        for m in mods:
            if bias:
                m.bias.normal_(0, 0.3)
    if not bias:
        for m in mods:
            m.bias = None
    x = torch.randn(5, 3, 28, 28, generator=torch.Generator().manual_seed(501))
    dact = torch.randn(5, G, generator=torch.Generator().manual_seed(502))
    act = net(x.to(dev))
    assert act.requires_grad
    (act * dact.to(dev)).sum().backward()
    ref_layers = [(m.weights.detach().cpu().double().requires_grad_(True),
                   m.bias.detach().cpu().double().requires_grad_(True) if bias else None) for m in mods]
    act_o = O.custom_equivariant_network(x.double(), ref_layers, n, reflect)
    (act_o * dact.double()).sum().backward()
    assert rel_err(act.detach().cpu().double(), act_o.detach()) < 1e-5
    for m, (wo, bo) in zip(mods, ref_layers):
        assert rel_err(m.weights.grad.cpu().double(), wo.grad) < 2e-5
        if bias:
            assert rel_err(m.bias.grad.cpu().double(), bo.grad) < 2e-5
    # eval() keeps using the fused inference stack and agrees with the training path's activations
    with torch.no_grad():
        assert rel_err(net.eval()(x.to(dev)).cpu().double(), act_o.detach()) < RTOL


def test_full_training_step_of_the_flagship_canonicalizer(cuda_device):
    """The reference's Lightning step (examples/images/classification/model.py:71-127) on this package's own network:
    canonicalize -> prediction network -> task loss + prior loss -> backward -> SGD.  Every parameter of the
    canonicalization network receives a finite gradient and the prior loss goes down over a few steps."""
    _, GEIC, _, Net = _mods()
    dev = cuda_device
    torch.manual_seed(510)
    net = Net((3, 32, 32), 4, 5, "rotation", 8, 3, device="cpu").to(dev)
    can = GEIC(net, SimpleNamespace(beta=1.0, input_crop_ratio=0.9, resize_shape=32), (3, 40, 40)).train()
    pred = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(4, 10)).to(dev)
    opt = torch.optim.SGD(list(can.parameters()) + list(pred.parameters()), lr=0.05)
    x = _smooth(16, 3, 40, 40, 511).to(dev)
    y = torch.randint(0, 10, (16,), generator=torch.Generator().manual_seed(512)).to(dev)
    priors = []
    for step in range(6):
        opt.zero_grad()
        logits = pred(can(x))
        prior = can.get_prior_regularization_loss()
        loss = torch.nn.functional.cross_entropy(logits, y) + 100.0 * prior
        loss.backward()
        for name, p in can.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), name
        assert any(float(p.grad.abs().sum()) > 0 for p in net.parameters())
        opt.step()
        priors.append(float(prior.detach()))
    assert priors[-1] < priors[0]


def test_training_step_is_cuda_graph_capturable(cuda_device):
    """Forward + losses + backward of the flagship canonicalizer (32 channels x C8 = 256: the tensor-core training kernels)
    captured as ONE CUDA graph on the stream it was warmed up on: the replay leaves the same loss and, within the run-to-run
    noise of the split-K atomics, the same parameter gradients as the eager step.  (The reference uploads its angle table on
    every call, discrete_group.py:110-117 -- a pageable copy no graph can contain; here the table stays on the device.)"""
    _, GEIC, _, Net = _mods()
    dev = cuda_device
    torch.manual_seed(520)
    net = Net((3, 32, 32), 32, 5, "rotation", 8, 3, device="cpu").to(dev)
    can = GEIC(net, SimpleNamespace(beta=1.0, input_crop_ratio=0.9, resize_shape=32), (3, 40, 40)).train()
    x = _smooth(6, 3, 40, 40, 521).to(dev)
    w = torch.randn(6, 3, 40, 40, generator=torch.Generator().manual_seed(522)).to(dev)
    params = [p for p in can.parameters() if p.requires_grad]

    def step():
        for p in params:
            p.grad = None
        loss = (can(x) * w).mean() + 100.0 * can.get_prior_regularization_loss()
        loss.backward()
        return loss

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            loss_eager = step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    g_eager = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
        loss_graph = (can(x) * w).mean() + 100.0 * can.get_prior_regularization_loss()
        loss_graph.backward()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize(dev)
    assert abs(float(loss_graph.detach()) - float(loss_eager.detach())) <= 1e-5 * abs(float(loss_eager.detach()))
    for p, ge in zip(params, g_eager):
        scale = float(ge.abs().max())
        if scale > 1e-4:                       # (a softmax-invariant bias has a mathematically zero gradient: noise only)
            assert float((p.grad - ge).abs().max()) < 1e-4 * scale
    del graph


# ---- N3: the optimisation-based variant trains any torch network through the cosine activations ----------------------
def test_cosine_activations_backward_vs_torch(cuda_device):
    ops = _mods()[0]
    dev = cuda_device
    gen = torch.Generator().manual_seed(600)
    b, G, v = 7, 8, 37
    vec = torch.randn(G * b, v, generator=gen)
    ref = torch.randn(1, v, generator=gen)
    dact = torch.randn(b, G, generator=gen)
    vo, ro = vec.double().requires_grad_(True), ref.double().requires_grad_(True)
    (O.cosine_group_activations(vo, ro, G) * dact.double()).sum().backward()
    vd, rd = vec.to(dev).requires_grad_(True), ref.to(dev).requires_grad_(True)
    act = ops.cosine_group_activations(vd, rd, G)
    (act * dact.to(dev)).sum().backward()
    assert rel_err(vd.grad.cpu().double(), vo.grad) < 1e-5
    assert rel_err(rd.grad.cpu().double(), ro.grad) < 1e-5
    # only the network output requires grad (learn_ref_vec = False)
    vd2 = vec.to(dev).requires_grad_(True)
    (ops.cosine_group_activations(vd2, ref.to(dev), G) * dact.to(dev)).sum().backward()
    assert rel_err(vd2.grad.cpu().double(), vo.grad) < 1e-5
    assert not ops.cosine_group_activations(vec.to(dev), ref.to(dev), G).requires_grad


@pytest.mark.parametrize("group_type", ["rotation", "roto-reflection"])
def test_optimized_variant_training_gradients_vs_oracle_chain(group_type, cuda_device):
    """OptimizedGroupEquivariantImageCanonicalization in train() around a small torch CNN (the reference's use: any
    non-equivariant network): task loss through the canonicalize warp + straight-through element, prior loss and the
    optimisation-specific loss.  Gradients of the CNN and of the reference vector equal torch autograd through the
    oracle chain (orbit expand -> CNN -> cosine similarity -> ..., discrete_group.py:387-512), the rotation probed at
    +-1e-3 degrees for the symmetric derivative at quarter turns."""
    _, _, OGEIC, _ = _mods()
    import copy
    dev = cuda_device
    n, reflect = 4, group_type == "roto-reflection"
    G = n * (2 if reflect else 1)

    class CNN(torch.nn.Module):
        out_vector_size = 12

        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 5, 5, stride=2)
            self.fc = torch.nn.Linear(5, 12)

        def forward(self, x):
            return self.fc(torch.tanh(self.conv(x)).mean(dim=(2, 3)))

    torch.manual_seed(610 + reflect)
    net = CNN()
    hp = SimpleNamespace(beta=1.0, input_crop_ratio=1.0, resize_shape=32, group_type=group_type, num_rotations=n,
                         artifact_err_wt=0, learn_ref_vec=True)
    can = OGEIC(copy.deepcopy(net).to(dev), hp, (3, 32, 32)).to(dev).train()
    ref_vec = can.reference_vector.detach().cpu().double().clone()
    x = _smooth(6, 3, 32, 32, 611)
    wtask = torch.randn(6, 3, 32, 32, generator=torch.Generator().manual_seed(612))
    xc = can(x.to(dev))
    loss = (xc * wtask.to(dev)).sum() + 10.0 * can.get_prior_regularization_loss() + 3.0 * can.get_optimization_specific_loss()
    loss.backward()
    idx = can.canonicalization_info_dict["group_element"].index.cpu().long()
    ours = {k: p.grad.detach().cpu().double() for k, p in can.canonicalization_network.named_parameters()}
    ours["reference_vector"] = can.reference_vector.grad.detach().cpu().double()

    ref_net = copy.deepcopy(net).double()
    rv = ref_vec.clone().requires_grad_(True)
    angles = torch.linspace(0.0, 360.0, n + 1)[:n].double()
    comp = torch.cat([angles, angles]) if reflect else angles
    xd = x.double()

    def reference_step(delta):
        for p in ref_net.parameters():
            p.grad = None
        rv.grad = None
        vec = ref_net(O.group_augment(xd, n, reflect, 32))
        act = O.cosine_group_activations(vec, rv, G)
        assert torch.equal(act.argmax(-1), idx)
        onehot = torch.nn.functional.one_hot(act.argmax(-1), G).double()
        soft = torch.softmax(hp.beta * act, -1)
        st = onehot + soft - soft.detach()
        rot = (st * comp).sum(-1) + delta
        refl = (st * torch.cat([torch.zeros(n), torch.ones(n)]).double()).sum(-1) if reflect else None
        lo = ((O.canonicalize_image(xd, rot, refl) * wtask.double()).sum()
              + 10.0 * torch.nn.functional.cross_entropy(act, torch.zeros(6, dtype=torch.long))
              + 3.0 * O.optimization_specific_loss(vec, G, 12))
        lo.backward()
        g = {k: p.grad.clone() for k, p in ref_net.named_parameters()}
        g["reference_vector"] = rv.grad.clone()
        return float(lo.detach()), g

    lo, _ = reference_step(0.0)
    assert abs(float(loss.detach()) - lo) < 1e-3 * abs(lo)
    _, gp = reference_step(1e-3)
    _, gm = reference_step(-1e-3)
    for k in gp:
        assert rel_err(ours[k], 0.5 * (gp[k] + gm[k])) < 2e-3, k


# ---- N3: the frame path (point clouds, n-body) is differentiable -------------------------------------------------------
@pytest.mark.parametrize("modified", [False, True])
def test_gram_schmidt_backward_vs_oracle_autograd(modified, cuda_device):
    ops = _mods()[0]
    gen = torch.Generator().manual_seed(700 + modified)
    v = torch.randn(33, 3, 3, generator=gen)
    dR = torch.randn(33, 3, 3, generator=gen)
    vo = v.double().requires_grad_(True)
    ((O.modified_gram_schmidt(vo) if modified else O.gram_schmidt(vo)) * dR.double()).sum().backward()
    vd = v.to(cuda_device).requires_grad_(True)
    R = ops.gram_schmidt3(vd, modified=modified)
    (R * dR.to(cuda_device)).sum().backward()
    assert rel_err(vd.grad.cpu().double(), vo.grad) < 2e-5
    assert not ops.gram_schmidt3(v.to(cuda_device), modified=modified).requires_grad


def test_frame_apply_backward_vs_oracle_autograd(cuda_device):
    ops = _mods()[0]
    dev = cuda_device
    gen = torch.Generator().manual_seed(710)
    # SO(3) on point clouds
    x, R, dy = torch.randn(5, 3, 1500, generator=gen), torch.randn(5, 3, 3, generator=gen), torch.randn(5, 3, 1500, generator=gen)
    xo, Ro = x.double().requires_grad_(True), R.double().requires_grad_(True)
    (O.so3_canonicalize(xo, Ro) * dy.double()).sum().backward()
    xd, Rd = x.to(dev).requires_grad_(True), R.to(dev).requires_grad_(True)
    (ops.so3_apply(xd, Rd) * dy.to(dev)).sum().backward()
    assert rel_err(xd.grad.cpu().double(), xo.grad) < 1e-5 and rel_err(Rd.grad.cpu().double(), Ro.grad) < 1e-5
    Rd2 = R.to(dev).requires_grad_(True)                      # only the frame requires grad (the training case)
    (ops.so3_apply(x.to(dev), Rd2) * dy.to(dev)).sum().backward()
    assert rel_err(Rd2.grad.cpu().double(), Ro.grad) < 1e-5
    # E(3) on particle rows
    m = 45
    loc, vel, t = (torch.randn(m, 3, generator=gen) for _ in range(3))
    Rm = torch.randn(m, 3, 3, generator=gen)
    gl, gv, gi = (torch.randn(m, 3, generator=gen) for _ in range(3))
    ref = [a.double().requires_grad_(True) for a in (loc, vel, Rm, t)]
    cl, cv = O.e3_canonicalize(*ref)
    ((cl * gl.double()).sum() + (cv * gv.double()).sum()).backward()
    ours = [a.to(dev).requires_grad_(True) for a in (loc, vel, Rm, t)]
    cl2, cv2 = ops.e3_apply(*ours)
    ((cl2 * gl.to(dev)).sum() + (cv2 * gv.to(dev)).sum()).backward()
    for a, b in zip(ours, ref):
        assert rel_err(a.grad.cpu().double(), b.grad) < 1e-5
    ref = [a.double().requires_grad_(True) for a in (loc, Rm, t)]
    (O.e3_invert(*ref) * gi.double()).sum().backward()
    ours = [a.to(dev).requires_grad_(True) for a in (loc, Rm, t)]
    (ops.e3_invert(*ours) * gi.to(dev)).sum().backward()
    for a, b in zip(ours, ref):
        assert rel_err(a.grad.cpu().double(), b.grad) < 1e-5


def test_pointcloud_canonicalizer_trains_a_torch_frame_network(cuda_device):
    """EquivariantPointcloudCanonicalization around a torch network: task loss through R x and the MSE prior reach the
    network's parameters; gradients equal torch autograd through the oracle chain (gram_schmidt -> bmm, MSE(R, I))."""
    from equiadapt_b200.pointcloud.canonicalization.continuous_group import EquivariantPointcloudCanonicalization
    import copy
    dev = cuda_device

    class FrameNet(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mix = torch.nn.Parameter(torch.randn(3, 6) * 0.5)

        def forward(self, x):                    # (B,3,N) -> (B,3,3): three equivariant vectors (rows)
            feats = torch.stack([x.mean(-1), (x * x.norm(dim=1, keepdim=True)).mean(-1), x[..., 0], x[..., 1], x[..., 2], x[..., 3]], 1)
            return torch.einsum("vk,bkc->bvc", self.mix, feats)

    torch.manual_seed(720)
    net = FrameNet()
    can = EquivariantPointcloudCanonicalization(copy.deepcopy(net).to(dev), SimpleNamespace()).train()
    x = torch.randn(6, 3, 200, generator=torch.Generator().manual_seed(721))
    wt = torch.randn(6, 3, 200, generator=torch.Generator().manual_seed(722))
    y = can(x.to(dev))
    loss = (y * wt.to(dev)).sum() + 5.0 * can.get_prior_regularization_loss()
    loss.backward()
    ref = copy.deepcopy(net).double()
    R = O.gram_schmidt(ref(x.double()))
    lo = (O.so3_canonicalize(x.double(), R) * wt.double()).sum() + 5.0 * torch.nn.functional.mse_loss(R, torch.eye(3).double().expand(6, 3, 3))
    lo.backward()
    assert abs(float(loss.detach()) - float(lo.detach())) < 1e-4 * abs(float(lo.detach()))
    assert rel_err(can.canonicalization_network.mix.grad.cpu().double(), ref.mix.grad) < 2e-5


# ---- N3: the continuous (SO(2) / O(2)) image warp is differentiable in the group element ------------------------------
@pytest.mark.parametrize("with_reflection,shape", [(False, (3, 40, 40)), (True, (3, 36, 36)), (False, (1, 28, 28))])
def test_continuous_warp_gradients_vs_oracle_autograd(with_reflection, shape, cuda_device):
    """d canonicalize / d rotation matrix (and / d reflection indicator) of ContinuousGroupImageCanonicalization vs torch
    autograd through the fp64 oracle (flip blend, pad, kornia warp_affine restated, crop: continuous_group.py:162-210)."""
    from unittest import mock
    from equiadapt_b200.images.canonicalization.continuous_group import ContinuousGroupImageCanonicalization
    dev = cuda_device
    gen = torch.Generator().manual_seed(800 + with_reflection + shape[0])
    b = 6
    x = _smooth(b, *shape, 801)
    wt = torch.randn(b, *shape, generator=gen)
    ang = torch.rand(b, generator=gen) * 2 * torch.pi
    rot = torch.stack([torch.stack([torch.cos(ang), torch.sin(ang)], 1), torch.stack([-torch.sin(ang), torch.cos(ang)], 1)], 1)
    refl = torch.randint(0, 2, (b, 1, 1, 1), generator=gen).float() if with_reflection else None
    # The rotation centre is an integer pixel, a FIXED POINT of the map: its source coordinate sits exactly on a kink of the
    # bilinear interpolant for every angle, and rounding noise alone picks the side (the reference's float32 and float64
    # autograd differ by up to 8e-3 of the gradient through that one pixel).  Its loss weight is zeroed; everywhere
    # else float32 and float64 agree to 2e-6.
    import math
    p_ = 0 if shape[0] == 1 else math.ceil(shape[-1] * 0.5)
    c_ = (shape[-1] + 2 * p_) // 2 - p_
    wt[:, :, c_, c_] = 0
    ro = rot.double().requires_grad_(True)
    fo = refl.double().requires_grad_(True) if with_reflection else None
    (O.canonicalize_image_continuous(x.double(), ro, fo) * wt.double()).sum().backward()
    # ours, through the class (the element comes from a patched get_groupelement as in the reference's own fixture)
    can = ContinuousGroupImageCanonicalization(torch.nn.Identity(), SimpleNamespace(input_crop_ratio=0.9, resize_shape=(16, 16)), shape)
    rd = rot.to(dev).requires_grad_(True)
    fd = refl.to(dev).requires_grad_(True) if with_reflection else None
    element = {"rotation": rd.clone()}
    if with_reflection:
        element["reflection"] = fd * 1.0
    with mock.patch.object(can, "get_groupelement", return_value=element):
        y = can.canonicalize(x.to(dev))
    assert y.requires_grad
    (y * wt.to(dev)).sum().backward()
    assert rel_err(y.detach().cpu().double(), O.canonicalize_image_continuous(x.double(), rot.double(), refl.double() if with_reflection else None)) < RTOL
    assert rel_err(rd.grad.cpu().double(), ro.grad) < RTOL
    if with_reflection:
        assert rel_err(fd.grad.cpu().double(), fo.grad) < RTOL


def test_steerable_canonicalizer_trains_a_torch_network(cuda_device):
    """SteerableImageCanonicalization in train() around a torch network with parameters: task loss through the warp and
    the MSE prior reach the parameters; gradients equal torch autograd through the oracle chain."""
    from equiadapt_b200.images.canonicalization.continuous_group import SteerableImageCanonicalization
    import copy
    dev = cuda_device

    class Net(torch.nn.Module):
        group_type = "rotation"

        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 4, 5)
            self.fc = torch.nn.Linear(4, 4)

        def forward(self, x):
            return self.fc(torch.tanh(self.conv(x)).mean(dim=(2, 3))).reshape(-1, 2, 2)

    torch.manual_seed(810)
    net = Net()
    can = SteerableImageCanonicalization(copy.deepcopy(net).to(dev), SimpleNamespace(input_crop_ratio=0.8, resize_shape=(32, 32)),
                                         (3, 64, 64)).train()
    x = _smooth(5, 3, 64, 64, 811)
    wt = torch.randn(5, 3, 64, 64, generator=torch.Generator().manual_seed(812))
    wt[:, :, 32, 32] = 0                   # the rotation's fixed-point pixel (see the test above)
    y = can(x.to(dev))
    loss = (y * wt.to(dev)).sum() + 50.0 * can.get_prior_regularization_loss()
    loss.backward()
    ref = copy.deepcopy(net)              # the reference chain in its own float32 arithmetic
    vec = ref(O.pre_network_transform(x, (3, 64, 64), 0.8, (32, 32)))[:, 0]
    v1 = vec / vec.norm(dim=1, keepdim=True)
    rot = torch.stack([v1, torch.stack([-v1[:, 1], v1[:, 0]], 1)], 1)
    neg = rot.clone()
    neg[:, [0, 1], [1, 0]] *= -1          # the representation the prior sees after the reference's in-place flip (:180)
    lo = (O.canonicalize_image_continuous(x, rot, None) * wt).sum() + 50.0 * O.prior_loss_continuous(neg)
    lo.backward()
    assert abs(float(loss.detach()) - float(lo.detach())) < 1e-3 * abs(float(lo.detach()))
    for (k, p), (_, q) in zip(can.canonicalization_network.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad.cpu().double(), q.grad.double()) < 1e-3, k


@pytest.mark.parametrize("tag", ["c8", "d4"])
def test_training_gradients_vs_unmodified_reference_golden(tag, cuda_device):
    """The flagship canonicalizer in train() on the golden weights: activations, prior loss and every parameter gradient of
    loss = 100 * prior + <act, w> equal what the UNMODIFIED reference produced with torch autograd on CPU
    (tests/golden/train_step_*.npz)."""
    g = load_golden("train_step_" + tag)
    dev = cuda_device
    can = build_canonicalizer(g, dev).train()
    can(g["x"].to(dev))
    act = can.canonicalization_info_dict["group_activations"]
    assert act.requires_grad
    prior = can.get_prior_regularization_loss()
    loss = 100.0 * prior + (act * g["wact"].to(dev)).sum()
    loss.backward()
    assert rel_err(act.detach().cpu(), g["act"]) < RTOL
    assert abs(float(prior.detach()) - float(g["prior"])) < 1e-5
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    mods = [m for m in can.canonicalization_network.eqv_network if hasattr(m, "weights")]
    for i, m in enumerate(mods):
        assert rel_err(m.weights.grad.cpu(), g[f"gw{2 * i}"]) < RTOL, f"weights of layer {i}"
        assert rel_err(m.bias.grad.cpu(), g[f"gb{2 * i}"]) < RTOL, f"bias of layer {i}"


@pytest.mark.parametrize("tag", ["c8", "d4"])
def test_optimized_variant_training_gradients_vs_unmodified_reference_golden(tag, cuda_device):
    """OptimizedGroupEquivariantImageCanonicalization in train(): gradients of 10 * prior + 3 * optimisation-specific loss
    w.r.t. the consumer network's output and the learnable reference vector equal the UNMODIFIED reference's
    (tests/golden/opt_train_step_*.npz; the network is replayed by its recorded output, as in the forward golden test)."""
    _, _, OGEIC, _ = _mods()
    g = load_golden("opt_train_step_" + tag)
    dev = cuda_device
    vec = g["vector_out"].to(dev).requires_grad_(True)

    class Net(torch.nn.Module):
        out_vector_size = vec.shape[1]

        def forward(self, xa):
            return vec

    hp = SimpleNamespace(beta=1.0, input_crop_ratio=g["crop_ratio"], resize_shape=int(g["resize"]), group_type=g["group_type"],
                         num_rotations=g["num_rotations"], artifact_err_wt=0, learn_ref_vec=True)
    can = OGEIC(Net(), hp, tuple(int(v) for v in g["in_shape"])).to(dev).train()
    with torch.no_grad():
        can.reference_vector.copy_(g["reference_vector"].to(dev))
    can(g["x"].to(dev))
    prior, opt = can.get_prior_regularization_loss(), can.get_optimization_specific_loss()
    loss = 10.0 * prior + 3.0 * opt
    loss.backward()
    assert rel_err(can.canonicalization_info_dict["group_activations"].detach().cpu(), g["act"]) < RTOL
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert rel_err(vec.grad.cpu(), g["g_vector_out"]) < RTOL
    assert rel_err(can.reference_vector.grad.cpu(), g["g_reference_vector"]) < RTOL


def test_frame_path_training_gradients_vs_unmodified_reference_golden(cuda_device):
    """Point-cloud and n-body canonicalizers in train(): gradients w.r.t. the frame network's outputs equal the UNMODIFIED
    reference's torch autograd (tests/golden/pointcloud_train.npz, nbody_train.npz)."""
    from equiadapt_b200.nbody.canonicalization.euclidean_group import EuclideanGroupNBody
    from equiadapt_b200.pointcloud.canonicalization.continuous_group import EquivariantPointcloudCanonicalization
    dev = cuda_device
    g = load_golden("pointcloud_train")
    vecs = g["vectors"].to(dev).requires_grad_(True)

    class PNet(torch.nn.Module):
        def forward(self, _x):
            return vecs * 1.0

    can = EquivariantPointcloudCanonicalization(PNet(), SimpleNamespace()).train()
    loss = (can(g["x"].to(dev)) * g["w"].to(dev)).sum() + 5.0 * can.get_prior_regularization_loss()
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert rel_err(vecs.grad.cpu(), g["g_vectors"]) < RTOL

    g = load_golden("nbody_train")
    rv, t = g["rot_vectors"].to(dev).requires_grad_(True), g["translation"].to(dev).requires_grad_(True)

    class NNet(torch.nn.Module):
        def forward(self, *a):
            return rv * 1.0, t * 1.0

    can = EuclideanGroupNBody(NNet()).train()
    loc, vel = g["loc"].to(dev), g["vel"].to(dev)
    nodes = torch.sqrt(torch.sum(vel ** 2, dim=1)).unsqueeze(1)
    cl, cv = can(nodes, None, loc=loc, edges=None, vel=vel, edge_attr=None, charges=None)
    inv = can.invert_canonicalization(g["pred"].to(dev))
    loss = (cl * g["wl"].to(dev)).sum() + (cv * g["wv"].to(dev)).sum() + (inv * g["wi"].to(dev)).sum()
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert rel_err(rv.grad.cpu(), g["g_rot_vectors"]) < RTOL
    assert rel_err(t.grad.cpu(), g["g_translation"]) < RTOL


# ---- round 2: captured step, cache invalidation, second device, sampled element ----------------------------------------
def _small_c8(dev, seed=70, in_hw=64, resize=32, cout=8):
    _, GEIC, _, Net = _mods()
    torch.manual_seed(seed)
    net = Net((3, resize, resize), cout, 5, "rotation", 8, 3, device="cpu").to(dev)
    return GEIC(net, SimpleNamespace(beta=1.0, input_crop_ratio=0.8, resize_shape=resize), (3, in_hw, in_hw)).eval()


def test_capture_step_replays_the_public_step_bit_exactly(cuda_device):
    """canonicalizer.capture_step(): ONE graph launch == canonicalize + invert + prior loss + identity metric of the eager
    public calls, on new data copied into the captured buffer and on the buffer itself."""
    dev = cuda_device
    can = _small_c8(dev)
    gen = torch.Generator().manual_seed(71)
    x0, x1 = torch.rand(16, 3, 64, 64, generator=gen).to(dev), torch.rand(16, 3, 64, 64, generator=gen).to(dev)
    step = can.capture_step(x0.clone(), induced_rep_type="scalar")
    with torch.no_grad():
        for x in (x1, x0):
            y, z, loss, ident = step(x)
            torch.cuda.synchronize()
            got = (y.clone(), z.clone(), float(loss), float(ident))
            yr = can(x)
            zr = can.invert_canonicalization(yr, induced_rep_type="scalar")
            ref = (yr, zr, float(can.get_prior_regularization_loss()), float(can.get_identity_metric()))
            assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]) and got[2:] == ref[2:]
        step.inputs[0].copy_(x1)
        y, z, loss, ident = step()                 # no argument: replay on whatever the captured buffer holds
        torch.cuda.synchronize()
        assert torch.equal(y, can(x1))
    with pytest.raises(ValueError):
        step(x1[:8])


def test_packed_operand_cache_follows_parameter_writes(cuda_device):
    """in-place ops bump the version and repack; writes through .data do not, so invalidate_packed() (also
    called by load_state_dict / _apply / train) or cache_packed = False is the contract for those."""
    dev = cuda_device
    can = _small_c8(dev, seed=72)
    net = can.canonicalization_network
    x = torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(73)).to(dev)
    lift = net.eqv_network[0]
    with torch.no_grad():
        a0 = net(x).clone()
        lift.weights.mul_(1.5)                              # in-place op on the parameter: version bump -> repacked
        a1 = net(x).clone()
        assert not torch.equal(a0, a1)
        lift.weights.data.mul_(2.0)                         # write through .data: invisible to the version counter
        stale = net(x).clone()
        assert torch.equal(stale, a1)
        net.invalidate_packed()
        a2 = net(x).clone()
        assert not torch.equal(a2, a1)
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        lift.weights.data.mul_(0.5)
        net.load_state_dict(sd)                             # restores the doubled weights AND drops the cache
        assert torch.equal(net(x), a2)
        net.cache_packed = False                            # the reference's behaviour: rebuilt on every forward
        lift.weights.data.mul_(0.5)
        assert torch.equal(net(x), a1)


def test_second_device_in_one_process(cuda_device):
    """function attributes / the SM count / the stall-report symbol are per DEVICE; the stack, the TMA warp and the
    frame kernels must run on cuda:1 after cuda:0 in one process."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    outs = []
    x = torch.rand(8, 3, 64, 64, generator=torch.Generator().manual_seed(75))
    for d in (0, 1):
        dev = torch.device("cuda", d)
        can = _small_c8(dev, seed=74)
        with torch.no_grad():
            y = can(x.to(dev))
            z = can.invert_canonicalization(y, induced_rep_type="scalar")
            outs.append((y.cpu(), z.cpu(), float(can.get_prior_regularization_loss())))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]


def test_gumbel_softmax_element_is_the_sampled_one(cuda_device):
    """with gradient_trick = "gumbel_softmax" the reference derives the element from the SAMPLED one-hot
    (discrete_group.py:94-135); the warp must be driven by that same sample, not by the arg-max."""
    dev = cuda_device
    can = _small_c8(dev, seed=76)
    can.gradient_trick = "gumbel_softmax"
    ops = _mods()[0]
    x = torch.rand(32, 3, 64, 64, generator=torch.Generator().manual_seed(77)).to(dev)
    with torch.no_grad():
        torch.manual_seed(5)
        y = can(x)
        el = can.canonicalization_info_dict["group_element"]
        angles = torch.linspace(0.0, 360.0, 9)[:8].to(dev)
        assert torch.equal(el["rotation"], angles[el.index.long()])
        assert torch.equal(y, ops.warp_canonicalize(x, el.index, 8, False))
        act = can.canonicalization_info_dict["group_activations"]
        assert not torch.equal(el.index.long(), act.argmax(-1))      # 32 samples with gaps of 1e-6: the sample differs somewhere


def test_input_shape_must_match_in_shape(cuda_device):
    can = _small_c8(cuda_device, seed=78)
    with pytest.raises(ValueError):
        can(torch.rand(2, 3, 48, 48, device=cuda_device))


def test_stack_activations_do_not_depend_on_batch_mates(cuda_device):
    """the fp16 operand split is scaled PER IMAGE, so a sample's activations are bit-identical whatever
    shares its batch, and an image 1e6 times fainter than its neighbour keeps full relative accuracy (with one scale per
    call it lost its low-order bits).  Pair kernel shape (N = 256) and single-CTA shape (N = 64)."""
    _, _, _, Net = _mods()
    dev = cuda_device
    for cout, n_rot in ((32, 8), (16, 4)):
        torch.manual_seed(80 + cout)
        net = Net((3, 40, 40), cout, 5, "rotation", n_rot, 3, device="cpu")
        layers = [(m.weights.detach().clone(), m.bias.detach().clone()) for m in net.eqv_network if hasattr(m, "weights")]
        net = net.to(dev).eval()
        g = torch.Generator().manual_seed(81)
        x = torch.rand(6, 3, 40, 40, generator=g)
        scales = torch.tensor([1.0, 1e-6, 1e4, 3e-3, 1.0, 77.0]).view(-1, 1, 1, 1)
        xa = (x * scales).to(dev)
        xb = xa.clone()
        xb[2] = xb[2] * 1e-4            # change ONE image: every other row must not move by a bit
        xb[5] = 0.0
        with torch.no_grad():
            a, b = net(xa), net(xb)
            keep = [0, 1, 3, 4]
            assert torch.equal(a[keep], b[keep])
            assert bool(torch.isfinite(b).all())                       # an all-zero image scales by 1
            ref = O.custom_equivariant_network(xa.cpu().double(), [(w.double(), bb.double()) for w, bb in layers], n_rot, False)
            err = (a.cpu().double() - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)
            assert float(err.max()) < 2e-6, err
            single = torch.cat([net(xa[i:i + 1]) for i in range(6)])
            assert rel_err(single.cpu(), a.cpu()) < 1e-6


def test_tma_resize_equals_scalar_kernel_and_leaves_per_image_absmax(cuda_device, monkeypatch):
    """The TMA-staged crop + resize returns the same bits as the scalar kernel (same taps, same FMA order), for aligned and
    unaligned crop offsets, and eqb_crop_resize_aa_absmax leaves max |y[b]| per image."""
    ops = _mods()[0]
    dev = cuda_device
    g = torch.Generator().manual_seed(82)
    for (c, h, w, top, left, ch, cw, oh, ow) in ((3, 224, 224, 22, 22, 180, 180, 96, 96), (3, 64, 64, 3, 3, 58, 58, 32, 32),
                                                 (1, 40, 48, 0, 4, 40, 40, 33, 17), (2, 32, 32, 2, 2, 29, 29, 32, 32)):
        x = (torch.randn(5, c, h, w, generator=g) * torch.tensor([1.0, 1e-3, 50.0, 1.0, 0.0]).view(-1, 1, 1, 1)).to(dev)
        y = ops.crop_resize_aa(x, top, left, ch, cw, oh, ow, with_absmax=True)
        monkeypatch.setenv("EQB_RESIZE_SCALAR", "1")
        y_ref = ops.crop_resize_aa(x, top, left, ch, cw, oh, ow, with_absmax=True)
        monkeypatch.delenv("EQB_RESIZE_SCALAR")
        assert torch.equal(y, y_ref)
        want = y_ref.abs().amax(dim=(1, 2, 3))
        assert torch.equal(y._eqb_absmax[:5], want) and torch.equal(y_ref._eqb_absmax[:5], want)


def test_select_fused_into_the_finish_kernel_equals_the_select_kernel(cuda_device):
    """eqb_gconv_stack_run_select == eqb_gconv_stack_run + eqb_group_pool_select, bit for bit (index, angle, reflection,
    one-hot, the 5-float statistic), for C8 and D4, on consecutive calls (the last-block ticket resets itself), and the
    canonicalizer picks the fused selection up without launching the select kernel."""
    ops, GEIC, _, Net = _mods()
    dev = cuda_device
    for group_type, n_rot, cout in (("rotation", 8, 32), ("roto-reflection", 4, 8)):
        torch.manual_seed(90)
        net = Net((3, 40, 40), cout, 5, group_type, n_rot, 3, device="cpu").to(dev).eval()
        reflect = group_type == "roto-reflection"
        for trial, b in enumerate((37, 5, 37)):
            x = torch.rand(b, 3, 40, 40, generator=torch.Generator().manual_seed(91 + trial)).to(dev)
            with torch.no_grad():
                net.fuse_select = True
                act = net(x)
                fused = act._eqb_selection
                net.fuse_select = False
                act_plain = net(x)
                assert not hasattr(act_plain, "_eqb_selection")
                assert torch.equal(act, act_plain)
                idx, rot, refl, onehot, stats = ops.group_pool_select(act_plain, n_rot, reflect)
            assert fused[0] == n_rot and fused[1] == reflect
            assert torch.equal(fused[2], idx) and torch.equal(fused[3], rot) and torch.equal(fused[5], onehot)
            assert (fused[4] is None and refl is None) or torch.equal(fused[4], refl)
            assert torch.equal(fused[6], stats), (fused[6], stats)
        net.fuse_select = True
        can = GEIC(net, SimpleNamespace(beta=1.0, input_crop_ratio=0.8, resize_shape=40), (3, 64, 64)).eval()
        xs = torch.rand(9, 3, 64, 64, generator=torch.Generator().manual_seed(95)).to(dev)
        with torch.no_grad():
            can(xs)                                # (eval() dropped the packed operands: the first call re-packs)
            n0 = ops.launch_count
            can(xs)
            used = ops.launch_count - n0           # crop + resize (1), stack + finish/select (2), warp (1)
            assert used == 4, used
            want = ops.group_pool_select(can.canonicalization_info_dict["group_activations"], n_rot, reflect)
            assert torch.equal(can.canonicalization_info_dict["group_element"].index, want[0])
            assert float(can.get_prior_regularization_loss()) == float(want[4][3])


def test_escnn_stack_activations_do_not_depend_on_batch_mates(cuda_device):
    """a7: the e2cnn-style stack chains its fp16 operand scales per IMAGE as well (layer records per image): rows are
    bit-identical whatever shares the batch, and a faint image keeps its relative accuracy next to a bright one."""
    from equiadapt_b200.images.canonicalization_networks.escnn_networks import ESCNNEquivariantNetwork
    dev = cuda_device
    torch.manual_seed(96)
    net = ESCNNEquivariantNetwork((3, 40, 40), 8, 5, "rotation", 4, 3, device=str(dev)).eval()
    x = torch.rand(5, 3, 40, 40, generator=torch.Generator().manual_seed(97))
    xa = (x * torch.tensor([1.0, 1e-5, 300.0, 2e-2, 1.0]).view(-1, 1, 1, 1)).to(dev)
    xb = xa.clone()
    xb[2] = xb[2] * 1e-3
    with torch.no_grad():
        a, b = net(xa), net(xb)
        assert torch.equal(a[[0, 1, 3, 4]], b[[0, 1, 3, 4]])
        scales, shifts = net.folded_affine()
        ref = O.expanded_conv_network(xa.cpu().double(), [f.detach().cpu().double() for f in net.filters],
                                      [bb.detach().cpu().double() for bb in net.biases],
                                      [t.cpu().double() for t in scales], [t.cpu().double() for t in shifts], 4)
    err = (a.cpu().double() - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)
    assert float(err.max()) < 1e-4, err
