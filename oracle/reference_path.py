"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the reference's per-batch canonicalization hot path (SURVEY.md section 8a,
rows a3..a17).  Every function cites the reference file:line it follows (paths relative to
/root/reference).  The reference is pure PyTorch, so the restatement is written on torch CPU ops
in the reference's own op order: in float32 it reproduces the reference's rounding, in float64
(`dtype=torch.float64`) it is the tie-breaker for near-tied group activations.

How it is pinned (tests/test_oracle_*.py, all `-m "not gpu"`):
  * the reference's one known-answer test: gram_schmidt(randn(1,3,3), seed 0)[0,0,0] == 0.5740
    (tests/common/test_utils.py:6-12);
  * golden vectors in tests/golden/*.npz, produced by oracle/make_golden.py, which imports the
    UNMODIFIED reference package from /root/reference (through the stand-in modules of
    oracle/shims/) and records its inputs and outputs;
  * torchvision (installed) for CenterCrop / Pad / Resize and torch for conv2d / grid_sample, which
    are the reference's own callees.
kornia's part is restated in oracle/kornia_restated.py (parity unpinned at that boundary, see its
header); e2cnn networks cannot be run at all (parity unpinned).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import kornia_restated as K

# --------------------------------------------------------------------------------------------
# a2 / a3  pre-network transform: CenterCrop + antialiased bilinear Resize
# --------------------------------------------------------------------------------------------


def center_crop_offsets(h: int, w: int, ch: int, cw: int) -> Tuple[int, int]:
    """torchvision.transforms.functional.center_crop offsets (Python banker's rounding).

    Reference use: equiadapt/images/canonicalization/discrete_group.py:67-86 (CenterCrop objects).
    """
    top = int(round((h - ch) / 2.0))
    left = int(round((w - cw) / 2.0))
    return top, left


def canonization_crop_size(in_shape: Sequence[int], input_crop_ratio: float) -> Tuple[int, int]:
    """discrete_group.py:76-85: ceil(H*ratio), ceil(W*ratio)."""
    return math.ceil(in_shape[-2] * input_crop_ratio), math.ceil(in_shape[-1] * input_crop_ratio)


def center_crop(x: torch.Tensor, ch: int, cw: int) -> torch.Tensor:
    h, w = x.shape[-2:]
    assert ch <= h and cw <= w, "oracle covers crop <= image (the reference never pads here)"
    top, left = center_crop_offsets(h, w, ch, cw)
    return x[..., top:top + ch, left:left + cw]


def resize_output_size(h: int, w: int, size) -> Tuple[int, int]:
    """torchvision Resize(size): int -> smaller edge matched (long edge int(size*long/short)); pair -> as is."""
    if isinstance(size, int):
        short, long = (w, h) if w <= h else (h, w)
        new_short, new_long = size, int(size * long / short)
        return (new_long, new_short) if w <= h else (new_short, new_long)
    if len(size) == 1:
        return resize_output_size(h, w, int(size[0]))
    return int(size[0]), int(size[1])


def aa_axis_weights(in_size: int, out_size: int, dtype=torch.float64) -> Tuple[List[int], List[torch.Tensor]]:
    """ATen `_upsample_bilinear2d_aa` separable weights for one axis (align_corners=False).

    aten/src/ATen/native/cpu/UpSampleKernel.cpp `HelperInterpLinear::aa_filter` /
    `_compute_indices_min_size_weights_aa`.  Returns per output index the first
    source index and the normalised weights.
    """
    scale = in_size / out_size
    support = scale if scale >= 1.0 else 1.0
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    starts, weights = [], []
    for i in range(out_size):
        center = scale * (i + 0.5)
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        ws = []
        for j in range(lo, hi):
            v = 1.0 - abs((j - center + 0.5) * invscale)
            ws.append(max(v, 0.0))
        t = torch.tensor(ws, dtype=dtype)
        t = t / t.sum()
        starts.append(lo)
        weights.append(t)
    return starts, weights


def aa_resize_restated(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """Closed-form separable antialiased bilinear resize (what the CUDA kernel implements)."""
    b, c, h, w = x.shape
    sx, wx = aa_axis_weights(w, out_w, x.dtype)
    sy, wy = aa_axis_weights(h, out_h, x.dtype)
    tmp = x.new_zeros(b, c, h, out_w)
    for i in range(out_w):
        tmp[..., i] = (x[..., sx[i]:sx[i] + len(wx[i])] * wx[i]).sum(-1)
    out = x.new_zeros(b, c, out_h, out_w)
    for i in range(out_h):
        out[..., i, :] = (tmp[..., sy[i]:sy[i] + len(wy[i]), :] * wy[i][:, None]).sum(-2)
    return out


def pre_network_transform(x: torch.Tensor, in_shape: Sequence[int], input_crop_ratio: float, resize_shape) -> torch.Tensor:
    """discrete_group.py:174-188 with the transforms built at :60-92.

    Grayscale (C==1) inputs skip both (Identity).  Resize == F.interpolate(bilinear, antialias=True,
    align_corners=False), which is what torchvision.transforms.Resize calls on tensors.
    """
    if in_shape[0] == 1:
        return x
    ch, cw = canonization_crop_size(in_shape, input_crop_ratio)
    x = center_crop(x, ch, cw)
    oh, ow = resize_output_size(x.shape[-2], x.shape[-1], resize_shape)
    return F.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False, antialias=True)


# --------------------------------------------------------------------------------------------
# a4 / a5  filter orbits of the hand-rolled group convolutions
# --------------------------------------------------------------------------------------------


def _angles(num_rotations: int, dtype, n: Optional[int] = None) -> torch.Tensor:
    # linspace is always produced in float32 by the reference, then used at the tensor's dtype
    a = torch.linspace(0.0, 360.0, steps=num_rotations + 1, dtype=torch.float32)[:num_rotations]
    return a.to(dtype)


def lift_filter_orbit(weights: torch.Tensor, num_rotations: int, reflect: bool) -> torch.Tensor:
    """custom_group_equivariant_layers.py:62-90 (C_N) and :169-199 (D_N).

    weights (Cout,Cin,k,k) -> (Cout*|G|, Cin, k, k), channel = o*|G| + g.
    """
    cout, cin, k, _ = weights.shape
    w = weights.flatten(0, 1).unsqueeze(0).repeat(num_rotations, 1, 1, 1)
    rot = K.rotate(w, _angles(num_rotations, weights.dtype))
    if reflect:
        rot = torch.cat([rot, K.hflip(rot)], dim=0)
    g = rot.shape[0]
    return rot.reshape(g, cout, cin, k, k).transpose(0, 1).flatten(0, 1)


def regular_permutation_indices(num_rotations: int, reflect: bool) -> torch.Tensor:
    """Index table idx[g,h] = which input-group slice of W feeds (g,h).

    C_N: custom_group_equivariant_layers.py:283-293;  D_N: :420-449.
    """
    n = num_rotations
    ar = torch.arange(n)
    fwd = (ar[None, :] - ar[:, None]) % n  # [g,h] = (h-g) mod n
    if not reflect:
        return fwd
    inv = (ar[None, :] + ar[:, None]) % n
    upper = torch.cat([fwd, inv + n], dim=1)
    lower = torch.cat([inv + n, fwd], dim=1)
    return torch.cat([upper, lower], dim=0)


def regular_filter_orbit(weights: torch.Tensor, num_rotations: int, reflect: bool) -> torch.Tensor:
    """custom_group_equivariant_layers.py:298-334 (C_N) and :461-507 (D_N).

    weights (Cout,Cin,|G|,k,k) -> (Cout*|G|, Cin*|G|, k, k).
    """
    cout, cin, g, k, _ = weights.shape
    idx = regular_permutation_indices(num_rotations, reflect)  # (G,G)
    w = weights.flatten(0, 1).unsqueeze(0).repeat(g, 1, 1, 1, 1)  # (G, Cout*Cin, G, k, k)
    gather_idx = idx[:, None, :, None, None].expand(g, cout * cin, g, k, k)
    perm = torch.gather(w, 2, gather_idx)
    ang = _angles(num_rotations, weights.dtype)
    if reflect:
        ang = torch.cat([ang, ang])
    rot = K.rotate(perm.flatten(1, 2), ang)
    if reflect:
        rot = torch.cat([rot[:num_rotations], K.hflip(rot[num_rotations:])])
    return rot.reshape(g, cout, cin, g, k, k).transpose(0, 1).reshape(cout * g, cin * g, k, k)


def custom_equivariant_network(x: torch.Tensor, layers: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]],
                               num_rotations: int, reflect: bool, return_feature_map: bool = False) -> torch.Tensor:
    """CustomEquivariantNetwork.forward, custom_equivariant_networks.py:49-93.

    `layers` = [(W_lift (Cout,Cin,k,k), b), (W (Cout,Cout,|G|,1,1), b), ...]; ReLU between layers
    (:56,:70), mean over (channel, H, W) at the end (:91).
    """
    g = num_rotations * (2 if reflect else 1)
    b = x.shape[0]
    for li, (w, bias) in enumerate(layers):
        if li == 0:
            wg = lift_filter_orbit(w, num_rotations, reflect)
        else:
            x = torch.relu(x).flatten(1, 2)
            wg = regular_filter_orbit(w, num_rotations, reflect)
        x = F.conv2d(x, wg)
        x = x.reshape(b, w.shape[0], g, x.shape[2], x.shape[3])
        if bias is not None:
            x = x + bias[None, :, None, None, None]
    if return_feature_map:
        return x
    return torch.mean(x, dim=(1, 3, 4))


def expanded_conv_network(x: torch.Tensor, filters: Sequence[torch.Tensor], biases: Sequence[Optional[torch.Tensor]],
                          scales: Sequence[Optional[torch.Tensor]], shifts: Sequence[Optional[torch.Tensor]],
                          num_group: int) -> torch.Tensor:
    """ESCNNEquivariantNetwork.forward in eval(), escnn_networks.py:93-117 (modules :66-91), on the EXPANDED
    tensors e2cnn caches: R2Conv = conv2d(filter, expanded_bias), valid padding, stride 1; InnerBatchNorm (eval) =
    per-channel affine, equal within a field; ReLU; PointwiseDropout = identity; then
    reshape (B, Cout, |G|, H', W') and mean over (1, 3, 4) (:107-115).  e2cnn itself is absent here: the basis
    expansion that produces `filter` is NOT restated (parity unpinned); this restates what follows it.
    """
    n_layers = len(filters)
    for li, w in enumerate(filters):
        x = F.conv2d(x, w, None if biases[li] is None else biases[li])
        if li < n_layers - 1:
            if scales[li] is not None:
                x = x * scales[li][None, :, None, None]
            if shifts[li] is not None:
                x = x + shifts[li][None, :, None, None]
            x = torch.relu(x)
    b, n = x.shape[0], x.shape[1]
    x = x.reshape(b, n // num_group, num_group, x.shape[2], x.shape[3])
    return torch.mean(x, dim=(1, 3, 4))


# --------------------------------------------------------------------------------------------
# a9  activations -> group element
# --------------------------------------------------------------------------------------------


def activations_to_onehot(act: torch.Tensor, num_group: int, beta: float, training: bool = False) -> torch.Tensor:
    """common/basecanonicalization.py:221-256 (straight_through trick)."""
    onehot = F.one_hot(torch.argmax(act, dim=-1), num_group).float().to(act.dtype)
    if training:
        soft = F.softmax(beta * act, dim=-1)
        return onehot + soft - soft.detach()
    return onehot


def activations_to_group_element(act: torch.Tensor, num_rotations: int, reflect: bool, beta: float = 1.0,
                                 training: bool = False) -> Dict[str, torch.Tensor]:
    """discrete_group.py:94-135: rotation (degrees) and reflection (0/1) from the one-hot."""
    g = num_rotations * (2 if reflect else 1)
    onehot = activations_to_onehot(act, g, beta, training)
    angles = torch.linspace(0.0, 360.0, num_rotations + 1)[:num_rotations].to(act.dtype)
    rot_comp = torch.cat([angles, angles]) if reflect else angles
    out = {"rotation": torch.sum(onehot * rot_comp, dim=-1)}
    if reflect:
        ident = torch.cat([torch.zeros(num_rotations), torch.ones(num_rotations)]).to(act.dtype)
        out["reflection"] = torch.sum(onehot * ident, dim=-1)
    return out


# --------------------------------------------------------------------------------------------
# a10  canonicalize (pad -> flip blend -> rotate(-angle) -> crop)
# --------------------------------------------------------------------------------------------


def canonicalize_image(x: torch.Tensor, rotation_deg: torch.Tensor, reflection: Optional[torch.Tensor]) -> torch.Tensor:
    """discrete_group.py:204-215 with pad/crop from :62-71 (Identity for 1-channel inputs)."""
    gray = x.shape[1] == 1
    h, w = x.shape[-2:]
    if not gray:
        p = math.ceil(w * 0.5)
        x = F.pad(x, (p, p, p, p), mode="replicate")
    if reflection is not None:
        r = reflection[:, None, None, None]
        x = (1 - r) * x + r * K.hflip(x)
    x = K.rotate(x, -rotation_deg)
    if not gray:
        x = center_crop(x, h, w)
    return x


def rotate_closed_form(x: torch.Tensor, angle_deg: torch.Tensor, clamp: bool) -> torch.Tensor:
    """what rotate() amounts to, in float64, with exact coefficients for
    multiples of 90 degrees.  `clamp=True` reproduces pad(replicate)+rotate+crop, `False` the bare
    zero-fill rotate.  Used to bound how far the fp32 reference is from the exact map."""
    b, c, h, w = x.shape
    xd = x.double()
    out = torch.zeros_like(xd)
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing="ij")
    for i in range(b):
        a = float(angle_deg[i]) % 360.0
        q, rem = divmod(a, 90.0)
        if rem == 0.0:
            cs, sn = [(1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)][int(q) % 4]
        else:
            cs, sn = math.cos(math.radians(a)), math.sin(math.radians(a))
        sxs = cx + cs * (xs - cx) - sn * (ys - cy)
        sys_ = cy + sn * (xs - cx) + cs * (ys - cy)
        if clamp:
            sxs = sxs.clamp(0, w - 1)
            sys_ = sys_.clamp(0, h - 1)
        x0 = torch.floor(sxs)
        y0 = torch.floor(sys_)
        fx = sxs - x0
        fy = sys_ - y0
        acc = torch.zeros(c, h, w, dtype=torch.float64)
        for dy, wy in ((0, 1 - fy), (1, fy)):
            for dx, wx in ((0, 1 - fx), (1, fx)):
                xi = (x0 + dx).long()
                yi = (y0 + dy).long()
                ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
                v = xd[i][:, yi.clamp(0, h - 1), xi.clamp(0, w - 1)]
                acc += v * (wy * wx * ok)[None]
        out[i] = acc
    return out


# --------------------------------------------------------------------------------------------
# a11  invert_canonicalization on feature maps
# --------------------------------------------------------------------------------------------


def roll_by_gather(feature_map: torch.Tensor, shifts: torch.Tensor) -> torch.Tensor:
    """images/utils.py:8-29: out[:,:,g] = in[:,:,(g - long(shift)) mod G]."""
    b, c, g, h, w = feature_map.shape
    ar = torch.arange(g).view(1, 1, g, 1, 1).repeat(b, c, 1, h, w)
    idx = (ar - shifts[:, None, None, None, None].long()) % g
    return torch.gather(feature_map, 2, idx)


def invert_image_features(f: torch.Tensor, rotation_deg: torch.Tensor, reflection: Optional[torch.Tensor],
                          num_rotations: int, num_group: int, induced_rep_type: str = "regular") -> torch.Tensor:
    """images/utils.py:32-94 (get_action_on_image_features) as called by discrete_group.py:240-259."""
    assert f.dim() == 4
    b, c, h, w = f.shape
    if induced_rep_type not in ("regular", "scalar", "vector"):
        raise ValueError("induced_rep_type must be regular, scalar or vector")
    if induced_rep_type == "vector":
        raise NotImplementedError("Action for vector representation is not implemented")
    if induced_rep_type == "regular":
        assert c % num_group == 0
    out = K.rotate(f, rotation_deg)
    if reflection is not None:
        r = reflection[:, None, None, None]
        out = out * r + K.hflip(out) * (1 - r)  # utils.py:62-64 (polarity opposite to canonicalize)
    if induced_rep_type == "scalar":
        return out
    out = out.reshape(b, c // num_group, num_group, h, w)
    shift = rotation_deg / 360.0 * num_rotations
    if reflection is not None:
        out = torch.cat([roll_by_gather(out[:, :, :num_rotations], shift),
                         roll_by_gather(out[:, :, num_rotations:], -shift)], dim=2)
    else:
        out = roll_by_gather(out, shift)
    return out.reshape(b, -1, h, w)


# --------------------------------------------------------------------------------------------
# a12  optimisation-based variant: orbit expand + cosine similarity
# --------------------------------------------------------------------------------------------


def group_augment(x: torch.Tensor, num_rotations: int, reflect: bool, resize_shape: int) -> torch.Tensor:
    """discrete_group.py:387-427: |G| copies, group-major; rotate THEN flip (quirk A.4-6)."""
    gray = x.shape[1] == 1
    degrees = torch.linspace(0, 360, num_rotations + 1)[:-1].to(x.dtype)
    outs = []
    for refl in ([False, True] if reflect else [False]):
        for d in degrees:
            xr = x
            if not gray:
                p = math.ceil(resize_shape * 0.5)
                xr = F.pad(xr, (p, p, p, p), mode="replicate")
            xr = K.rotate(xr, -d)
            if refl:
                xr = K.hflip(xr)
            if not gray:
                xr = center_crop(xr, resize_shape, resize_shape)
            outs.append(xr)
    return torch.cat(outs, dim=0)


def cosine_group_activations(vector_out: torch.Tensor, reference_vector: torch.Tensor, num_group: int) -> torch.Tensor:
    """discrete_group.py:475-481."""
    s = F.cosine_similarity(reference_vector.repeat(vector_out.shape[0], 1), vector_out)
    return s.reshape(num_group, -1).T


def optimization_specific_loss(vector_out: torch.Tensor, num_group: int, out_vector_size: int) -> torch.Tensor:
    """discrete_group.py:483-512 with artifact_err_wt == 0."""
    v = vector_out.reshape(num_group, -1, out_vector_size).permute(1, 0, 2)
    d = v @ v.permute(0, 2, 1)
    mask = 1.0 - torch.eye(num_group, dtype=v.dtype)
    return torch.abs(d * mask).mean()


# --------------------------------------------------------------------------------------------
# a13 / a14  prior regularisation statistic and identity metric
# --------------------------------------------------------------------------------------------


def prior_loss_discrete(act: torch.Tensor) -> torch.Tensor:
    """basecanonicalization.py:290-301: CE(act, class 0), mean over the batch (no beta)."""
    return F.cross_entropy(act, torch.zeros(act.shape[0], dtype=torch.long))


def identity_metric_discrete(act: torch.Tensor) -> torch.Tensor:
    """basecanonicalization.py:303-311."""
    return (act.argmax(dim=-1) == 0).float().mean()


def prior_loss_continuous(rep: torch.Tensor) -> torch.Tensor:
    """basecanonicalization.py:390-408: MSE(R, I)."""
    eye = torch.eye(rep.shape[-1], dtype=rep.dtype).repeat(rep.shape[0], 1, 1)
    return F.mse_loss(rep, eye)


def identity_metric_continuous(rep: torch.Tensor) -> torch.Tensor:
    """basecanonicalization.py:410-430."""
    return 1.0 - prior_loss_continuous(rep)


# --------------------------------------------------------------------------------------------
# a15 .. a17  frames: Gram-Schmidt, SO(3) apply, E(3) apply / invert
# --------------------------------------------------------------------------------------------


def gram_schmidt(v: torch.Tensor) -> torch.Tensor:
    """common/utils.py:22-51 (classical GS on the three ROWS, no eps, no handedness fix)."""
    v1 = v[:, 0]
    v1 = v1 / torch.norm(v1, dim=1, keepdim=True)
    v2 = v[:, 1] - torch.sum(v[:, 1] * v1, dim=1, keepdim=True) * v1
    v2 = v2 / torch.norm(v2, dim=1, keepdim=True)
    v3 = (v[:, 2] - torch.sum(v[:, 2] * v1, dim=1, keepdim=True) * v1
          - torch.sum(v[:, 2] * v2, dim=1, keepdim=True) * v2)
    v3 = v3 / torch.norm(v3, dim=1, keepdim=True)
    return torch.stack([v1, v2, v3], dim=1)


def modified_gram_schmidt(v: torch.Tensor) -> torch.Tensor:
    """nbody/canonicalization/euclidean_group.py:139-157."""
    v1 = v[:, 0]
    v1 = v1 / torch.norm(v1, dim=1, keepdim=True)
    v2 = v[:, 1] - torch.sum(v[:, 1] * v1, dim=1, keepdim=True) * v1
    v2 = v2 / torch.norm(v2, dim=1, keepdim=True)
    v3 = v[:, 2] - torch.sum(v[:, 2] * v1, dim=1, keepdim=True) * v1
    v3 = v3 - torch.sum(v3 * v2, dim=1, keepdim=True) * v2
    v3 = v3 / torch.norm(v3, dim=1, keepdim=True)
    return torch.stack([v1, v2, v3], dim=1)


def so3_canonicalize(x: torch.Tensor, rotation: torch.Tensor) -> torch.Tensor:
    """pointcloud/canonicalization/continuous_group.py:66-81: x (B,3,N) -> bmm(x^T, R^T)^T == R x."""
    return torch.bmm(x.transpose(1, 2), rotation.transpose(1, 2)).transpose(1, 2)


def e3_canonicalize(loc: torch.Tensor, vel: torch.Tensor, rotation: torch.Tensor, translation: torch.Tensor):
    """nbody/canonicalization/euclidean_group.py:108-124 (rows are particles, R and t per row)."""
    rinv = rotation.transpose(1, 2)
    cl = torch.bmm(loc[:, None, :], rinv).squeeze(1) - torch.bmm(translation[:, None, :], rinv).squeeze(1)
    cv = torch.bmm(vel[:, None, :], rinv).squeeze(1)
    return cl, cv


def e3_invert(x: torch.Tensor, rotation: torch.Tensor, translation: torch.Tensor) -> torch.Tensor:
    """euclidean_group.py:126-137: x R + t."""
    return torch.bmm(x[:, None, :], rotation).squeeze(1) + translation


# --------------------------------------------------------------------------------------------
# N1  frame-predicting vector-neuron networks (eval mode), on the reference modules' state dicts
# --------------------------------------------------------------------------------------------
VN_EPS = 1e-6  # vector_neuron_layers.py:13, nbody custom_group_equivariant_layers.py:4


def _vn_batchnorm(x: torch.Tensor, sd: dict, prefix: str, bn_eps: float = 1e-5) -> torch.Tensor:
    """VNBatchNorm in eval mode (vector_neuron_layers.py:309-322): x (B, C, 3, ...)."""
    kind = "bn2d" if prefix + "bn2d.weight" in sd else "bn1d"
    w, b = sd[f"{prefix}{kind}.weight"], sd[f"{prefix}{kind}.bias"]
    rm, rv = sd[f"{prefix}{kind}.running_mean"], sd[f"{prefix}{kind}.running_var"]
    norm = torch.norm(x, dim=2) + VN_EPS
    norm_bn = torch.nn.functional.batch_norm(norm, rm, rv, w, b, False, 0.0, bn_eps)   # nn.BatchNorm{1,2}d in eval()
    return x / norm.unsqueeze(2) * norm_bn.unsqueeze(2)


def _vn_linear_leaky_relu(x: torch.Tensor, sd: dict, prefix: str, negative_slope: float = 0.0) -> torch.Tensor:
    """VNLinearLeakyReLU.forward (vector_neuron_layers.py:253-273)."""
    p = torch.nn.functional.linear(x.transpose(1, -1), sd[prefix + "map_to_feat.weight"]).transpose(1, -1)
    p = _vn_batchnorm(p, sd, prefix + "batchnorm.")
    d = torch.nn.functional.linear(x.transpose(1, -1), sd[prefix + "map_to_dir.weight"]).transpose(1, -1)
    dotprod = (p * d).sum(2, keepdim=True)
    mask = (dotprod >= 0).to(x.dtype)
    d_norm_sq = (d * d).sum(2, keepdim=True)
    return negative_slope * p + (1 - negative_slope) * (mask * p + (1 - mask) * (p - (dotprod / (d_norm_sq + VN_EPS)) * d))


def vnsmall_forward(x: torch.Tensor, sd: dict, n_knn: int) -> torch.Tensor:
    """VNSmall.forward, pooling "mean", eval (equivariant_networks.py:128-150; knn :15-33; get_graph_feature_cross
    :36-76).  x (B,3,N) -> (B,3,3)."""
    b, _, n = x.shape
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    idx = (-xx - inner - xx.transpose(2, 1)).topk(k=n_knn, dim=-1)[1]                    # (B, N, k)
    xt = x.transpose(2, 1).contiguous()                                                   # (B, N, 3)
    feature = torch.gather(xt.unsqueeze(1).expand(b, n, n, 3), 2, idx.unsqueeze(-1).expand(b, n, n_knn, 3))
    feature = feature.view(b, n, n_knn, 1, 3)
    xr = xt.view(b, n, 1, 1, 3).repeat(1, 1, n_knn, 1, 1)
    cross = torch.cross(feature, xr, dim=-1)
    feat = torch.cat((feature - xr, xr, cross), dim=3).permute(0, 3, 4, 1, 2).contiguous()  # (B, 3, 3, N, k)
    out = _vn_linear_leaky_relu(feat, sd, "conv_pos.")
    out = out.mean(dim=-1)                                                                 # mean_pool (:141)
    out = _vn_batchnorm(_vn_linear_leaky_relu(out, sd, "conv1."), sd, "bn1.")
    out = _vn_linear_leaky_relu(out, sd, "conv2.")
    return out.mean(dim=-1)[:, :3]


def _segment_reduce(src: torch.Tensor, index: torch.Tensor, n: int, reduce: str) -> torch.Tensor:
    """torch_scatter.scatter(src, index, 0, reduce=) (third-party, absent here; unambiguous segment sum / mean)."""
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    res.index_add_(0, index, src)
    if reduce == "mean":
        cnt = torch.zeros(n, dtype=src.dtype).index_add_(0, index, torch.ones(index.shape[0], dtype=src.dtype))
        res = res / cnt.clamp(min=1).view((n,) + (1,) * (src.dim() - 1))
    return res


def vndeepsets_forward(loc: torch.Tensor, vel: torch.Tensor, charges: torch.Tensor, edges: torch.Tensor, sd: dict,
                       num_layers: int, nonlinearity: str, canon_feature: str, layer_pooling: str, final_pooling: str,
                       canon_translation: bool):
    """VNDeepSets.forward, eval (nbody custom_equivariant_networks.py:106-172; VNDeepSetLayer :228-252; VNLeakyReLU /
    VNSoftplus custom_group_equivariant_layers.py:31-52, :84-99) -> rotation vectors (M,3,3), translation (M,3)."""
    m = loc.shape[0]
    systems = m // 5
    batch_indices = torch.arange(systems).reshape(-1, 1).repeat(1, 5).reshape(-1)
    mean_loc = _segment_reduce(loc, batch_indices, systems, layer_pooling)
    mean_loc = mean_loc.repeat(5, 1, 1).transpose(0, 1).reshape(-1, 3)
    cl = loc - mean_loc
    chans = [cl]
    if "v" in canon_feature:
        chans.append(vel)
    if "a" in canon_feature:
        chans.append(torch.linalg.cross(cl, vel, dim=1))
    if "c" in canon_feature:
        chans.append(cl * charges)
    x = torch.stack(chans, dim=2)                                                          # (M, 3, Cin)
    ns = 0.2 if nonlinearity == "leakyrelu" else 0.0
    for l in range(num_layers):
        pre = "first_set_layer." if l == 0 else f"set_layers.{l - 1}."
        identity = torch.nn.functional.linear(x, sd[pre + "identity_linear.weight"], sd[pre + "identity_linear.bias"])
        pooled = _segment_reduce(torch.index_select(x, 0, edges[0]), edges[1], m, layer_pooling)
        pooling = torch.nn.functional.linear(pooled, sd[pre + "pooling_linear.weight"], sd[pre + "pooling_linear.bias"])
        y = (identity + pooling).transpose(1, -1)                                          # (M, H, 3)
        d = torch.nn.functional.linear(y.transpose(1, -1), sd[pre + "nonlinear_function.map_to_dir.weight"]).transpose(1, -1)
        dotprod = (y * d).sum(2, keepdim=True)
        if nonlinearity == "softplus":
            ang = torch.acos(dotprod / (torch.norm(y, dim=2, keepdim=True) * torch.norm(d, dim=2, keepdim=True) + VN_EPS))
            mask = torch.cos(ang / 2) ** 2
        else:
            mask = (dotprod >= 0).to(y.dtype)
        d_norm_sq = (d * d).sum(2, keepdim=True)
        out = ns * y + (1 - ns) * (mask * y + (1 - mask) * (y - (dotprod / (d_norm_sq + VN_EPS)) * d))
        out = out.transpose(1, -1)
        x = out + x if l > 0 else out
    xs = _segment_reduce(x, batch_indices, systems, final_pooling)
    output = torch.nn.functional.linear(xs, sd["output_layer.weight"], sd["output_layer.bias"])
    output = output.repeat(5, 1, 1, 1).transpose(0, 1).reshape(-1, 3, 4)
    rotation_vectors = output[:, :, :3]
    translation = output[:, :, 3:] if canon_translation else 0.0
    translation = translation + mean_loc[:, :, None]
    return rotation_vectors, translation.squeeze()


# --------------------------------------------------------------------------------------------
# N2  continuous-group image canonicalization
# --------------------------------------------------------------------------------------------
def canonicalize_image_continuous(x: torch.Tensor, rotation_matrices: torch.Tensor,
                                  reflection: Optional[torch.Tensor]) -> torch.Tensor:
    """ContinuousGroupImageCanonicalization.canonicalize, images/canonicalization/continuous_group.py:162-210, given the
    group element (rotation (B,2,2) as get_group_from_out_vectors returns it, reflection (B,1,1,1) or None).
    Does NOT modify `rotation_matrices` (the reference negates its off-diagonal entries in place, :180)."""
    from oracle import kornia_restated as K
    rot = rotation_matrices.clone()
    rot[:, [0, 1], [1, 0]] *= -1
    if reflection is not None:
        x = (1 - reflection) * x + reflection * K.hflip(x)
    c, h, w = x.shape[1:]
    if c != 1:
        p = math.ceil(w * 0.5)
        x = F.pad(x, (p, p, p, p), mode="replicate")
    alpha, beta = rot[:, 0, 0], rot[:, 0, 1]
    cx, cy = x.shape[-2] // 2, x.shape[-1] // 2
    affine_part = torch.stack([(1 - alpha) * cx - beta * cy, beta * cx + (1 - alpha) * cy], dim=1)
    m = torch.cat([rot, affine_part.unsqueeze(-1)], dim=-1)
    x = K.warp_affine(x, m, dsize=(x.shape[-2], x.shape[-1]))
    if c != 1:
        x = center_crop(x, h, w)
    return x


def group_augment_continuous(x: torch.Tensor, angles: torch.Tensor, reflect: Optional[torch.Tensor]):
    """OptimizedSteerableImageCanonicalization.group_augment, continuous_group.py:362-412, with the random draws
    (`angles` in radians, `reflect` = +-1 or None) supplied -> (augmented images, (B,2,2) matrices)."""
    b, c, h, w = x.shape
    cos_a, sin_a = torch.cos(angles), torch.sin(angles)
    rm = torch.zeros(b, 2, 3, dtype=x.dtype)
    rm[:, :2, :2] = torch.stack((cos_a, -sin_a, sin_a, cos_a)).reshape(-1, 2, 2)
    if reflect is not None:
        rm[:, 0, 0] *= reflect
    if c != 1:
        p = math.ceil(w * 0.5)
        x = F.pad(x, (p, p, p, p), mode="replicate")
    grid = F.affine_grid(rm, list(x.size()), align_corners=False)
    aug = F.grid_sample(x, grid, align_corners=False)
    if c != 1:
        aug = center_crop(aug, h, w)
    rm[:, [0, 1], [1, 0]] *= -1
    return aug, rm[:, :, :2]


# --------------------------------------------------------------------------------------------
# N4  evaluation-time group orbit (examples/images/classification/inference_utils.py:97-122)
# --------------------------------------------------------------------------------------------


def linspace_degrees(num_rotations: int) -> List[float]:
    """`torch.linspace(0, 360, n + 1)[:-1]` as float32 values (inference_utils.py:99): torch fills the first half as
    start + i*step and the second half as end - (steps-1-i)*step, all in float32 (the scalar formula of ATen's
    linspace kernel; its vectorised CPU path adds lane*step to a per-vector base instead, so when 360/n is not a float32
    the reference's own value moves by an ulp with the SIMD width of the host -- only rounding-tie pixels can notice)."""
    import numpy as np
    steps = num_rotations + 1
    step = np.float32(360.0) / np.float32(steps - 1)
    half = steps // 2
    out = []
    for i in range(num_rotations):
        v = np.float32(0.0) + step * np.float32(i) if i < half else np.float32(360.0) - step * np.float32(steps - 1 - i)
        out.append(float(np.float32(v)))
    return out


def torchvision_rotate_matrix(angle_deg: float) -> List[float]:
    """torchvision.transforms.functional.rotate -> _get_inverse_affine_matrix([0,0], -angle, [0,0], 1, [0,0]) in
    Python doubles (torchvision 0.26 transforms/functional.py); the six entries of the output->input map in
    centred pixel coordinates."""
    rot = math.radians(-angle_deg)
    a, b, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)
    return [d, -b, 0.0, -c, a, 0.0]


def group_inference_orbit(x: torch.Tensor, num_rotations: int, reflect: bool, return_margin: bool = False):
    """GroupInference.get_group_element_wise_logits' inputs (inference_utils.py:97-122): for every group element
    Pad(ceil(0.4 H), edge) -> [hflip] -> torchvision rotate (NEAREST, zero fill, expand=False) -> CenterCrop(H, W).
    Returns (G, B, C, H, W) with rotations first, then the reflected rotations (dict keys rot, rot + n).

    torchvision's tensor rotate is affine grid (float32 bmm) + grid_sample(nearest, zeros, align_corners=False);
    restated here as explicit index arithmetic.  With `return_margin` also returns (G, H, W) float64 distances of
    the source coordinate from the nearest rounding tie (x.5): pixels with a tiny margin (the diagonals of the
    45-degree elements sit EXACTLY on ties) depend on the order of float32 operations of the matmul backend, so
    parity is asserted away from them."""
    b, c, h, w = x.shape
    p = math.ceil(h * 0.4)
    hp, wp = h + 2 * p, w + 2 * p
    top, left = center_crop_offsets(hp, wp, h, w)
    xs32 = (torch.arange(w, dtype=torch.float32) + left) + (-wp * 0.5 + 0.5)
    ys32 = (torch.arange(h, dtype=torch.float32) + top) + (-hp * 0.5 + 0.5)
    outs, margins = [], []
    for flip in ([False, True] if reflect else [False]):
        for deg in linspace_degrees(num_rotations):
            m = torchvision_rotate_matrix(deg)
            theta = torch.tensor(m, dtype=torch.float32).reshape(2, 3)
            resc = theta.t() / torch.tensor([0.5 * wp, 0.5 * hp], dtype=torch.float32)       # (3, 2)
            gx = (xs32[None, :] * resc[0, 0] + ys32[:, None] * resc[1, 0]) + resc[2, 0]
            gy = (xs32[None, :] * resc[0, 1] + ys32[:, None] * resc[1, 1]) + resc[2, 1]
            ix = ((gx + 1) * wp - 1) / 2
            iy = ((gy + 1) * hp - 1) / 2
            jx, jy = torch.round(ix).long(), torch.round(iy).long()        # torch.round = nearbyint (ties to even)
            inside = (jx >= 0) & (jx < wp) & (jy >= 0) & (jy < hp)
            if flip:
                jx = wp - 1 - jx
            sx, sy = (jx - p).clamp(0, w - 1), (jy - p).clamp(0, h - 1)     # edge padding
            out = x[:, :, sy, sx] * inside.to(x.dtype)
            outs.append(out)
            if return_margin:
                xs64, ys64 = xs32.double(), ys32.double()
                fx = xs64[None, :] * m[0] + ys64[:, None] * m[1] + (wp - 1) / 2
                fy = xs64[None, :] * m[3] + ys64[:, None] * m[4] + (hp - 1) / 2
                tie = lambda v: (v - torch.floor(v) - 0.5).abs()
                margins.append(torch.minimum(tie(fx), tie(fy)))
    orbit = torch.stack(outs, 0)
    return (orbit, torch.stack(margins, 0)) if return_margin else orbit
