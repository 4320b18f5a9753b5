"""ORACLE (test infrastructure, never on the product path).

Restatement of the three `kornia.geometry` functions the reference calls on the
canonicalization hot path.  kornia is a third-party dependency that is NOT vendored
under /root/reference and is NOT installed in this image:

    kornia, pinned `kornia=0.7.0` in /root/reference/conda_env.yaml:44
    (unpinned in setup.cfg:48).

The restatement follows the published kornia 0.7.0 algorithm
(`kornia/geometry/transform/affwarp.py::rotate/affine/_compute_rotation_matrix`,
`imgwarp.py::warp_affine/get_rotation_matrix2d`,
`conversions.py::angle_to_rotation_matrix/deg2rad/normalize_homography/normal_transform_pixel`)
op for op, in the same order, on torch CPU ops, so its fp32 rounding is the reference's.
Parity at this boundary is **unpinned**: the reference's own tests hold no numeric
assertion on a kornia output (tests/images/canonicalization/test_discrete_group.py:43-69
asserts nothing).  It is anchored by (i) the reference's own hand-derived affine matrix in
equiadapt/images/canonicalization/continuous_group.py:195-204, which is the same
OpenCV `getRotationMatrix2D` formula, and (ii) the closed-form properties checked in
tests/test_oracle_golden.py (rotate(+90) == rot90(k=1) up to fp32 coefficient noise,
1x1 kernels unchanged, closed-form sampling positions).

Reference call sites of these functions:
  equiadapt/images/canonicalization/discrete_group.py:211,213,404,406,456,463
  equiadapt/images/utils.py:57,61,82,85
  equiadapt/images/canonicalization_networks/custom_group_equivariant_layers.py:77,184,190,313,481,487
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def deg2rad(t: torch.Tensor) -> torch.Tensor:
    # kornia.geometry.conversions.deg2rad: tensor * pi / 180 with pi cast to tensor dtype
    pi = torch.tensor(math.pi, dtype=t.dtype, device=t.device)
    return t * pi / 180.0


def angle_to_rotation_matrix(angle: torch.Tensor) -> torch.Tensor:
    ang = deg2rad(angle)
    c = torch.cos(ang)
    s = torch.sin(ang)
    return torch.stack([c, s, -s, c], dim=-1).view(*angle.shape, 2, 2)


def get_rotation_matrix2d(center: torch.Tensor, angle: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """OpenCV getRotationMatrix2D as kornia 0.7.0 computes it (B,2,3)."""
    rot = angle_to_rotation_matrix(angle)
    b = center.shape[0]
    scaling = torch.zeros(2, 2, dtype=rot.dtype, device=rot.device).fill_diagonal_(1).repeat(b, 1, 1)
    scaling = scaling * scale.unsqueeze(dim=2).repeat(1, 1, 2)
    scaled = rot @ scaling
    alpha = scaled[:, 0, 0]
    beta = scaled[:, 0, 1]
    x = center[..., 0]
    y = center[..., 1]
    m = torch.zeros(b, 2, 3, dtype=center.dtype, device=center.device)
    m[..., 0:2, 0:2] = scaled
    m[..., 0, 2] = (1.0 - alpha) * x - beta * y
    m[..., 1, 2] = beta * x + (1.0 - alpha) * y
    return m


def normal_transform_pixel(height: int, width: int, dtype, device, eps: float = 1e-14) -> torch.Tensor:
    tr = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], dtype=dtype, device=device)
    wd = eps if width == 1 else width - 1.0
    hd = eps if height == 1 else height - 1.0
    tr[0, 0] = tr[0, 0] * 2.0 / wd
    tr[1, 1] = tr[1, 1] * 2.0 / hd
    return tr.unsqueeze(0)


def _inverse(m: torch.Tensor) -> torch.Tensor:
    # kornia.utils.helpers._torch_inverse_cast: LU inverse in the tensor's own fp32/fp64
    dt = m.dtype if m.dtype in (torch.float32, torch.float64) else torch.float32
    return torch.linalg.inv(m.to(dt)).to(m.dtype)


def normalize_homography(dst_pix_trans_src_pix: torch.Tensor, dsize_src, dsize_dst) -> torch.Tensor:
    sh, sw = dsize_src
    dh, dw = dsize_dst
    src_norm_trans_src_pix = normal_transform_pixel(sh, sw, dst_pix_trans_src_pix.dtype, dst_pix_trans_src_pix.device)
    src_pix_trans_src_norm = _inverse(src_norm_trans_src_pix)
    dst_norm_trans_dst_pix = normal_transform_pixel(dh, dw, dst_pix_trans_src_pix.dtype, dst_pix_trans_src_pix.device)
    return dst_norm_trans_dst_pix @ (dst_pix_trans_src_pix @ src_pix_trans_src_norm)


def warp_affine(src: torch.Tensor, M: torch.Tensor, dsize, mode: str = "bilinear",
                padding_mode: str = "zeros", align_corners: bool = True) -> torch.Tensor:
    b, c, h, w = src.shape
    m3 = F.pad(M, [0, 0, 0, 1], "constant", 0.0)
    m3[..., -1, -1] += 1.0
    dst_norm_trans_src_norm = normalize_homography(m3, (h, w), dsize)
    src_norm_trans_dst_norm = _inverse(dst_norm_trans_src_norm)
    grid = F.affine_grid(src_norm_trans_dst_norm[:, :2, :], [b, c, dsize[0], dsize[1]], align_corners=align_corners)
    return F.grid_sample(src, grid, align_corners=align_corners, mode=mode, padding_mode=padding_mode)


def rotate(tensor: torch.Tensor, angle: torch.Tensor, center=None, mode: str = "bilinear",
           padding_mode: str = "zeros", align_corners: bool = True) -> torch.Tensor:
    """kornia.geometry.rotate: anti-clockwise rotation by `angle` degrees about the image centre."""
    if not torch.is_tensor(angle):
        angle = torch.tensor(angle, dtype=tensor.dtype, device=tensor.device)
    angle = angle.to(tensor.dtype)
    if center is None:
        h, w = tensor.shape[-2:]
        center = torch.tensor([float(w - 1) / 2, float(h - 1) / 2], dtype=tensor.dtype, device=tensor.device)
    angle = angle.expand(tensor.shape[0])
    center = center.expand(tensor.shape[0], -1)
    m = get_rotation_matrix2d(center, angle, torch.ones_like(center))
    m = m[..., :2, :3].expand(tensor.shape[0], -1, -1)
    return warp_affine(tensor, m, (tensor.shape[-2], tensor.shape[-1]), mode, padding_mode, align_corners)


def hflip(x: torch.Tensor) -> torch.Tensor:
    return x.flip(-1)
