"""Shim: kornia.geometry.{rotate,hflip,warp_affine} restated in oracle/kornia_restated.py."""
from oracle.kornia_restated import hflip, rotate, warp_affine  # noqa: F401
