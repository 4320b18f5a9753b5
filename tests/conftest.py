import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """tests/golden/<name>.npz (written by oracle/make_golden.py from the unmodified reference)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        v = z[k]
        if v.dtype.kind in "US":
            out[k] = str(v)
        elif v.ndim == 0 and v.dtype.kind in "iu":
            out[k] = int(v)
        elif v.ndim == 0 and k in ("crop_ratio", "beta"):
            out[k] = float(v)
        else:
            out[k] = torch.from_numpy(np.ascontiguousarray(v))
    return out


IMAGE_CASES = ["image_c4_cfg1", "image_c8_small", "image_d4_small", "image_d8_small", "image_c4_gray",
               "image_c8_rect"]


def golden_layers(g):
    """[(W, b), ...] of the CustomEquivariantNetwork stored in an image golden file."""
    layers = []
    i = 0
    while f"w{i}" in g:
        layers.append((g[f"w{i}"], g[f"b{i}"]))
        i += 2
    return layers


def resize_arg(g):
    r = g["resize"]
    if torch.is_tensor(r):
        return tuple(int(v) for v in r.tolist())
    return int(r)


def rel_err(a, b):
    """||a-b||_inf / ||b||_inf  (the metric SURVEY.md section 7 defines for the 1e-4 bar)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
