"""Time the evaluation-orbit kernel (eqb_orbit_rotate_nearest) against the torchvision loop it replaces
(examples/images/classification/inference_utils.py:97-122), on the GPU (development aid)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200 import ops


def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


def torchvision_loop(x, n, reflect):
    from torchvision import transforms
    pad = transforms.Pad(math.ceil(x.shape[-2] * 0.4), padding_mode="edge")
    crop = transforms.CenterCrop(tuple(x.shape[-2:]))
    degs = torch.linspace(0, 360, n + 1)[:-1]
    out = [crop(transforms.functional.rotate(pad(x), d.item())) for d in degs]
    if reflect:
        out += [crop(transforms.functional.rotate(transforms.functional.hflip(pad(x)), d.item())) for d in degs]
    return out


for (b, c, h, w, n, reflect) in ((64, 3, 224, 224, 8, False), (64, 3, 224, 224, 8, True), (128, 3, 32, 32, 4, False), (512, 3, 64, 64, 4, True)):
    x = torch.rand(b, c, h, w, device="cuda")
    g = n * (2 if reflect else 1)
    us = timed(lambda: ops.orbit_rotate_nearest(x, n, reflect))
    us_tv = timed(lambda: torchvision_loop(x, n, reflect), 5)
    same = all(torch.equal(a, o) for a, o in zip(torchvision_loop(x, n, reflect), ops.orbit_rotate_nearest(x, n, reflect)))
    byt = (g + 1) * x.numel() * 4
    print(f"orbit {b}x{c}x{h}x{w} |G|={g}: {us:.1f} us, {byt / us / 1e3:.0f} GB/s algorithmic ((|G|+1) planes); torchvision loop on the same GPU "
          f"{us_tv:.0f} us ({us_tv / us:.1f}x); bit-identical to it: {same}")
