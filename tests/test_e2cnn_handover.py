"""a7 / N4: the e2cnn hand-over (`ESCNNEquivariantNetwork.load_e2cnn`, escnn_networks.py:59-117) exercised on a stand-in
module tree with the attribute names it walks.  e2cnn is not installable here, so the tree below is built from plain torch
modules that carry what e2cnn's modules carry in eval(): `R2Conv.filter` / `.expanded_bias` (the cached expanded tensors),
`InnerBatchNorm` with one `BatchNorm3d` child per representation size acting on (B, fields, |G|, H, W).  It is executable, so
it doubles as an independent torch statement of the dense arithmetic the kernel must reproduce.  Parity of e2cnn's own BASIS
EXPANSION stays unpinned: nothing here can check how e2cnn fills `filter`."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F


class R2Conv(nn.Module):
    def __init__(self, cin, fields, g, k, gen):
        super().__init__()
        self.register_buffer("filter", torch.randn(fields * g, cin, k, k, generator=gen) * (cin * k * k) ** -0.5)
        self.register_buffer("expanded_bias", torch.randn(fields, generator=gen).repeat_interleave(g) * 0.1)

    def forward(self, x):
        return F.conv2d(x, self.filter, self.expanded_bias)


class InnerBatchNorm(nn.Module):
    def __init__(self, fields, g, gen):
        super().__init__()
        self.g = g
        bn = nn.BatchNorm3d(fields, momentum=0.9)
        with torch.no_grad():
            bn.weight.copy_(torch.rand(fields, generator=gen) + 0.5)
            bn.bias.copy_(torch.randn(fields, generator=gen) * 0.1)
            bn.running_mean.copy_(torch.randn(fields, generator=gen) * 0.1)
            bn.running_var.copy_(torch.rand(fields, generator=gen) + 0.5)
        setattr(self, f"batch_norm_[{g}]", bn)          # e2cnn names the child after the representation size

    def forward(self, x):
        b, c, h, w = x.shape
        bn = next(self.children())
        return bn(x.reshape(b, c // self.g, self.g, h, w)).reshape(b, c, h, w)


class ReLU(nn.ReLU):
    pass


class PointwiseDropout(nn.Dropout):
    pass


def fake_reference_network(cin, fields, g, k, layers, seed):
    gen = torch.Generator().manual_seed(seed)
    mods = []
    for l in range(layers):
        mods.append(R2Conv(cin if l == 0 else fields * g, fields, g, k, gen))
        if l < layers - 1:
            mods += [InnerBatchNorm(fields, g, gen), ReLU(), PointwiseDropout(0.5)]
    net = nn.Module()
    net.eqv_network = nn.Sequential(*mods)
    return net.eval()


def _ours(fields, k, g, layers, device):
    from equiadapt_b200.images.canonicalization_networks.escnn_networks import ESCNNEquivariantNetwork
    return ESCNNEquivariantNetwork((3, 32, 32), fields, k, "rotation", g, layers, device=device).eval()


def test_load_e2cnn_copies_the_cached_tensors_and_folds_the_batch_norms():
    fields, g, k, layers = 6, 4, 3, 3
    ref = fake_reference_network(3, fields, g, k, layers, seed=3)
    net = _ours(fields, k, g, layers, "cpu").load_e2cnn(ref)
    convs = [m for m in ref.eqv_network if isinstance(m, R2Conv)]
    bns = [m for m in ref.eqv_network if isinstance(m, InnerBatchNorm)]
    for l, c in enumerate(convs):
        assert torch.equal(net.filters[l], c.filter) and torch.equal(net.biases[l], c.expanded_bias)
    scales, shifts = net.folded_affine()
    x = torch.randn(2, fields * g, 5, 5, generator=torch.Generator().manual_seed(4))
    for l, b in enumerate(bns):
        want = b(x)
        got = x * scales[l].view(1, -1, 1, 1) + shifts[l].view(1, -1, 1, 1)
        assert torch.allclose(got, want, atol=1e-6)
    # the folded affine is cached per batch-norm version, not recomputed per forward ...
    again = net.folded_affine()
    assert again[0][0] is scales[0]
    # ... and follows in-place updates of the statistics
    with torch.no_grad():
        getattr(net, "bn_running_mean_0").add_(1.0)
    moved = net.folded_affine()
    assert not torch.equal(moved[1][0], shifts[0])
    with pytest.raises(ValueError):
        _ours(fields, k, g, layers + 1, "cpu").load_e2cnn(ref)


@pytest.mark.gpu
def test_loaded_network_matches_the_stand_in_tree_on_the_gpu(cuda_device):
    fields, g, k, layers = 8, 4, 5, 3
    ref = fake_reference_network(3, fields, g, k, layers, seed=5)
    net = _ours(fields, k, g, layers, str(cuda_device)).load_e2cnn(ref.to(cuda_device))
    x = torch.rand(6, 3, 32, 32, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        out = ref.double().cpu().eqv_network(x.double())
        want = out.reshape(6, fields, g, out.shape[-2], out.shape[-1]).mean(dim=(1, 3, 4))      # escnn_networks.py:107-115
        got = net(x.to(cuda_device)).cpu().double()
    assert float((got - want).abs().max() / want.abs().max()) < 1e-4
