"""T2: Monte-Carlo methods of ProbabilisticLogits (reference vlm.py:68-103, 142-159) on the fused kernel against the
reference's own operation sequence in torch on the same GPU with the shared device RNG."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference_softmax(mean, var, num_samples, seed):
    torch.manual_seed(seed)
    std = torch.sqrt(var)
    probas = torch.zeros_like(mean)
    for _ in range(num_samples):  # vlm.py:86-89
        eps = torch.randn(std.shape, device=std.device) * std
        probas += torch.nn.functional.softmax(mean + eps, dim=-1)
    return probas / num_samples


def _reference_entropy(mean, var, num_samples, seed):
    torch.manual_seed(seed)
    entropy = 0
    for _ in range(num_samples):  # vlm.py:145-149
        eps = torch.randn(var.shape, device=var.device) * torch.sqrt(var)
        probas = torch.nn.functional.softmax(mean + eps, dim=-1)
        entropy += -(probas * probas.log()).sum(dim=-1)
    return entropy / num_samples


@pytest.mark.parametrize("n,c,k", [(333, 10, 7), (100, 1000, 5), (64, 257, 9), (2048, 65, 33), (5, 1024, 3)])
def test_mc_softmax_and_entropy_vs_reference_sequence(n, c, k, monkeypatch):
    from bayesvlm_b200 import _lib, vlm
    from bayesvlm_b200.vlm import ProbabilisticLogits

    gen = torch.Generator().manual_seed(n + c)
    mean = (torch.randn(n, c, generator=gen) * 3).cuda()
    var = (torch.rand(n, c, generator=gen) * 4 + 0.05).cuda()
    pl = ProbabilisticLogits(mean, var)
    l0 = _lib.launch_count()
    p = pl.softmax(num_samples=k, seed=11)
    assert _lib.launch_count() > l0, "the fused kernel did not run"
    ref = _reference_softmax(mean, var, k, 11)
    assert float((p - ref).abs().max()) <= 2e-6
    assert float((p.sum(-1) - 1).abs().max()) <= 1e-5
    torch.manual_seed(5)
    h = pl.expected_aleatoric_entropy(num_samples=k)
    href = _reference_entropy(mean, var, k, 5)
    assert float(((h - href).abs() / href.abs().clamp_min(1e-3)).max()) <= 2e-5
    # several launches (noise buffer smaller than the draws): same result, same RNG stream
    monkeypatch.setattr(vlm, "_MC_NOISE_BYTES", 4 * n * c * 2)
    p2 = pl.softmax(num_samples=k, seed=11)
    assert float((p2 - ref).abs().max()) <= 2e-6


def test_mc_fallbacks_keep_the_torch_expression():
    from bayesvlm_b200 import _lib
    from bayesvlm_b200.vlm import ProbabilisticLogits

    gen = torch.Generator().manual_seed(1)
    mean, var = torch.randn(8, 1500, generator=gen).cuda(), (torch.rand(8, 1500, generator=gen) + 0.1).cuda()
    l0 = _lib.launch_count()
    p = ProbabilisticLogits(mean, var).softmax(num_samples=3, seed=2)  # C > 1024: generic device expression
    assert _lib.launch_count() == l0
    assert float((p - _reference_softmax(mean, var, 3, 2)).abs().max()) == 0.0
    # num_samples = 0 keeps the reference's diagonal quirk (vlm.py:74-78)
    m2, v2 = mean[:, :8], var[:, :8]
    q = ProbabilisticLogits(m2, v2).softmax(num_samples=0)
    ref = torch.softmax(m2 / torch.sqrt(1 + torch.pi / 8 * v2.diagonal()), dim=-1)
    assert torch.equal(q, ref)
