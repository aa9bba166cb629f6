"""C1 / §8(f)#1 on the device: one eigendecomposition per factor feeds the prior-precision optimiser and the covariance."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _spd(gen, d, scale):
    w = torch.randn(4 * d, d, generator=gen, dtype=torch.float64)
    return ((w.T @ w) / math.sqrt(4 * d) * scale).float()


@pytest.mark.parametrize("d_in,d_emb", [(768, 512), (1024, 768)])
def test_factor_spectrum_on_device(d_in, d_emb):
    """Regularised inverses from the spectra equal torch.linalg.inv of the same matrices (reference hessians.py:170-184) and the
    spectral log-determinant equals torch.logdet (hessians.py:276-280), at the config 2 / 3 factor sizes."""
    from bayesvlm_b200.hessians import FactorSpectrum, _compute_covariance, covariance_from_spectra

    gen = torch.Generator().manual_seed(d_in)
    A, B = _spd(gen, d_in, 3e3).cuda(), _spd(gen, d_emb, 20.0).cuda()
    n, lam = 1.0, 605.25
    spec = (FactorSpectrum.of(A), FactorSpectrum.of(B))
    assert spec[0].evecs.is_cuda and spec[0].evals.dtype == torch.float64
    cov = covariance_from_spectra(spec[0], spec[1], n, lam)
    ref = _compute_covariance(A.double(), B.double(), torch.tensor(n, dtype=torch.float64, device="cuda"),
                              torch.tensor(lam, dtype=torch.float64, device="cuda"))
    for ours, theirs in ((cov.A_inv, ref.A_inv), (cov.B_inv, ref.B_inv)):
        assert ours.dtype == torch.float32
        assert float((ours.double() - theirs).abs().max() / theirs.abs().max()) <= 1e-6
    eye = torch.eye(d_in, device="cuda", dtype=torch.float64)
    ld = torch.logdet(A.double() * math.sqrt(n) + math.sqrt(lam) * eye)
    assert abs(float(spec[0].logdet_regularised(math.sqrt(n), math.sqrt(lam))) - float(ld)) <= 1e-8 * abs(float(ld))


def test_prior_precision_with_shared_spectra_on_device(golden):
    """optimize_prior_precision on the device, with and without precomputed spectra, against the reference's Adam result
    (tests/golden/make_golden.py: reference hessians.py:219-265 on the CPU)."""
    from bayesvlm_b200.hessians import FactorSpectrum, optimize_prior_precision

    lam0, n, lr, steps = golden["prior_cfg"]
    proj = torch.nn.Linear(24, 16, bias=False).cuda()
    with torch.no_grad():
        proj.weight.copy_(torch.from_numpy(golden["prior_W"]))
    A, B = torch.from_numpy(golden["prior_A"]).cuda(), torch.from_numpy(golden["prior_B"]).cuda()
    kw = dict(lmbda_init=float(lam0), n=float(n), lr=float(lr), num_steps=int(steps), device="cuda")
    lam_plain = optimize_prior_precision(proj, A, B, **kw)
    lam_spec = optimize_prior_precision(proj, A, B, spectra=(FactorSpectrum.of(A), FactorSpectrum.of(B)), **kw)
    ref = float(golden["prior_lambda"][0])
    assert abs(lam_plain.item() - ref) <= 2e-3 * ref
    assert abs(lam_spec.item() - lam_plain.item()) <= 1e-5 * ref
