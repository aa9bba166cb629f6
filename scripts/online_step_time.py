"""Cost of the per-pick posterior update of select_epig_online (eigendecompositions, prior precision, covariance):
python scripts/online_step_time.py"""
import math, sys, time, torch
sys.path.insert(0, ".")
from bayesvlm_b200.hessians import FactorSpectrum, covariance_from_spectra, optimize_prior_precision, _compute_covariance
g = torch.Generator().manual_seed(0)
spd = lambda d, sc: (lambda w: (w.T @ w) / math.sqrt(4 * d) * sc)(torch.randn(4 * d, d, generator=g))
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for d_in, d in ((768, 512), (1024, 768)):
    A, B = spd(d_in, 3e3).cuda(), spd(d, 20.0).cuda()
    proj = torch.nn.Linear(d_in, d, bias=False).cuda()
    print(d_in, d, "eigh fp64 A+B ms", round(t(lambda: (FactorSpectrum.of(A), FactorSpectrum.of(B))), 2),
          "| eigh fp32 A+B ms", round(t(lambda: (torch.linalg.eigh(A), torch.linalg.eigh(B))), 2),
          "| inv x4 fp32 ms", round(t(lambda: [torch.linalg.inv(M) for M in (A, B, A, B)]), 2))
    sp = (FactorSpectrum.of(A), FactorSpectrum.of(B))
    print("   prior precision 20 steps with spectra ms", round(t(lambda: optimize_prior_precision(proj, A, B, 600.0, 1.0, 1e-3, 20, "cuda", spectra=sp)), 2),
          "| without ms", round(t(lambda: optimize_prior_precision(proj, A, B, 600.0, 1.0, 1e-3, 20, "cuda")), 2),
          "| covariance from spectra ms", round(t(lambda: covariance_from_spectra(sp[0], sp[1], 1.0, 600.0)), 2))
