"""Small-shape pass over every kernel family (target for compute-sanitizer --tool memcheck)."""
import math, sys, numpy as np, torch
sys.path.insert(0, ".")
from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC, _ggn, syrk_accumulate
from bayesvlm_b200.vlm import CLIP, EncoderResult, sample_probas_from_noise
from bayesvlm_b200.epig import epig_from_probs_using_matmul
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g)
spd = lambda d, sc: (lambda w: (w.T @ w) / math.sqrt(4 * d) * sc)(rn(4 * d, d))
inv = lambda F, lam: torch.linalg.inv(F.double() + math.sqrt(lam) * torch.eye(F.shape[0], dtype=torch.float64)).float()
D, d_i, d_t, N, C = 96, 130, 70, 300, 37
for prec in ("fp16x3", "fp16+fp8", "fp16"):
    m = CLIP(logit_scale=math.log(100.0), device="cuda", precision=prec)
    m.set_covariances(KFC(inv(spd(d_i, 3e3), 600).cuda(), inv(spd(D, 20), 600).cuda()), KFC(inv(spd(d_t, 3e3), 220).cuda(), inv(spd(D, 20), 220).cuda()))
    with torch.no_grad():
        out, pr = m._compute_probabilistic_logits_smith(EncoderResult(rn(N, D).cuda(), rn(N, d_i).cuda()), EncoderResult(rn(C, D).cuda(), rn(C, d_t).cuda()), return_probs=True)
    assert torch.isfinite(out.mean).all() and torch.isfinite(pr).all()
for siglip in (False, True):
    for prec in ("fp16", "fp16x3"):
        H = _ggn(rn(333, 72).cuda(), rn(515, 72).cuda(), math.log(50.0), -5.0, siglip, precision=prec)
        assert torch.isfinite(H).all()
A = syrk_accumulate(rn(777, 129).cuda(), append_one=True)
assert torch.isfinite(A).all()
mk = lambda n, k, cl: sample_probas_from_noise((rn(n, cl) * 2).cuda(), (torch.rand(n, cl, generator=g) * 3 + 0.1).cuda(), rn(k, n, cl).cuda())
s = epig_from_probs_using_matmul(mk(130, 20, 7), mk(90, 20, 7), chunk_size=256)
assert torch.isfinite(s).all()
torch.cuda.synchronize(); print("sanitize pass ok")
