"""Sizes beyond 2^31 output elements: predictive 300 000 x 10 000 (3e9 pairs, 24 GB of mean + var) checked on sampled rows against the
fp64 oracle, probit softmax on a row slice; GGN with a 65536-target class batch:  python scripts/big_shapes.py"""
import math, sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from oracle import laplace_oracle as O
from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC, compute_hessian_analytic_InfoNCE
from bayesvlm_b200.vlm import CLIP, EncoderResult
g = torch.Generator().manual_seed(9)
N, C, D, di, dt = 300_000, 10_000, 512, 768, 512
spd = lambda d, sc, lam: torch.linalg.inv(bench.surrogate_spd(g, d, sc).double() + math.sqrt(lam) * torch.eye(d, dtype=torch.float64)).float()
Ai, Bi, At, Bt = spd(di, 3e3, 600.0), spd(D, 20.0, 600.0), spd(dt, 3e3, 200.0), spd(D, 20.0, 200.0)
ie, ia = torch.randn(N, D, generator=g), torch.randn(N, di, generator=g)
te, ta = torch.randn(C, D, generator=g), torch.randn(C, dt, generator=g)
m = CLIP(logit_scale=bench.LS, device="cuda")
m.set_covariances(KFC(Ai.cuda(), Bi.cuda()), KFC(At.cuda(), Bt.cuda()))
with torch.no_grad():
    out = m(EncoderResult(ie.cuda(), ia.cuda()), EncoderResult(te.cuda(), ta.cuda()))
torch.cuda.synchronize()
assert out.mean.shape == (N, C) and out.mean.numel() > 2 ** 31
rows = torch.cat([torch.arange(0, 64), torch.randint(0, N, (128,), generator=g), torch.arange(N - 64, N)])
rm, rv = O.predictive(ie[rows].numpy(), ia[rows].numpy(), te.numpy(), ta.numpy(), Ai.numpy(), Bi.numpy(), At.numpy(), Bt.numpy(), bench.LS, dtype=np.float64)
mean, var = out.mean[rows.cuda()].double().cpu().numpy(), out.var[rows.cuda()].double().cpu().numpy()
em = (np.abs(mean - rm) / (1e-3 * np.maximum(np.abs(rm), 1.0))).max()
ev = (np.abs(var - rv) / (1e-3 * np.abs(rv))).max()
print(f"predictive {N} x {C}: mean excess {em:.3f}, var excess {ev:.3f} (<= 1 passes)")
assert em <= 1 and ev <= 1
from bayesvlm_b200.vlm import probit_softmax
sl = slice(N - 4096, N)
pr = probit_softmax(out.mean[sl], out.var[sl]).double().cpu().numpy()
rp = O.probit_softmax(out.mean[sl].double().cpu().numpy(), out.var[sl].double().cpu().numpy(), dtype=np.float64)
print("probit max abs", np.abs(pr - rp).max()); assert np.abs(pr - rp).max() <= 1e-4
del out; torch.cuda.empty_cache()
# GGN: 65536 targets, 4096 sources (oracle-sized), D = 512
z = torch.randn(65536, 512, generator=g)
X, Y = (z + 1.5 * torch.randn(65536, 512, generator=g))[:4096].contiguous(), z + 1.5 * torch.randn(65536, 512, generator=g)
H = compute_hessian_analytic_InfoNCE(X.cuda(), Y.cuda(), torch.tensor(bench.LS)).double().cpu().numpy()
ref = O.infonce_ggn_collapsed(X[:4096].numpy(), Y.numpy(), bench.LS)
rf, rx = np.linalg.norm(H - ref) / np.linalg.norm(ref), np.abs(H - ref).max() / np.abs(ref).max()
print(f"GGN 4096 x 65536: frob {rf:.2e} max {rx:.2e}"); assert rf <= 1e-3 and rx <= 1e-3
print("ok")
