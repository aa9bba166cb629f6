"""Diagnostic: error magnitudes of the KFAC kernels vs the fp64 oracle (run on the GPU box)."""
import math, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import laplace_oracle as O
from bayesvlm_b200.hessians import _ggn, syrk_accumulate
LS = math.log(100.0)
def paired(gen, n, d, noise=1.5):
    z = torch.randn(n, d, generator=gen)
    return z + noise * torch.randn(n, d, generator=gen), z + noise * torch.randn(n, d, generator=gen)
def err(F, ref):
    F = F.double().cpu().numpy(); ref = np.asarray(ref, np.float64)
    return np.linalg.norm(F - ref) / np.linalg.norm(ref), np.abs(F - ref).max() / np.abs(ref).max()
def ggn64(X, Y, ls, lb, siglip, chunk=4096):
    """fp64 torch evaluation of the naive per-sample definition, chunked over sources (checker only)."""
    X = X.double().cuda(); Y = Y.double().cuda(); s = math.exp(ls)
    Yh = Y / Y.norm(dim=1, keepdim=True); D = X.shape[1]
    H = torch.zeros(D, D, dtype=torch.float64, device="cuda")
    for lo in range(0, X.shape[0], chunk):
        x = X[lo:lo + chunk]; nx = x.norm(dim=1, keepdim=True); xh = x / nx; w = 1 / nx[:, 0] ** 2
        L = xh @ Yh.T
        if siglip:
            sg = torch.sigmoid(L * s + lb); Wt = sg * (1 - sg)
            S1 = (Yh * (w[:, None] * Wt).sum(0)[:, None]).T @ Yh
            u = (Wt * L) @ Yh
        else:
            P = torch.softmax(L * s, dim=1)
            m = P @ Yh
            S1 = (Yh * (w[:, None] * P).sum(0)[:, None]).T @ Yh - (m * w[:, None]).T @ m
            u = (P * L) @ Yh - m * (m * xh).sum(1, keepdim=True)
        a = (u * xh).sum(1)
        H += S1 - (xh * w[:, None]).T @ u - (u * w[:, None]).T @ xh + (xh * (w * a)[:, None]).T @ xh
    return (H * s * s).cpu().numpy()
for (B, C, D) in [(7, 40, 24), (5, 300, 64), (129, 257, 96), (640, 1024, 512), (1000, 2048, 768)]:
    gen = torch.Generator().manual_seed(B * 31 + C + D)
    Xa, Ya = paired(gen, max(B, C), D)
    X, Y = Xa[:B].contiguous(), Ya[:C].contiguous()
    ref = O.infonce_ggn_collapsed(X.numpy(), Y.numpy(), LS)
    print("chk64 %.2g" % err(torch.from_numpy(ggn64(X, Y, LS, 0, False)), ref)[0], end=" ")
    for prec in ("fp16", "fp16x3"):
        H = _ggn(X.cuda(), Y.cuda(), LS, 0.0, False, precision=prec)
        print("infonce", (B, C, D), prec, "fro %.3g max %.3g" % err(H, ref), end=" | ")
    print()
    ref = O.siglip_ggn_collapsed(X.numpy(), Y.numpy(), 4.765, -12.93)
    for prec in ("fp16", "fp16x3"):
        H = _ggn(X.cuda(), Y.cuda(), 4.765, -12.93, True, precision=prec)
        print("siglip ", (B, C, D), prec, "fro %.3g max %.3g" % err(H, ref), end=" | ")
    print(flush=True)
for noise in (1.5, 2.0, 2.5):
    for (n, D) in [(8192, 512), (32768, 512), (32768, 768)]:
        gen = torch.Generator().manual_seed(n + D)
        X, Y = paired(gen, n, D, noise)
        for siglip, ls, lb in ((False, LS, 0.0), (True, 4.765, -12.93)):
            ref = ggn64(X, Y, ls, lb, siglip)
            for prec in ("fp16", "fp16x3"):
                torch.cuda.synchronize(); t0 = time.time()
                H = _ggn(X.cuda(), Y.cuda(), ls, lb, siglip, precision=prec); torch.cuda.synchronize()
                print("noise", noise, "siglip" if siglip else "infonce", (n, D), prec, "fro %.3g max %.3g" % err(H, ref),
                      "t=%.1fms" % ((time.time() - t0) * 1e3), flush=True)
