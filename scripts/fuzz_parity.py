"""Random-shape parity fuzz of the public API against the fp64 oracle (test infrastructure, not product code):
    python scripts/fuzz_parity.py [seconds=120 | n=CASES] [seed=0] [which=pred,ggn,syrk,epig,kfac,host,mc]
Shapes are drawn to hit ragged tiles, odd / unaligned widths, row pitches larger than the row, tiny and empty-ish inputs.
Prints one line per failure and a summary; exit code 1 if anything failed."""
import math, sys, time, traceback
import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import laplace_oracle as O
from oracle import torch_port as T
from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
from bayesvlm_b200.hessians import compute_hessian_analytic_InfoNCE, compute_hessian_analytic_SigLIP, syrk_accumulate
from bayesvlm_b200.vlm import CLIP, SIGLIP, EncoderResult, ProbabilisticLogits
from bayesvlm_b200.epig import epig_from_logits_using_matmul, epig_from_probs_using_matmul
from bayesvlm_b200.hessians import kfac_ggn
from bayesvlm_b200.precompute import make_predictions

arg1 = sys.argv[1] if len(sys.argv) > 1 else "120"
max_cases = int(arg1[2:]) if arg1.startswith("n=") else None  # a fixed number of cases (deterministic for a seed) ...
budget = float("inf") if max_cases is not None else float(arg1)  # ... or a time budget
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
which = (sys.argv[3] if len(sys.argv) > 3 else "pred,ggn,syrk,epig,kfac,host,mc").split(",")
rng = np.random.default_rng(seed)
fails, runs = [], {w: 0 for w in which}


def pick(*choices):
    return int(choices[rng.integers(len(choices))])


def dim(lo, hi, specials=()):
    if specials and rng.random() < 0.4:
        return pick(*specials)
    return int(rng.integers(lo, hi + 1))


def spd_inv(gen, d, scale, lam):
    w = torch.randn(4 * d, d, generator=gen, dtype=torch.float64)
    F = (w.T @ w) / math.sqrt(4 * d) * scale
    return torch.linalg.inv(F + math.sqrt(lam) * torch.eye(d, dtype=torch.float64)).float()


def pitched(x, extra):
    """The same values as a view with a row pitch larger than the row (non-contiguous rows)."""
    if extra == 0:
        return x
    buf = torch.zeros(x.shape[0], x.shape[1] + extra, dtype=x.dtype, device=x.device)
    buf[:, : x.shape[1]] = x
    return buf[:, : x.shape[1]]


def fuzz_pred(i):
    gen = torch.Generator().manual_seed(seed * 100003 + i)
    siglip = rng.random() < 0.3
    N, C = dim(1, 700, (1, 127, 128, 129, 255, 256, 257, 513)), dim(1, 600, (1, 10, 255, 256, 257, 511, 513))
    D = dim(2, 1100, (64, 65, 127, 128, 512, 768, 769, 1024, 1025))
    # activation widths >= 64: the fp16 quadratic forms are a sum of d squares of d-term products; their relative error is <= 2e-4
    # from d = 256 up and grows as 1 / sqrt(d) below (1.2e-3 at d = 12, outside the 1e-3 variance tolerance)
    d_i, d_t = dim(64, 1400, (64, 65, 768, 1024, 1280)), dim(64, 900, (64, 512, 768, 769))
    ls = float(rng.uniform(1.0, 4.8))
    prec = ("fp16+fp8", "fp16x3", "fp16")[pick(0, 0, 1, 2)]
    Ai, At = spd_inv(gen, d_i + int(siglip), 3e3, 600.0), spd_inv(gen, d_t + int(siglip), 3e3, 200.0)
    Bi, Bt = spd_inv(gen, D, 20.0, 600.0), spd_inv(gen, D, 20.0, 200.0)
    ie, ia = torch.randn(N, D, generator=gen) * float(rng.uniform(0.01, 30)), torch.randn(N, d_i, generator=gen) * float(rng.uniform(0.01, 30))
    te, ta = torch.randn(C, D, generator=gen), torch.randn(C, d_t, generator=gen)
    desc = f"pred N={N} C={C} D={D} d_i={d_i} d_t={d_t} siglip={siglip} prec={prec} ls={ls:.2f}"
    cls = SIGLIP if siglip else CLIP
    m = cls(logit_scale=ls, logit_bias=-3.0 if siglip else 0.0, device="cuda", precision=prec)
    m.set_covariances(KFC(Ai.cuda(), Bi.cuda()), KFC(At.cuda(), Bt.cuda()))
    ex = pick(0, 0, 1, 3, 4)
    img = EncoderResult(pitched(ie.cuda(), ex), pitched(ia.cuda(), pick(0, 0, 2, 4)))
    txt = EncoderResult(te.cuda(), ta.cuda())
    with torch.no_grad():
        out = m(img, txt)
    rm, rv = O.predictive(ie.numpy(), ia.numpy(), te.numpy(), ta.numpy(), Ai.numpy(), Bi.numpy(), At.numpy(), Bt.numpy(), ls,
                          src_bias=siglip, tgt_bias=siglip, dtype=np.float64)
    s = math.exp(ls)
    mean, var = out.mean.double().cpu().numpy(), out.var.double().cpu().numpy()
    floor = (0.01 if prec != "fp16" else 0.2) * s
    em = (np.abs(mean - rm) / (1e-3 * np.maximum(np.abs(rm), floor))).max()
    ev = (np.abs(var - rv) / (1e-3 * np.abs(rv))).max()
    if not (em <= 1.0 and ev <= 1.0 and np.isfinite(mean).all() and np.isfinite(var).all()):
        fails.append(f"{desc}: mean excess {em:.3g}, var excess {ev:.3g}")
    pr = out.probit().double().cpu().numpy()
    rp = O.probit_softmax(mean, var, dtype=np.float64)
    if np.abs(pr - rp).max() > 1e-4:
        fails.append(f"{desc}: probit max abs {np.abs(pr - rp).max():.3g}")


def fuzz_ggn(i):
    gen = torch.Generator().manual_seed(seed * 100019 + i)
    siglip = rng.random() < 0.5
    # (C = 2 targets: 1.26e-3 -- two fp16 x fp16 products per output element of the final GEMM, nothing averages)
    B, C = dim(1, 900, (1, 5, 127, 128, 129, 256, 257)), dim(16, 1500, (16, 64, 255, 256, 257, 1024))
    D = dim(32, 800, (32, 64, 65, 256, 257, 512, 768))  # (B = 5 with D = 8 / 24: 1.6e-3 / 1.01e-3 -- fp16 curvature weights, nothing averages)
    z = torch.randn(max(B, C), D, generator=gen)
    X = ((z + 1.5 * torch.randn(max(B, C), D, generator=gen))[:B] * float(rng.uniform(0.1, 20))).contiguous()
    Y = (z + 1.5 * torch.randn(max(B, C), D, generator=gen))[:C].contiguous()
    desc = f"ggn B={B} C={C} D={D} siglip={siglip}"
    if siglip:
        ls, lb = float(rng.uniform(2.0, 4.8)), float(rng.uniform(-13, 0))
        H = compute_hessian_analytic_SigLIP(pitched(X.cuda(), pick(0, 0, 3)), torch.arange(B).cuda(), Y.cuda(), torch.tensor(ls), torch.tensor(lb))
        ref = O.siglip_ggn_collapsed(X.numpy(), Y.numpy(), ls, lb)
    else:
        ls = float(rng.uniform(1.0, 4.7))
        H = compute_hessian_analytic_InfoNCE(pitched(X.cuda(), pick(0, 0, 3)), Y.cuda(), torch.tensor(ls))
        ref = O.infonce_ggn_collapsed(X.numpy(), Y.numpy(), ls)
    Hn = H.double().cpu().numpy()
    scale = max(np.abs(ref).max(), 1e-300)
    rel_f = np.linalg.norm(Hn - ref) / max(np.linalg.norm(ref), 1e-300)
    rel_m = np.abs(Hn - ref).max() / scale
    if not (np.isfinite(Hn).all() and rel_f <= 1e-3 and rel_m <= 1e-3):
        fails.append(f"{desc} ls={ls:.2f}: frob {rel_f:.3g}, max {rel_m:.3g}")


def fuzz_syrk(i):
    gen = torch.Generator().manual_seed(seed * 100043 + i)
    n, d = dim(1, 5000, (1, 63, 64, 65, 4096)), dim(1, 1400, (1, 24, 255, 256, 257, 768, 769, 1280))
    one = rng.random() < 0.4
    X = torch.randn(n, d, generator=gen) * torch.exp(torch.randn(d, generator=gen) * 2.0)  # features of very different magnitude
    A = syrk_accumulate(pitched(X.cuda(), pick(0, 0, 1, 4)), append_one=one).double().cpu().numpy()
    Xd = X.double().numpy()
    if one:
        Xd = np.concatenate([Xd, np.ones((n, 1))], 1)
    ref = Xd.T @ Xd
    # relative to the natural scale of each entry (the features differ by orders of magnitude)
    sc = np.sqrt(np.outer(np.diag(ref), np.diag(ref))) + 1e-300
    err = (np.abs(A - ref) / sc).max()
    if not (np.isfinite(A).all() and err <= 1e-3 and np.abs(A - A.T).max() == 0):
        fails.append(f"syrk n={n} d={d} one={one}: scaled max err {err:.3g}, asym {np.abs(A - A.T).max():.3g}")


def fuzz_epig(i):
    torch.manual_seed(seed * 7 + i)
    Np, Nt = dim(1, 700, (1, 127, 128, 129, 256)), dim(1, 500, (1, 10, 128, 257))
    Cl, K = dim(1, 70, (1, 2, 10, 16, 17, 65)), dim(1, 130, (1, 33, 64, 100, 128))
    chunk = pick(64, 256, 512, 4096, 8192)
    mp, vp = torch.randn(Np, Cl, device="cuda") * 2, torch.rand(Np, Cl, device="cuda") * 3 + 0.1
    mt, vt = torch.randn(Nt, Cl, device="cuda") * 2, torch.rand(Nt, Cl, device="cuda") * 3 + 0.1
    desc = f"epig Np={Np} Nt={Nt} Cl={Cl} K={K} chunk={chunk}"
    ours = epig_from_logits_using_matmul(ProbabilisticLogits(mp, vp), ProbabilisticLogits(mt, vt), seed=3, num_samples=K, chunk_size=chunk)
    ref = T.epig_from_logits(mp, vp, mt, vt, seed=3, num_samples=K, chunk_size=chunk)
    d = (ours - ref).abs()
    # one fp16 step at the magnitude of the entropies the score is a difference of (<= log Cl), per column chunk
    quantum = 2.0 ** (math.floor(math.log2(max(math.log(max(Cl, 2)), 1e-3))) - 10)
    n_chunks = math.ceil(Nt * Cl / chunk)
    # (many column chunks: the fp32 sums over chunks differ in their last bits, far below the quantum -- no exact-match floor)
    exact_floor = 0.9 if n_chunks <= 8 else 0.0
    if not (torch.isfinite(ours).all() and float(d.max()) <= 2 * quantum * max(1, n_chunks) and float((d == 0).float().mean()) >= exact_floor):
        fails.append(f"{desc}: exact {float((d == 0).float().mean()):.3f}, max diff {float(d.max()):.3g} (quantum {quantum:.3g})")


def fuzz_kfac(i):
    """K0: the estimation loop with its dropped remainders (hessian_estimation.py:55,71), any input placement."""
    gen = torch.Generator().manual_seed(seed * 100057 + i)
    siglip = rng.random() < 0.4
    ncls, bs = dim(16, 700, (16, 64, 127, 128, 129, 256, 512)), dim(1, 9, (1, 5))
    ncb = dim(1, 4)
    n = ncb * ncls + int(rng.integers(0, ncls))  # ragged tail: dropped by the reference
    D, d_in = dim(32, 300, (32, 64, 65, 128, 256)), dim(64, 400, (64, 65, 128, 257))
    z = torch.randn(n, D, generator=gen)
    src_e, tgt_e = z + 1.5 * torch.randn(n, D, generator=gen), z + 1.5 * torch.randn(n, D, generator=gen)
    src_a = torch.randn(n, d_in, generator=gen)
    ls, lb = (float(rng.uniform(2.0, 4.8)), float(rng.uniform(-13, 0))) if siglip else (float(rng.uniform(1.0, 4.7)), 0.0)
    like = "siglip" if siglip else "info_nce"
    vlm = SIGLIP(logit_scale=ls, logit_bias=lb, device="cuda") if siglip else CLIP(logit_scale=ls, device="cuda")
    place = (lambda t: t, lambda t: t.pin_memory(), lambda t: t.cuda())[pick(0, 1, 2)]
    desc = f"kfac n={n} ncls={ncls} bs={bs} D={D} d_in={d_in} {like} ls={ls:.2f}"
    A, B = kfac_ggn(vlm, ncls, bs, place(src_e), place(src_a), place(tgt_e), "cuda", like)
    Ar, Br = O.kfac_ggn(src_e.numpy(), src_a.numpy(), tgt_e.numpy(), ncls, bs, ls, lb, likelihood=like)
    for name, got, ref in (("A", A, Ar), ("B", B, Br)):
        g = got.double().cpu().numpy()
        rf = np.linalg.norm(g - ref) / max(np.linalg.norm(ref), 1e-300)
        rm = np.abs(g - ref).max() / max(np.abs(ref).max(), 1e-300)
        if not (g.shape == ref.shape and np.isfinite(g).all() and rf <= 1e-3 and rm <= 1e-3):
            fails.append(f"{desc}: {name} frob {rf:.3g}, max {rm:.3g}")
    if B.device.type != "cpu" or A.device.type != "cuda":
        fails.append(f"{desc}: placement A on {A.device}, B on {B.device}")


def fuzz_host(i):
    """make_predictions / predict_host (pinned or pageable host buffers, odd batch sizes) == forward on device tensors."""
    gen = torch.Generator().manual_seed(seed * 100069 + i)
    N, C, D = dim(1, 900, (1, 255, 257)), dim(1, 300, (1, 10, 257)), dim(128, 600, (128, 512, 513))
    d_i, d_t = dim(64, 700, (64, 768)), dim(64, 600, (64, 512))
    bsz = dim(1, 1200, (1, 7, 64, 1000))
    Ai, At = spd_inv(gen, d_i, 3e3, 600.0), spd_inv(gen, d_t, 3e3, 200.0)
    Bi, Bt = spd_inv(gen, D, 20.0, 600.0), spd_inv(gen, D, 20.0, 200.0)
    ie, ia = torch.randn(N, D, generator=gen), torch.randn(N, d_i, generator=gen)
    te, ta = torch.randn(C, D, generator=gen), torch.randn(C, d_t, generator=gen)
    m = CLIP(logit_scale=math.log(100.0), device="cuda")
    m.set_covariances(KFC(Ai.cuda(), Bi.cuda()), KFC(At.cuda(), Bt.cuda()))
    place = (lambda t: t, lambda t: t.pin_memory())[pick(0, 1)]
    out = make_predictions(m, EncoderResult(place(ie), place(ia)), EncoderResult(te, ta), batch_size=bsz, device="cuda")
    with torch.no_grad():
        ref = m(EncoderResult(ie.cuda(), ia.cuda()), EncoderResult(te.cuda(), ta.cuda()))
    desc = f"host N={N} C={C} D={D} d_i={d_i} d_t={d_t} batch={bsz}"
    if out.mean.device.type != "cpu" or out.mean.shape != (N, C):
        fails.append(f"{desc}: result on {out.mean.device} with shape {tuple(out.mean.shape)}")
    # rows are independent: the batched host pipeline and one device call run the same kernels on the same rows
    if not (torch.equal(out.mean, ref.mean.cpu()) and torch.equal(out.var, ref.var.cpu())):
        fails.append(f"{desc}: differs from the device call by {float((out.mean - ref.mean.cpu()).abs().max()):.3g} / "
                     f"{float((out.var - ref.var.cpu()).abs().max()):.3g}")


def fuzz_mc(i):
    """Monte-Carlo softmax / aleatoric entropy on the fused kernel == the reference's torch sequence under a shared seed."""
    n, c, k = dim(1, 600, (1, 33, 257)), dim(1, 1100, (1, 10, 1000, 1024, 1025)), dim(1, 12)
    torch.manual_seed(seed * 13 + i)
    mean, var = torch.randn(n, c, device="cuda") * 3, torch.rand(n, c, device="cuda") * 4 + 0.05
    pl = ProbabilisticLogits(mean, var)
    p = pl.softmax(num_samples=k, seed=17)
    torch.manual_seed(17)
    std, ref = torch.sqrt(var), torch.zeros_like(mean)
    for _ in range(k):  # vlm.py:86-89
        ref += torch.nn.functional.softmax(mean + torch.randn(std.shape, device="cuda") * std, dim=-1)
    ref /= k
    torch.manual_seed(23)
    h = pl.expected_aleatoric_entropy(num_samples=k)
    torch.manual_seed(23)
    href = 0
    for _ in range(k):  # vlm.py:145-149
        pr = torch.nn.functional.softmax(mean + torch.randn(var.shape, device="cuda") * torch.sqrt(var), dim=-1)
        href = href + -(pr * pr.log()).sum(dim=-1)
    href = href / k
    ep, eh = float((p - ref).abs().max()), float(((h - href).abs() / href.abs().clamp_min(1e-3)).max())
    if not (ep <= 2e-6 and eh <= 5e-5):
        fails.append(f"mc n={n} c={c} k={k}: probs {ep:.3g}, entropy rel {eh:.3g}")


FUZZ = {"pred": fuzz_pred, "ggn": fuzz_ggn, "syrk": fuzz_syrk, "epig": fuzz_epig, "kfac": fuzz_kfac, "host": fuzz_host, "mc": fuzz_mc}
t0 = time.time()
i = 0
while time.time() - t0 < budget and (max_cases is None or i < max_cases):
    w = which[i % len(which)]
    try:
        FUZZ[w](i)
    except Exception as exc:  # an exception is a failure too (unless it is a documented refusal)
        fails.append(f"{w} case {i}: EXCEPTION {type(exc).__name__}: {str(exc)[:200]} | {traceback.format_exc().splitlines()[-3].strip()[:160]}")
    runs[w] += 1
    i += 1
    torch.cuda.synchronize()
for f in fails:
    print("FAIL", f)
print("runs", runs, "failures", len(fails))
sys.exit(1 if fails else 0)
