"""P1/P2/P3 parity: CUDA path (through the Python mirror -> ctypes -> C ABI) against the reference-pinned oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import laplace_oracle as O

pytestmark = pytest.mark.gpu
LS = math.log(100.0)


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _model(tag, g, precision, ls, lb=0.0):
    from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
    from bayesvlm_b200.vlm import CLIP, SIGLIP

    cls = SIGLIP if tag == "siglip" else CLIP
    m = cls(logit_scale=ls, logit_bias=lb, device="cuda", precision=precision)
    m.set_covariances(KFC(_cuda(g["A_img_inv"]), _cuda(g["B_img_inv"])), KFC(_cuda(g["A_txt_inv"]), _cuda(g["B_txt_inv"])))
    return m


def _check_logits(mean, var, ref_mean, ref_var, s, strict):
    """Tolerances (fp32 reference on identical inputs):
    mean: |d| <= 1e-3 * max(|ref|, 0.01 s) element-wise in the split-precision (default) mode;
          fp16 single pass: norm-wise 1e-3 and |d| <= 1e-3 * max(|ref|, 0.2 s);
    var : |d| <= 1e-3 * |ref| element-wise."""
    mean, var = mean.double().cpu().numpy(), var.double().cpu().numpy()
    floor = (0.01 if strict else 0.2) * s
    tol = 1e-3 * np.maximum(np.abs(ref_mean), floor)
    dm = np.abs(mean - ref_mean)
    assert (dm <= tol).all(), f"mean: max excess {(dm / tol).max():.3g}, max abs {dm.max():.3g}"
    assert np.linalg.norm(mean - ref_mean) <= 1e-3 * np.linalg.norm(ref_mean)
    dv = np.abs(var - ref_var) / np.abs(ref_var)
    assert dv.max() <= 1e-3, f"var: max rel {dv.max():.3g}"


@pytest.mark.parametrize("precision", ["fp16x3", "fp16+fp8", "fp16"])
@pytest.mark.parametrize("tag", ["clip", "siglip"])
def test_golden_small(golden, tag, precision):
    from bayesvlm_b200.vlm import EncoderResult

    g = {k[len(f"pred_{tag}_"):]: v for k, v in golden.items() if k.startswith(f"pred_{tag}_")}
    ls, lb = (golden["ggn_siglip_params"] if tag == "siglip" else (LS, 0.0))
    model = _model(tag, g, precision, float(ls), float(lb))
    img = EncoderResult(_cuda(g["img_emb"]), _cuda(g["img_act"]))
    txt = EncoderResult(_cuda(g["txt_emb"]), _cuda(g["txt_act"]))
    with torch.no_grad():
        out = model(img, txt)
        out_map = model(img, txt, map_estimate=True)
    _check_logits(out.mean, out.var, g["mean"], g["var"], math.exp(ls), strict=precision != "fp16")
    np.testing.assert_allclose(out_map.mean.cpu().numpy(), g["map"], rtol=1e-4, atol=1e-4)
    assert (out_map.var == 0).all()
    np.testing.assert_allclose(out.probit().cpu().numpy(), O.probit_softmax(out.mean.cpu().numpy(), out.var.cpu().numpy()),
                               atol=1e-5)
    np.testing.assert_allclose(out.softmax(num_samples=0).cpu().numpy(), g["softmax0"], atol=2e-3)
    # plain tensors -> deterministic logits (vlm.py:710), bias included on the MAP path only
    np.testing.assert_allclose(model(img.embeds, txt.embeds).detach().cpu().numpy(), g["map"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("precision", ["fp16x3", "fp16+fp8", "fp16"])
def test_config1_shipped_b32_factors(golden_b32, precision):
    """BASELINE config 1: shipped CLIP ViT-B-32 factors, seeded 10k x 10 features; golden = reference CLIP.forward."""
    from bayesvlm_b200.hessians import compute_covariances
    from bayesvlm_b200.vlm import CLIP, EncoderResult

    b = golden_b32
    info = dict(zip(("n_img", "n_txt", "lambda_img", "lambda_txt"), (float(v) for v in b["info"])))
    cov_img, cov_txt = compute_covariances(_cuda(b["A_img"]), _cuda(b["B_img"]), _cuda(b["A_txt"]), _cuda(b["B_txt"]), info)
    gen = torch.Generator().manual_seed(int(b["seed"][0]))
    rn = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float32).cuda()
    img = EncoderResult(rn(10000, 512), rn(10000, 768))
    txt = EncoderResult(rn(10, 512), rn(10, 512))
    model = CLIP(logit_scale=LS, device="cuda", precision=precision)
    model.set_covariances(cov_img, cov_txt)
    with torch.no_grad():
        out = model(img, txt)
    _check_logits(out.mean, out.var, b["mean"].astype(np.float64), b["var"].astype(np.float64), 100.0,
                  strict=precision != "fp16")


@pytest.mark.parametrize("precision", ["fp16x3", "fp16+fp8"])
@pytest.mark.parametrize("which", ["l14", "siglip"])
def test_shipped_l14_and_siglip_factors(which, precision):
    """BASELINE configs 3 / 5 on the factors the reference ships for ViT-L-14 (A_txt, B_img, B_txt) and SigLIP (A_txt 769^2,
    B_img, B_txt): the real three-decade spectra go through the fp16 quadratic forms.  Golden = the reference's own
    CLIP / SIGLIP.forward on 48 seeded rows (tests/golden/make_golden.py `shipped`); all N rows are predicted (so the golden
    rows sit in different row panels) and the covariances are assembled on the device from the packed factors."""
    import sys

    from conftest import GOLDEN

    sys.path.insert(0, str(GOLDEN))
    from make_golden import shipped_problem, sym_from_lower

    from bayesvlm_b200.hessians import compute_covariances
    from bayesvlm_b200.vlm import CLIP, SIGLIP, EncoderResult

    fx = dict(np.load(GOLDEN / f"shipped_{which}.npz"))
    cfg, t = shipped_problem(which)
    fac = {k: _cuda(sym_from_lower(fx[k + "_tril"], cfg["D"] + (cfg["bias"] if k == "A_txt" else 0))) for k in ("A_txt", "B_img", "B_txt")}
    info = dict(zip(("n_img", "n_txt", "lambda_img", "lambda_txt"), (float(v) for v in fx["info"])))
    cov_img, cov_txt = compute_covariances(t["A_img"].cuda(), fac["B_img"], fac["A_txt"], fac["B_txt"], info)
    cls = SIGLIP if cfg["bias"] else CLIP
    model = cls(logit_scale=cfg["logit_scale"], logit_bias=cfg["logit_bias"], device="cuda", precision=precision)
    model.set_covariances(cov_img, cov_txt)
    with torch.no_grad():
        out = model(EncoderResult(t["img_e"].cuda(), t["img_a"].cuda()), EncoderResult(t["txt_e"].cuda(), t["txt_a"].cuda()))
    rows = torch.from_numpy(fx["rows"]).cuda()
    _check_logits(out.mean[rows], out.var[rows], fx["mean"].astype(np.float64), fx["var"].astype(np.float64),
                  math.exp(cfg["logit_scale"]), strict=True)


def _surrogate_spd(gen, d, scale):
    w = torch.randn(4 * d, d, generator=gen, dtype=torch.float64)
    return ((w.T @ w) / math.sqrt(4 * d) * scale).float()


@pytest.mark.parametrize("precision", ["fp16x3", "fp16+fp8"])
@pytest.mark.parametrize("cfg", [dict(N=50000, C=1000, D=768, d_img=1024, d_txt=768, bias=False, seed=3001),
                                 dict(N=6000, C=1000, D=1024, d_img=1280, d_txt=1024, bias=False, seed=4001),
                                 dict(N=3000, C=257, D=768, d_img=3072, d_txt=768, bias=True, seed=5001)])
def test_full_size_rows_vs_oracle_and_row_independence(cfg, precision):
    """BASELINE configs 3/4/5 shapes (L-14 at the full 50k x 1000): a random row subset is checked against the fp64
    oracle, and rows are independent: predicting the subset alone reproduces the same rows bit for bit."""
    from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
    from bayesvlm_b200.vlm import CLIP, SIGLIP, EncoderResult

    gen = torch.Generator().manual_seed(cfg["seed"])
    N, C, D, bias = cfg["N"], cfg["C"], cfg["D"], int(cfg["bias"])
    A_img = _surrogate_spd(gen, cfg["d_img"] + bias, 3e3)
    A_txt = _surrogate_spd(gen, cfg["d_txt"] + bias, 3e3)
    B_img, B_txt = _surrogate_spd(gen, D, 20.0), _surrogate_spd(gen, D, 20.0)
    lam_i, lam_t = 605.255, 220.124
    inv = lambda F, lam: torch.linalg.inv(F.double() + math.sqrt(lam) * torch.eye(F.shape[0], dtype=torch.float64)).float()
    covs = [inv(A_img, lam_i), inv(B_img, lam_i), inv(A_txt, lam_t), inv(B_txt, lam_t)]
    rn = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float32)
    img_e, img_a, txt_e, txt_a = rn(N, D), rn(N, cfg["d_img"]), rn(C, D), rn(C, cfg["d_txt"])
    ls, lb = (4.765, -12.93) if bias else (LS, 0.0)
    model = (SIGLIP if bias else CLIP)(logit_scale=ls, logit_bias=lb, device="cuda", precision=precision)
    model.set_covariances(KFC(covs[0].cuda(), covs[1].cuda()), KFC(covs[2].cuda(), covs[3].cuda()))
    img = EncoderResult(img_e.cuda(), img_a.cuda())
    txt = EncoderResult(txt_e.cuda(), txt_a.cuda())
    with torch.no_grad():
        out = model(img, txt)
    assert torch.isfinite(out.mean).all() and torch.isfinite(out.var).all() and (out.var > 0).all()
    rows = torch.randperm(N, generator=gen)[:384].sort().values
    rm, rv = O.predictive(img_e[rows].numpy(), img_a[rows].numpy(), txt_e.numpy(), txt_a.numpy(), *(c.numpy() for c in covs),
                          ls, bool(bias), bool(bias), dtype=np.float64)
    _check_logits(out.mean[rows.cuda()], out.var[rows.cuda()], rm, rv, math.exp(ls), strict=True)
    with torch.no_grad():
        sub = model(img[rows.cuda()], txt)
    assert torch.equal(sub.mean, out.mean[rows.cuda()]) and torch.equal(sub.var, out.var[rows.cuda()])
    # probit softmax: rows sum to one, matches the oracle on the subset
    pr = out.probit()
    assert (pr.sum(-1) - 1).abs().max().item() < 1e-5
    np.testing.assert_allclose(pr[rows.cuda()].cpu().numpy(),
                               O.probit_softmax(out.mean[rows.cuda()].cpu().numpy(), out.var[rows.cuda()].cpu().numpy()),
                               atol=1e-5)


def test_edge_cases_and_errors(golden):
    from bayesvlm_b200.vlm import EncoderResult, ProbabilisticLogits

    g = {k[len("pred_clip_"):]: v for k, v in golden.items() if k.startswith("pred_clip_")}
    model = _model("clip", g, "fp16x3", LS)
    txt = EncoderResult(_cuda(g["txt_emb"]), _cuda(g["txt_act"]))
    img = EncoderResult(_cuda(g["img_emb"]), _cuda(g["img_act"]))
    with torch.no_grad():
        empty = model(EncoderResult(img.embeds[:0], img.activations[:0]), txt)
        one = model(EncoderResult(img.embeds[:1], img.activations[:1]), EncoderResult(txt.embeds[:1], txt.activations[:1]))
        strided = model(EncoderResult(torch.cat([img.embeds, img.embeds], 1)[:, : img.embeds.shape[1]], img.activations), txt)
        full = model(img, txt)
    assert empty.mean.shape == (0, txt.embeds.shape[0])
    np.testing.assert_allclose(one.mean.cpu().numpy(), g["mean"][:1, :1], rtol=1e-3)
    assert torch.equal(strided.mean, full.mean)
    with pytest.raises(NotImplementedError):
        model._compute_probabilistic_logits_smith(img, txt, compute_covariance=True)
    with pytest.raises(RuntimeError):  # CPU tensors: no CPU fallback
        model(EncoderResult(img.embeds.cpu(), img.activations.cpu()), txt)
    with pytest.raises(ValueError):
        ProbabilisticLogits(torch.zeros(3), torch.zeros(3)).sample_probas(2)
    # autograd: the N = 1 case of the online EPIG loop stays differentiable (reference epig.py:214-227)
    e = img.embeds[:1].clone().requires_grad_(True)
    out = model(EncoderResult(e, img.activations[:1]), txt)
    out.mean.sum().backward()
    assert e.grad is not None and torch.isfinite(e.grad).all()
    np.testing.assert_allclose(out.mean.detach().cpu().numpy(), g["mean"][:1], rtol=1e-4, atol=1e-3)


def test_predict_host_end_to_end(golden):
    """Host-buffer path (make_predictions data flow): pinned H2D per batch, kernels, D2H of mean/var."""
    from bayesvlm_b200.precompute import make_predictions
    from bayesvlm_b200.vlm import EncoderResult

    g = {k[len("pred_clip_"):]: v for k, v in golden.items() if k.startswith("pred_clip_")}
    model = _model("clip", g, "fp16x3", LS)
    img = EncoderResult(torch.from_numpy(g["img_emb"]), torch.from_numpy(g["img_act"]))
    txt = EncoderResult(torch.from_numpy(g["txt_emb"]), torch.from_numpy(g["txt_act"]))
    out = make_predictions(model, img, txt, batch_size=16, device="cuda")
    assert out.mean.device.type == "cpu"
    _check_logits(out.mean, out.var, g["mean"], g["var"], 100.0, strict=True)
