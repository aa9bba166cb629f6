"""The CPU oracle against outputs of the reference implementation itself (tests/golden/make_golden.py)."""
import math
import sys

import numpy as np
import pytest

from oracle import laplace_oracle as O
from conftest import GOLDEN as GOLDEN_DIR
from conftest import relerr

LS = math.log(100.0)


def test_infonce_naive_and_collapsed_match_reference(golden):
    X, Y = golden["ggn_X"], golden["ggn_Y"]
    ref64, ref32 = golden["ggn_infonce_H64"], golden["ggn_infonce_H"]
    assert relerr(O.infonce_ggn_naive(X, Y, LS), ref64) < 1e-12
    assert relerr(O.infonce_ggn_collapsed(X, Y, LS), ref64) < 1e-12
    assert relerr(O.infonce_ggn_collapsed(X, Y, LS, np.float32), ref32) < 5e-6
    H = O.infonce_ggn_collapsed(X, Y, LS)
    assert np.abs(H - H.T).max() <= 1e-12 * np.abs(H).max()


def test_siglip_naive_and_collapsed_match_reference(golden):
    X, Y = golden["ggn_X"], golden["ggn_Y"]
    s, b = golden["ggn_siglip_params"]
    ref64, ref32 = golden["ggn_siglip_H64"], golden["ggn_siglip_H"]
    idx = np.arange(X.shape[0])
    assert relerr(O.siglip_ggn_naive(X, idx, Y, s, b), ref64) < 1e-12
    assert relerr(O.siglip_ggn_collapsed(X, Y, s, b), ref64) < 1e-12
    assert relerr(O.siglip_ggn_collapsed(X, Y, s, b, np.float32), ref32) < 5e-6
    # labels only flip the sign inside sigma(1-sigma): result independent of indices_batch
    assert relerr(O.siglip_ggn_naive(X, idx[::-1].copy(), Y, s, b), ref64) < 1e-12


def test_siglip_dim_mismatch_asserts(golden):
    with pytest.raises(AssertionError):
        O.siglip_ggn_naive(golden["ggn_X"], np.arange(7), golden["ggn_Y"][:, :-1], 1.0, 0.0)


@pytest.mark.parametrize("likelihood", ["info_nce", "siglip"])
@pytest.mark.parametrize("literal", [False, True])
def test_kfac_loop_quirks(golden, likelihood, literal):
    ncls, bs = (int(v) for v in golden["kfac_cfg"])
    if likelihood == "info_nce":
        ls, lb, tag = LS, 0.0, "infonce"
    else:
        (ls, lb), tag = golden["ggn_siglip_params"], "siglip"
    A, B = O.kfac_ggn(golden["kfac_emb_s"], golden["kfac_act_s"], golden["kfac_emb_t"], ncls, bs, ls, lb, likelihood,
                      literal_batches=literal)
    assert relerr(A, golden[f"kfac_{tag}_A"]) < 1e-6
    assert relerr(B, golden[f"kfac_{tag}_B"]) < 1e-5
    if likelihood == "siglip":  # ones column appended: (d_in+1)^2 and A[-1,-1]*sqrt(n) == n
        n = (len(golden["kfac_emb_t"]) // ncls) * ncls
        assert A.shape[0] == golden["kfac_act_s"].shape[1] + 1
        assert abs(A[-1, -1] * math.sqrt(n) - n) < 1e-9


def test_kfac_data_batch_remainder_is_load_bearing(golden):
    """Keeping the last num_classes % batch_size sources changes B by tens of percent (quirk ii of K0)."""
    ncls, bs = (int(v) for v in golden["kfac_cfg"])
    _, B_keep = O.kfac_ggn(golden["kfac_emb_s"], golden["kfac_act_s"], golden["kfac_emb_t"], ncls, 1, LS)
    assert relerr(B_keep, golden["kfac_infonce_B"]) > 1e-2


def test_kfac_errors():
    x = np.zeros((10, 4))
    with pytest.raises(ValueError):
        O.kfac_ggn(x, x, x, 64, 5, 0.0)
    with pytest.raises(ValueError):
        O.kfac_ggn(x, x, x, 5, 5, 0.0, likelihood="bogus")


@pytest.mark.parametrize("tag", ["clip", "siglip"])
def test_covariance_and_predictive(golden, tag):
    g = {k[len(f"pred_{tag}_"):]: v for k, v in golden.items() if k.startswith(f"pred_{tag}_")}
    n_img, n_txt, l_img, l_txt = golden["pred_info"]
    Ai, Bi = O.covariance(g["A_img"], g["B_img"], n_img, l_img)
    At, Bt = O.covariance(g["A_txt"], g["B_txt"], n_txt, l_txt)
    assert relerr(Ai, g["A_img_inv"]) < 1e-4 and relerr(Bi, g["B_img_inv"]) < 1e-4
    assert relerr(At, g["A_txt_inv"]) < 1e-4 and relerr(Bt, g["B_txt_inv"]) < 1e-4
    bias = tag == "siglip"
    ls = golden["ggn_siglip_params"][0] if bias else LS
    lb = golden["ggn_siglip_params"][1] if bias else 0.0
    mean, var = O.predictive(g["img_emb"], g["img_act"], g["txt_emb"], g["txt_act"], g["A_img_inv"], g["B_img_inv"],
                             g["A_txt_inv"], g["B_txt_inv"], ls, bias, bias)
    np.testing.assert_allclose(mean, g["mean"], rtol=2e-5, atol=2e-5 * np.abs(g["mean"]).max())
    np.testing.assert_allclose(var, g["var"], rtol=2e-5)
    np.testing.assert_allclose(O.map_logits(g["img_emb"], g["txt_emb"], ls, lb), g["map"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(O.probit_softmax(g["mean"], g["var"]), g["probit"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(O.probit_softmax_method_quirk(g["mean"], g["var"]), g["softmax0"], rtol=1e-5, atol=1e-7)
    # the method's diagonal quirk is NOT the canonical probit
    if tag == "clip":  # (with SigLIP's small logit scale the variances are tiny and the two nearly coincide)
        assert np.abs(g["softmax0"] - g["probit"]).max() > 1e-4


def test_predictive_config1_shipped_b32_factors(golden_b32):
    """Config 1 (CPU-runnable): shipped CLIP ViT-B-32 factors, seeded 10k x 10 synthetic features."""
    import torch

    b = golden_b32
    n_img, n_txt, l_img, l_txt = b["info"]
    Ai, Bi = O.covariance(b["A_img"], b["B_img"], n_img, l_img)
    At, Bt = O.covariance(b["A_txt"], b["B_txt"], n_txt, l_txt)
    g = torch.Generator().manual_seed(int(b["seed"][0]))
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32).numpy()
    img_e, img_a, txt_e, txt_a = rn(10000, 512), rn(10000, 768), rn(10, 512), rn(10, 512)
    mean, var = O.predictive(img_e, img_a, txt_e, txt_a, Ai, Bi, At, Bt, LS, dtype=np.float64)
    assert np.abs(mean - b["mean"]).max() <= 1e-3 * max(np.abs(b["mean"]).max(), 1.0) * 0.1
    np.testing.assert_allclose(var, b["var"], rtol=1e-3)


@pytest.mark.parametrize("which", ["l14", "siglip"])
def test_predictive_on_shipped_l14_and_siglip_factors(which):
    """The factors the reference SHIPS for ViT-L-14 (A_txt, B_img, B_txt) and SigLIP (A_txt 769^2, B_img, B_txt): three-decade
    spectra, d = 768 / 769.  Golden = the reference's CLIP / SIGLIP.forward on 48 seeded rows (tests/golden/make_golden.py
    `shipped`); the oracle must reproduce it from the same packed factors."""
    sys.path.insert(0, str(GOLDEN_DIR))
    from make_golden import shipped_problem, sym_from_lower

    fx = dict(np.load(GOLDEN_DIR / f"shipped_{which}.npz"))
    cfg, t = shipped_problem(which)
    n_img, n_txt, l_img, l_txt = fx["info"]
    fac = {k: sym_from_lower(fx[k + "_tril"], cfg["D"] + (cfg["bias"] if k == "A_txt" else 0)) for k in ("A_txt", "B_img", "B_txt")}
    Ai, Bi = O.covariance(t["A_img"].numpy(), fac["B_img"], n_img, l_img)
    At, Bt = O.covariance(fac["A_txt"], fac["B_txt"], n_txt, l_txt)
    rows = fx["rows"]
    assert (rows == t["rows"].numpy()).all()
    mean, var = O.predictive(t["img_e"].numpy()[rows], t["img_a"].numpy()[rows], t["txt_e"].numpy(), t["txt_a"].numpy(), Ai, Bi,
                             At, Bt, cfg["logit_scale"], src_bias=bool(cfg["bias"]), tgt_bias=bool(cfg["bias"]), dtype=np.float64)
    s = math.exp(cfg["logit_scale"])
    assert np.abs(mean - fx["mean"]).max() <= 1e-4 * max(np.abs(fx["mean"]).max(), 0.01 * s)
    np.testing.assert_allclose(var, fx["var"], rtol=2e-4)
    # the spectra the fp16 quadratic forms have to survive: more than two decades between the extreme eigenvalues
    ev = np.linalg.eigvalsh(fac["A_txt"].astype(np.float64))
    assert ev.max() / max(ev.min(), 1e-30) > 500


def test_epig_fp16_rounding_points(golden):
    p32, t32 = golden["epig_probs_p"], golden["epig_probs_t"]
    np.testing.assert_allclose(O.sample_probas(golden["epig_mean_p"], golden["epig_var_p"], golden["epig_eps_p"]), p32,
                               rtol=2e-6, atol=1e-7)
    p16, t16 = p32.astype(np.float16), t32.astype(np.float16)
    chunk = int(golden["epig_cfg"][0])
    me = O.marginal_entropy_f16(p16)
    assert me.dtype == np.float16
    ref_me = golden["epig_marginal_p16"]
    ulp = np.abs(me.astype(np.float32) - ref_me.astype(np.float32)) / np.spacing(np.abs(ref_me)).astype(np.float32)
    assert (ulp == 0).mean() >= 0.95 and ulp.max() <= 1
    scores = O.epig_from_probs_f16(p16, t16, chunk)
    ref = golden["epig_scores_f16"]
    assert scores.dtype == np.float32
    # identical up to rare one-fp16-ulp flips caused by the fp32 summation order inside the reductions
    diff = np.abs(scores - ref)
    assert (diff == 0).mean() >= 0.8, (diff == 0).mean()
    assert diff.max() <= 2 * 2.0 ** -10 * np.abs(golden["epig_marginal_p16"].astype(np.float32)).max()
    # noise-free definition agrees with the reference evaluated on fp32 probabilities
    np.testing.assert_allclose(O.epig_from_probs_f32(p32, t32), golden["epig_scores_f32"], atol=5e-6)


def test_log_marglik_objective_is_maximised_by_reference_lambda(golden):
    lam0, n, _, _ = golden["prior_cfg"]
    lam = float(golden["prior_lambda"][0])
    W = golden["prior_W"]
    f = lambda l: O.log_marglik(golden["prior_A"], golden["prior_B"], n, l, float((W ** 2).sum()), W.size)
    assert f(lam) > f(lam0)  # 60 Adam ascent steps improved the objective the oracle restates


# ----------------------------------------------------------------------------------------------------------------------
# oracle/torch_port.py: the torch-CPU restatement bench.py times as the reference arm, against the same goldens
# ----------------------------------------------------------------------------------------------------------------------
def test_torch_port_matches_reference_outputs(golden):
    import torch

    from oracle import torch_port as T

    t = lambda k: torch.from_numpy(golden[k])
    X, Y = t("ggn_X"), t("ggn_Y")
    np.testing.assert_allclose(T.infonce_ggn(X, Y, LS).numpy(), golden["ggn_infonce_H"], rtol=2e-4, atol=2e-4)
    s, b = golden["ggn_siglip_params"]
    np.testing.assert_allclose(T.siglip_ggn(X, torch.arange(7), Y, s, b, chunk_size_j=16).numpy(), golden["ggn_siglip_H"],
                               rtol=2e-4, atol=1e-5)
    ncls, bs = (int(v) for v in golden["kfac_cfg"])
    A, B = T.kfac_ggn(t("kfac_emb_s"), t("kfac_act_s"), t("kfac_emb_t"), ncls, bs, LS)
    assert relerr(A.numpy(), golden["kfac_infonce_A"]) < 1e-6 and relerr(B.numpy(), golden["kfac_infonce_B"]) < 1e-4
    A, B = T.kfac_ggn(t("kfac_emb_s"), t("kfac_act_s"), t("kfac_emb_t"), ncls, bs, s, b, "siglip", siglip_chunk_size_j=20)
    assert relerr(A.numpy(), golden["kfac_siglip_A"]) < 1e-6 and relerr(B.numpy(), golden["kfac_siglip_B"]) < 1e-4
    for tag, bias, ls in (("clip", False, LS), ("siglip", True, s)):
        g = {k[len(f"pred_{tag}_"):]: torch.from_numpy(v) for k, v in golden.items() if k.startswith(f"pred_{tag}_")}
        mean, var = T.predictive(g["img_emb"], g["img_act"], g["txt_emb"], g["txt_act"], g["A_img_inv"], g["B_img_inv"],
                                 g["A_txt_inv"], g["B_txt_inv"], ls, bias, bias)
        np.testing.assert_allclose(mean.numpy(), g["mean"].numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(var.numpy(), g["var"].numpy(), rtol=1e-5)
        np.testing.assert_allclose(T.probit_softmax(mean, var).numpy(), g["probit"].numpy(), rtol=1e-5, atol=1e-7)
    p16, t16 = t("epig_probs_p").half(), t("epig_probs_t").half()
    chunk = int(golden["epig_cfg"][0])
    sc = T.epig_from_probs(p16, t16, chunk_size=chunk)
    assert np.abs(sc.float().numpy() - golden["epig_scores_f16"]).max() <= 2.0 ** -9
    np.testing.assert_allclose(T.epig_from_probs(t("epig_probs_p"), t("epig_probs_t"), chunk).numpy(),
                               golden["epig_scores_f32"], atol=5e-6)


def test_product_prior_precision_matches_reference(golden):
    """bayesvlm_b200.hessians.optimize_prior_precision (torch, eigenvalue form) against the reference's Adam result."""
    import torch

    from bayesvlm_b200.hessians import compute_log_det_kfac, optimize_prior_precision

    lam0, n, lr, steps = golden["prior_cfg"]
    proj = torch.nn.Linear(24, 16, bias=False)
    with torch.no_grad():
        proj.weight.copy_(torch.from_numpy(golden["prior_W"]))
    A, B = torch.from_numpy(golden["prior_A"]), torch.from_numpy(golden["prior_B"])
    lam = optimize_prior_precision(proj, A, B, lmbda_init=float(lam0), n=float(n), lr=float(lr), num_steps=int(steps),
                                   device="cpu")
    assert abs(lam.item() - float(golden["prior_lambda"][0])) <= 2e-3 * float(golden["prior_lambda"][0])
    # the reference's log-det quirk: p * logdet(A) + q * logdet(B)
    ld = compute_log_det_kfac(A + torch.eye(24), B + torch.eye(16))
    ref = torch.logdet(A + torch.eye(24)) * 24 + torch.logdet(B + torch.eye(16)) * 16
    assert torch.allclose(ld, ref)


def test_factor_spectrum_feeds_prior_precision_and_covariance(golden):
    """§8(f)#1: ONE eigendecomposition per factor gives the reference's Adam result for lambda (hessians.py:219-265) and the
    reference's regularised inverses (hessians.py:170-184)."""
    import torch

    from bayesvlm_b200.hessians import (FactorSpectrum, _compute_covariance, covariance_from_spectra,
                                        optimize_prior_precision)

    lam0, n, lr, steps = golden["prior_cfg"]
    proj = torch.nn.Linear(24, 16, bias=False)
    with torch.no_grad():
        proj.weight.copy_(torch.from_numpy(golden["prior_W"]))
    A, B = torch.from_numpy(golden["prior_A"]), torch.from_numpy(golden["prior_B"])
    spectra = (FactorSpectrum.of(A), FactorSpectrum.of(B))
    lam = optimize_prior_precision(proj, A, B, lmbda_init=float(lam0), n=float(n), lr=float(lr), num_steps=int(steps),
                                   device="cpu", spectra=spectra)
    assert abs(lam.item() - float(golden["prior_lambda"][0])) <= 2e-3 * float(golden["prior_lambda"][0])
    cov = covariance_from_spectra(spectra[0], spectra[1], float(n), lam.item())
    ref = _compute_covariance(A, B, torch.tensor(float(n)), torch.tensor(lam.item()))  # torch.linalg.inv, as the reference
    for ours, theirs in ((cov.A_inv, ref.A_inv), (cov.B_inv, ref.B_inv)):
        assert ours.dtype == torch.float32
        assert (ours - theirs).abs().max() <= 1e-5 * theirs.abs().max()


# ---------------------------------------------------------------------------------------------------------------------
# EPIG: CPU vs CUDA Half semantics of torch, the fused-chunk golden and the online loop
# ---------------------------------------------------------------------------------------------------------------------
def test_xlogy_half_semantics_of_both_torch_backends():
    """The oracle's two xlogy modes against tables recorded from torch itself: `xlogy_cpu_bits` by torch's CPU kernel and
    `xlogy_cuda_bits` by torch's CUDA kernel ON A B200 (scripts/probe_epig_parity.py) for every fp16 value in (0, 1]."""
    g = np.load(GOLDEN_DIR / "torch_cuda_half_semantics.npz")
    x = g["x_bits"].view(np.float16)
    cpu = O._xlogy_f16(x, "cpu").view(np.uint16)
    cuda = O._xlogy_f16(x, "cuda").view(np.uint16)
    assert (cpu == g["xlogy_cpu_bits"]).all()
    assert (cuda == g["xlogy_cuda_bits"]).mean() >= 0.9999  # one value of 15360: logf's last bit
    # the two backends really differ (which is why the CUDA kernels cannot be pinned on CPU goldens bit for bit)
    assert (g["xlogy_cpu_bits"] != g["xlogy_cuda_bits"]).mean() > 0.2
    import torch

    assert (torch.xlogy(torch.from_numpy(x), torch.from_numpy(x)).numpy().view(np.uint16) == g["xlogy_cpu_bits"]).all()


def test_epig_fused_chunk_golden_cpu_semantics():
    """Reference epig_from_probs_using_matmul (CPU) at chunk = 256, the fused kernel's tile width: bit-exact with the oracle's
    CPU mode and with the torch port; the CUDA mode differs by score quanta only."""
    import torch

    from oracle import torch_port as T

    g = np.load(GOLDEN_DIR / "epig_online_small.npz")
    chunk = int(g["fused_chunk"][0])
    s_cpu = O.epig_from_probs_f16(g["fused_p16"], g["fused_t16"], chunk, device="cpu")
    assert np.array_equal(s_cpu, g["fused_scores"])
    assert np.array_equal(O.marginal_entropy_f16(g["fused_p16"], "cpu"), g["fused_marginal"])
    s_port = T.epig_from_probs(torch.from_numpy(g["fused_p16"]), torch.from_numpy(g["fused_t16"]), chunk_size=chunk)
    assert np.array_equal(s_port.float().numpy(), g["fused_scores"])
    s_cuda = O.epig_from_probs_f16(g["fused_p16"], g["fused_t16"], chunk, device="cuda")
    assert np.abs(s_cuda - g["fused_scores"]).max() <= 4 * 2.0 ** -10


def test_torch_port_online_loop_matches_reference():
    """oracle/torch_port.select_epig_online (the checker of the GPU loop test) against the reference's own
    select_epig_online run on the CPU by make_golden.py: same picks, same scores."""
    import importlib.util

    from oracle import torch_port as T

    spec = importlib.util.spec_from_file_location("make_golden", GOLDEN_DIR / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    pr = mg.epig_online_problem()
    c = pr["cfg"]
    g = np.load(GOLDEN_DIR / "epig_online_small.npz")
    sel, sc, _, _ = T.select_epig_online(
        pr["label_e"], pr["label_a"], pr["pool_a"] @ pr["W"].T, pr["pool_a"], pr["targ_a"] @ pr["W"].T, pr["targ_a"], pr["ids"],
        pr["W"], c["logit_scale"], pr["A_img"], pr["A_txt"], pr["B_img"], pr["B_txt"], pr["info"], c["budget"], c["lr"],
        c["hessian_update_scale"], "cpu", c["num_samples"], c["seed"], c["pool_max_size"], c["target_max_size"], c["chunk_size"])
    assert sel == g["selected"].tolist()
    np.testing.assert_allclose(sc, g["scores"], rtol=0, atol=1e-7)
