"""E0/E1/E2/E3 parity: CUDA path against the oracle's restatement of the reference's fp16 rounding points and the
golden outputs of the reference's own epig functions (tests/golden/make_golden.py).

EPIG parity protocol: identical fp16 probabilities go to both sides; scores must agree up to rare
one-fp16-ulp flips of a per-chunk partial sum (fp32 summation order inside the reductions differs between any two
implementations, the reference's own CPU and GPU paths included); top-k sets must be identical modulo ties.
"""
import math

import numpy as np
import pytest
import torch

from oracle import laplace_oracle as O

pytestmark = pytest.mark.gpu


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _probs(gen, n, k, cl, spread=2.0):
    mean = torch.randn(n, cl, generator=gen) * spread
    var = torch.rand(n, cl, generator=gen) * 3 + 0.1
    eps = torch.randn(k, n, cl, generator=gen)
    return mean, var, eps


def _topk_identical_modulo_ties(scores, ref, k, ulp):
    """Any index in one top-k set and not the other must have a reference score within `ulp` of the k-th score."""
    a = set(np.argsort(-scores, kind="stable")[:k].tolist())
    b = set(np.argsort(-ref, kind="stable")[:k].tolist())
    kth = np.sort(ref)[::-1][k - 1]
    for i in a ^ b:
        assert abs(ref[i] - kth) <= ulp, (i, ref[i], kth, ulp)


def test_sample_probs_golden(golden):
    from bayesvlm_b200.vlm import sample_probas_from_noise

    p = sample_probas_from_noise(_cuda(golden["epig_mean_p"]), _cuda(golden["epig_var_p"]), _cuda(golden["epig_eps_p"]))
    assert p.dtype == torch.float16 and tuple(p.shape) == golden["epig_probs_p"].shape
    ref16 = golden["epig_probs_p"].astype(np.float16)
    d = np.abs(p.cpu().numpy().astype(np.float32) - ref16.astype(np.float32))
    # the fp32 softmax before the fp16 rounding may differ in the last fp32 bit -> at most one fp16 ulp, rarely
    assert (d == 0).mean() >= 0.99
    assert (d <= np.spacing(np.abs(ref16)).astype(np.float32)).all()


def test_marginal_entropy_golden(golden):
    from bayesvlm_b200.epig import marginal_entropy_from_probs

    p16 = _cuda(golden["epig_probs_p"].astype(np.float16))
    me = marginal_entropy_from_probs(p16)
    assert me.dtype == torch.float16
    ref = golden["epig_marginal_p16"]
    d = np.abs(me.cpu().numpy().astype(np.float32) - ref.astype(np.float32))
    assert (d <= np.spacing(np.abs(ref)).astype(np.float32)).all()
    assert (d == 0).mean() >= 0.9
    # fp32 input keeps torch's generic path (reference semantics for non-fp16 dtypes)
    me32 = marginal_entropy_from_probs(_cuda(golden["epig_probs_p"]))
    np.testing.assert_allclose(me32.cpu().numpy(), golden["epig_marginal_p32"], rtol=1e-5, atol=1e-6)


def test_epig_scores_golden(golden):
    from bayesvlm_b200.epig import epig_from_probs_using_matmul

    p16 = _cuda(golden["epig_probs_p"].astype(np.float16))
    t16 = _cuda(golden["epig_probs_t"].astype(np.float16))
    chunk = int(golden["epig_cfg"][0])  # 64: not a multiple of the fused kernel's 256-column tile -> generic device path
    s = epig_from_probs_using_matmul(p16, t16, chunk_size=chunk)
    ref = golden["epig_scores_f16"]
    assert np.abs(s.float().cpu().numpy() - ref).max() <= 4e-3
    s32 = epig_from_probs_using_matmul(_cuda(golden["epig_probs_p"]), _cuda(golden["epig_probs_t"]), chunk_size=chunk)
    np.testing.assert_allclose(s32.cpu().numpy(), golden["epig_scores_f32"], atol=2e-5)


@pytest.mark.parametrize("cfg", [dict(Np=45, Nt=30, K=16, Cl=5, chunk=256), dict(Np=300, Nt=200, K=100, Cl=10, chunk=512),
                                 dict(Np=257, Nt=129, K=100, Cl=65, chunk=4096), dict(Np=64, Nt=77, K=33, Cl=128, chunk=1024),
                                 dict(Np=500, Nt=1000, K=100, Cl=10, chunk=4096)])
def test_epig_fused_vs_oracle(cfg):
    """Fused joint-entropy kernel vs the numpy restatement of the reference's rounding points (same fp16 inputs)."""
    from bayesvlm_b200.epig import epig_from_probs_using_matmul
    from bayesvlm_b200.vlm import sample_probas_from_noise

    gen = torch.Generator().manual_seed(cfg["Np"] * 7 + cfg["Cl"])
    mp, vp, ep = _probs(gen, cfg["Np"], cfg["K"], cfg["Cl"])
    mt, vt, et = _probs(gen, cfg["Nt"], cfg["K"], cfg["Cl"])
    p16 = sample_probas_from_noise(mp.cuda(), vp.cuda(), ep.cuda())
    t16 = sample_probas_from_noise(mt.cuda(), vt.cuda(), et.cuda())
    s = epig_from_probs_using_matmul(p16, t16, chunk_size=cfg["chunk"]).float().cpu().numpy()
    ref = O.epig_from_probs_f16(p16.cpu().numpy(), t16.cpu().numpy(), cfg["chunk"])
    assert np.isfinite(s).all()
    n_chunks = math.ceil(cfg["Nt"] * cfg["Cl"] / cfg["chunk"])
    # one fp16 ulp of a per-chunk partial (|H_chunk| <= log(Cl^2) * chunk_fraction) per chunk, plus one for H_pool
    h_max = 2 * math.log(cfg["Cl"])
    ulp = 2.0 ** -10 * max(h_max / n_chunks, 2.0 ** -14) * 2
    d = np.abs(s - ref)
    assert d.max() <= ulp * n_chunks + 2.0 ** -10 * h_max, (d.max(), ulp, n_chunks)
    assert (d == 0).mean() >= 0.5, (d == 0).mean()
    k = min(50, cfg["Np"] // 2)
    _topk_identical_modulo_ties(s, ref, k, ulp * n_chunks + 2.0 ** -10 * h_max)


@pytest.mark.parametrize("cfg", [dict(Np=300, Nt=200, K=100, Cl=10, chunk=512), dict(Np=1000, Nt=700, K=100, Cl=10, chunk=4096),
                                 dict(Np=257, Nt=129, K=64, Cl=65, chunk=4096)])
def test_epig_fused_vs_reference_ops_on_same_gpu(cfg):
    """Parity protocol (1) of SURVEY.md section 8(d): identical fp16 probabilities go to the reference's operation sequence
    executed by torch ON THE SAME GPU (oracle/torch_port.epig_from_probs: fp16 matmul, `/K`, xlogy, sums -- torch's own CUDA
    kernels and their rounding) and to the fused kernel; scores agree to per-chunk fp16 ulps, top-k identical modulo ties."""
    from oracle import torch_port as T

    from bayesvlm_b200.epig import epig_from_probs_using_matmul
    from bayesvlm_b200.vlm import sample_probas_from_noise

    gen = torch.Generator().manual_seed(cfg["Np"] + 13 * cfg["Cl"])
    mp, vp, ep = _probs(gen, cfg["Np"], cfg["K"], cfg["Cl"])
    mt, vt, et = _probs(gen, cfg["Nt"], cfg["K"], cfg["Cl"])
    p16 = sample_probas_from_noise(mp.cuda(), vp.cuda(), ep.cuda())
    t16 = sample_probas_from_noise(mt.cuda(), vt.cuda(), et.cuda())
    s = epig_from_probs_using_matmul(p16, t16, chunk_size=cfg["chunk"]).float().cpu().numpy()
    ref = T.epig_from_probs(p16, t16, chunk_size=cfg["chunk"]).float().cpu().numpy()
    n_chunks = math.ceil(cfg["Nt"] * cfg["Cl"] / cfg["chunk"])
    h_max = 2 * math.log(cfg["Cl"])
    tol = 2.0 ** -10 * max(h_max / n_chunks, 2.0 ** -14) * 2 * n_chunks + 2.0 ** -10 * h_max
    d = np.abs(s - ref)
    assert d.max() <= tol, (d.max(), tol)
    assert (d == 0).mean() >= 0.4, (d == 0).mean()
    _topk_identical_modulo_ties(s, ref, min(50, cfg["Np"] // 2), tol)


def test_epig_from_logits_shared_rng():
    """E3: per-pool-chunk re-seeding (seed + row offset) with torch's own generator on the device."""
    from bayesvlm_b200.epig import epig_from_logits_using_matmul, epig_from_probs_using_matmul
    from bayesvlm_b200.vlm import ProbabilisticLogits

    gen = torch.Generator().manual_seed(9)
    mp, vp, _ = _probs(gen, 700, 1, 10)
    mt, vt, _ = _probs(gen, 300, 1, 10)
    lp = ProbabilisticLogits(mp.cuda(), vp.cuda())
    lt = ProbabilisticLogits(mt.cuda(), vt.cuda())
    s = epig_from_logits_using_matmul(lp, lt, seed=3, num_samples=32, chunk_size=256)
    assert s.shape == (700,) and s.dtype == torch.float32 and torch.isfinite(s).all()
    # chunk 1 (rows 256..511) reproduces from the reference recipe: same seed for target and pool draws
    pt = lt.sample_probas(32, seed=3 + 256).half()
    pp = ProbabilisticLogits(lp.mean[256:512], lp.var[256:512]).sample_probas(32, seed=3 + 256).half()
    s1 = epig_from_probs_using_matmul(pp, pt, chunk_size=256).float()
    assert (s[256:512] - s1).abs().max().item() <= 4e-3
    s_again = epig_from_logits_using_matmul(lp, lt, seed=3, num_samples=32, chunk_size=256)
    assert torch.equal(s, s_again)


def test_epig_errors():
    from bayesvlm_b200.epig import epig_from_probs_using_matmul, marginal_entropy_from_probs

    p = torch.rand(4, 3, 5).softmax(-1)
    with pytest.raises(RuntimeError):
        epig_from_probs_using_matmul(p, p)  # CPU tensors: no fallback
    with pytest.raises(AssertionError):
        marginal_entropy_from_probs(p[0].cuda())


def test_select_epig_online_smoke():
    """The online greedy loop (reference epig.py:44-273) end to end on device: budget picks, no duplicates, finite scores,
    covariances refreshed each step."""
    from bayesvlm_b200.epig import select_epig_online
    from bayesvlm_b200.vlm import CLIP, EncoderResult

    gen = torch.Generator().manual_seed(21)
    D, d_in, n_cls, n_pool, n_targ = 32, 40, 6, 600, 300
    rn = lambda *s: torch.randn(*s, generator=gen)
    spd = lambda d, sc: (lambda w: (w.T @ w) / math.sqrt(4 * d) * sc)(rn(4 * d, d))
    proj = torch.nn.Linear(d_in, D, bias=False)
    pool_act, targ_act = rn(n_pool, d_in), rn(n_targ, d_in)
    with torch.no_grad():
        pool = EncoderResult(proj(pool_act), pool_act)
        targ = EncoderResult(proj(targ_act), targ_act)
    labels = EncoderResult(rn(n_cls, D), rn(n_cls, D))
    info = {"n_img": 1.0, "n_txt": 1.0, "lambda_img": 600.0, "lambda_txt": 220.0}
    idx, scores = select_epig_online(
        label_features=labels, pool_features=pool, target_features=targ, pool_class_ids=torch.randint(0, n_cls, (n_pool,), generator=gen),
        image_projection=proj, clip=CLIP(logit_scale=math.log(20.0)), A_img=spd(d_in, 3e3), A_txt=spd(D, 3e3), B_img=spd(D, 20.0),
        B_txt=spd(D, 20.0), cov_info=info, budget=3, lr=1e-4, hessian_update_scale=10.0, device=torch.device("cuda"),
        num_samples=16, seed=0, pool_max_size=512, target_max_size=256, chunk_size=256)
    assert len(idx) == 3 and len(set(idx)) == 3 and all(0 <= i < n_pool for i in idx)
    assert all(math.isfinite(s) for s in scores)
