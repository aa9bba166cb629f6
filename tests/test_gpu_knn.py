"""Support-set search on the CUDA kernels (bayesvlm_b200.knn) against the oracle and the golden outputs of the reference's
find_similar_samples_* (tests/golden/knn_small.npz).

Parity bar (logit-mean tolerance at unit temperature): similarities within 1e-3 relative of the
fp32 oracle -- the mean GEMM runs in split fp16 (fp32-level), the error left is the fp16 quadratic-form kernel's, which
enters through the normalisers E|e|^2 (observed <= 2e-4); neighbour lists identical to the reference's, modulo candidates
whose similarity ties with the last kept one within that tolerance."""
import math
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import laplace_oracle as O

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def knn():
    return dict(np.load(GOLDEN / "knn_small.npz"))


def _inputs(g):
    from bayesvlm_b200.hessians import KroneckerFactorizedCovariance
    from bayesvlm_b200.vlm import EncoderResult

    t = lambda k: torch.from_numpy(g[k])
    cov = KroneckerFactorizedCovariance(A_inv=t("A_inv"), B_inv=t("B_inv"))
    return EncoderResult(t("train_e"), t("train_a")), EncoderResult(t("test_e"), t("test_a")), cov


def _golden_dict(g, tag):
    out = OrderedDict()
    for r, key in enumerate(g[f"{tag}_keys"]):
        n = int((g[f"{tag}_indices"][r] >= 0).sum())
        out[int(key)] = dict(score=float(g[f"{tag}_scores"][r]), indices=g[f"{tag}_indices"][r, :n].tolist(),
                             similarities=g[f"{tag}_sims"][r, :n].tolist())
    return out


@pytest.mark.parametrize("tag", ["cos", "wass", "cos_k5"])
def test_find_similar_samples_golden(knn, tag):
    from bayesvlm_b200 import knn as K

    train, test, cov = _inputs(knn)
    k_nearest, buf = (int(v) for v in knn[f"{tag}_cfg"])
    fn = K.find_similar_samples_wasserstein if tag.startswith("wass") else K.find_similar_samples_cosine
    res = fn(train, test, torch.from_numpy(knn["indices_test"]), torch.from_numpy(knn["values_test"]), k_nearest=k_nearest,
             source_covariance=cov, device="cuda", buffersize=buf)
    ref = _golden_dict(knn, tag)
    assert isinstance(res, OrderedDict) and list(res.keys()) == list(ref.keys())
    scale = 1.0 if tag.startswith("cos") else float(np.abs(knn["wass_sims"][np.isfinite(knn["wass_sims"])]).max())
    for k in ref:
        assert res[k]["indices"] == ref[k]["indices"], (k, res[k], ref[k])
        assert res[k]["score"] == pytest.approx(ref[k]["score"], rel=1e-6)
        np.testing.assert_allclose(res[k]["similarities"], ref[k]["similarities"], rtol=5e-4, atol=5e-4 * scale)
    ex = K.extract_test_train_indices(res)
    assert sorted(ex["train"]) == knn[f"{tag}_extract_train"].tolist()


@pytest.mark.parametrize("cfg", [dict(n_train=5000, n_test=64, D=512, d_act=768, bias=False),
                                 dict(n_train=3001, n_test=33, D=96, d_act=130, bias=True),
                                 dict(n_train=20000, n_test=257, D=768, d_act=1024, bias=False)])
def test_similarity_matrices_vs_oracle(cfg):
    from bayesvlm_b200 import knn as K
    from bayesvlm_b200.hessians import KroneckerFactorizedCovariance
    from bayesvlm_b200.vlm import EncoderResult

    gen = torch.Generator().manual_seed(cfg["n_train"])
    rn = lambda *s: torch.randn(*s, generator=gen)
    spd = lambda d, sc: (lambda w: (w.T @ w) / math.sqrt(4 * d) * sc)(rn(4 * d, d))
    dA = cfg["d_act"] + (1 if cfg["bias"] else 0)
    A_inv, B_inv = spd(dA, 3e-3), spd(cfg["D"], 0.05)
    centres = rn(40, cfg["D"])
    tr_e = centres[torch.randint(0, 40, (cfg["n_train"],), generator=gen)] + 0.5 * rn(cfg["n_train"], cfg["D"])
    te_e = centres[torch.randint(0, 40, (cfg["n_test"],), generator=gen)] + 0.5 * rn(cfg["n_test"], cfg["D"])
    tr_a, te_a = rn(cfg["n_train"], cfg["d_act"]), rn(cfg["n_test"], cfg["d_act"]) * 1.7
    cov = KroneckerFactorizedCovariance(A_inv=A_inv.cuda(), B_inv=B_inv.cuda())
    train, test = EncoderResult(tr_e.cuda(), tr_a.cuda()), EncoderResult(te_e.cuda(), te_a.cuda())
    ones = lambda a: np.concatenate([a, np.ones_like(a[:, :1])], 1) if cfg["bias"] else a
    args = (te_e.numpy(), ones(te_a.numpy()), tr_e.numpy(), ones(tr_a.numpy()), A_inv.numpy(), B_inv.numpy())

    cos = K.expected_cosine_similarity(test, train, cov, has_bias=cfg["bias"]).cpu().numpy()
    ref = O.knn_expected_cosine(*args, dtype=np.float64)
    assert np.abs(cos - ref).max() <= 5e-4 * np.abs(ref).max()

    rows = slice(0, min(cfg["n_test"], 24))  # the oracle's distance forms [rows, N_train, D] differences
    wass = K.negative_wasserstein_similarity(test, train, cov, has_bias=cfg["bias"]).cpu().numpy()[rows]
    refw = O.knn_neg_wasserstein(args[0][rows], args[1][rows], *args[2:], dtype=np.float64)
    assert np.abs(wass - refw).max() <= 5e-4 * np.abs(refw).max()
    # neighbour ranking: identical top-10 modulo near-ties
    for sim, rf in ((cos[rows], ref[rows]), (wass, refw)):
        tol = 1e-3 * np.abs(rf).max()
        top, top_ref = np.argsort(-sim, 1)[:, :10], np.argsort(-rf, 1)[:, :10]
        for r in range(top.shape[0]):
            kth = rf[r, top_ref[r, -1]]
            for i in set(top[r]) ^ set(top_ref[r]):
                assert abs(rf[r, i] - kth) <= tol


def test_epig_knn_subsampling_uses_kernel_path():
    """select_epig_online(pool_subsampling='knn_*') end to end (reference epig.py:109-161)."""
    from bayesvlm_b200.epig import select_epig_online
    from bayesvlm_b200.vlm import CLIP, EncoderResult

    gen = torch.Generator().manual_seed(33)
    D, d_in, n_cls, n_pool, n_targ = 32, 40, 5, 400, 60
    rn = lambda *s: torch.randn(*s, generator=gen)
    spd = lambda d, sc: (lambda w: (w.T @ w) / math.sqrt(4 * d) * sc)(rn(4 * d, d))
    proj = torch.nn.Linear(d_in, D, bias=False)
    pool_act, targ_act = rn(n_pool, d_in), rn(n_targ, d_in)
    with torch.no_grad():
        pool, targ = EncoderResult(proj(pool_act), pool_act), EncoderResult(proj(targ_act), targ_act)
    labels = EncoderResult(rn(n_cls, D), rn(n_cls, D))
    info = {"n_img": 1.0, "n_txt": 1.0, "lambda_img": 600.0, "lambda_txt": 220.0}
    for mode in ("knn_cosine", "knn_wasserstein"):
        idx, scores = select_epig_online(
            label_features=labels, pool_features=pool, target_features=targ,
            pool_class_ids=torch.randint(0, n_cls, (n_pool,), generator=gen), image_projection=proj,
            clip=CLIP(logit_scale=math.log(20.0)), A_img=spd(d_in, 3e3), A_txt=spd(D, 3e3), B_img=spd(D, 20.0),
            B_txt=spd(D, 20.0), cov_info=info, budget=2, lr=1e-4, hessian_update_scale=10.0, device=torch.device("cuda"),
            num_samples=8, seed=0, chunk_size=256, pool_subsampling=mode, k_nearest_neighbors=2)
        assert len(idx) == 2 and len(set(idx)) == 2 and all(math.isfinite(s) for s in scores)
