"""CPU oracle for the BayesVLM post-hoc Laplace hot path -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithms (MridulPandey17/BayesVLM).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it; the product
package ``bayesvlm_b200`` never does (it has no CPU path at all).

Parity status: PINNED.  The reference ships no tests or golden vectors for this path, so the oracle
is pinned against outputs of the reference implementation itself, generated in the build container by importing
``/root/reference`` (``tests/golden/make_golden.py``, fixtures committed under ``tests/golden/``);
``tests/test_oracle_golden.py`` checks every function below against those fixtures.

Each function cites the reference file:line it follows.  ``dtype=np.float32`` reproduces the reference's arithmetic
type; ``np.float64`` gives the tie-breaker ("who is closer") values.
"""
from __future__ import annotations

import math

import numpy as np

F16 = np.float16
F32 = np.float32


def _rownorm(x):
    return np.sqrt((x * x).sum(axis=-1, keepdims=True))


def _softmax(z, axis=-1):
    z = z - z.max(axis=axis, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=axis, keepdims=True)


# ----------------------------------------------------------------------------------------------------------------------
# K2 -- bayesvlm/hessians.py:10-48
# ----------------------------------------------------------------------------------------------------------------------
def infonce_ggn_naive(X, Y, logit_scale, dtype=np.float64):
    """Literal per-sample evaluation: H = sum_b s^2 J_b (Yh^T (diag p_b - p_b p_b^T) Yh) J_b^T."""
    X = np.asarray(X, dtype)
    Y = np.asarray(Y, dtype)
    s = dtype(math.exp(float(logit_scale)))
    nx = _rownorm(X)
    Xh = X / nx
    Yh = Y / _rownorm(Y)
    P = _softmax((Xh @ Yh.T) * s)                       # :24-27
    D = X.shape[1]
    H = np.zeros((D, D), dtype)
    eye = np.eye(D, dtype=dtype)
    for b in range(X.shape[0]):
        p = P[b]
        J_diag = (Yh * p[:, None]).T @ Yh               # :30
        yp = Yh.T @ p                                    # :33
        J_pp = np.outer(yp, yp)                          # :36
        J_norm = eye / nx[b] - np.outer(X[b], X[b]) / nx[b] ** 3   # :39-43
        H += J_norm @ (J_diag - J_pp) @ J_norm.T * s * s            # :46
    return H


def infonce_ggn_collapsed(X, Y, logit_scale, dtype=np.float64):
    """Same quantity without any [B,D,D] / [B,C,D] temporaries (the form the CUDA pipeline evaluates)."""
    X = np.asarray(X, dtype)
    Y = np.asarray(Y, dtype)
    s = dtype(math.exp(float(logit_scale)))
    nx = _rownorm(X)
    Xh = X / nx
    Yh = Y / _rownorm(Y)
    L = Xh @ Yh.T
    P = _softmax(L * s)
    w = (s * s) / (nx[:, 0] ** 2)
    q = (w[:, None] * P).sum(axis=0)
    m = P @ Yh
    t = (m * Xh).sum(-1, keepdims=True)
    u = (P * L) @ Yh - m * t
    a = (u * Xh).sum(-1)
    H = (Yh * q[:, None]).T @ Yh
    H -= (m * w[:, None]).T @ m
    H -= (Xh * w[:, None]).T @ u
    H -= (u * w[:, None]).T @ Xh
    H += (Xh * (w * a)[:, None]).T @ Xh
    return H


# ----------------------------------------------------------------------------------------------------------------------
# K3 -- bayesvlm/hessians.py:50-117
# ----------------------------------------------------------------------------------------------------------------------
def siglip_ggn_naive(X, indices, Y, logit_scale, logit_bias, dtype=np.float64):
    X = np.asarray(X, dtype)
    Y = np.asarray(Y, dtype)
    assert X.shape[1] == Y.shape[1], "The input and output dimensions must be the same"   # :77
    s = dtype(math.exp(float(logit_scale)))
    nx = _rownorm(X)
    Xh = X / nx
    Yh = Y / _rownorm(Y)
    logits = Xh @ Yh.T * s + dtype(logit_bias)                       # :88
    labels = 2 * np.eye(Y.shape[0], dtype=dtype) - 1                 # :89-90
    labels = labels[np.asarray(indices)]
    sig = 1.0 / (1.0 + np.exp(-(logits * labels)))                   # :93
    lam = s * s * sig * (1 - sig)                                    # :94
    D = X.shape[1]
    eye = np.eye(D, dtype=dtype)
    H = np.zeros((D, D), dtype)
    for b in range(X.shape[0]):
        hess = (Yh * lam[b][:, None]).T @ Yh                         # :103-106
        J = eye / nx[b] - np.outer(X[b], X[b]) / nx[b] ** 3          # :109-111
        H += J @ hess @ J                                            # :113
    return H


def siglip_ggn_collapsed(X, Y, logit_scale, logit_bias, dtype=np.float64):
    X = np.asarray(X, dtype)
    Y = np.asarray(Y, dtype)
    s = dtype(math.exp(float(logit_scale)))
    nx = _rownorm(X)
    Xh = X / nx
    Yh = Y / _rownorm(Y)
    L = Xh @ Yh.T
    sig = 1.0 / (1.0 + np.exp(-(L * s + dtype(logit_bias))))
    lam = s * s * sig * (1 - sig)
    w = 1.0 / (nx[:, 0] ** 2)
    q = (w[:, None] * lam).sum(axis=0)
    u = (lam * L) @ Yh
    a = (u * Xh).sum(-1)
    H = (Yh * q[:, None]).T @ Yh
    H -= (Xh * w[:, None]).T @ u
    H -= (u * w[:, None]).T @ Xh
    H += (Xh * (w * a)[:, None]).T @ Xh
    return H


# ----------------------------------------------------------------------------------------------------------------------
# K0 / K1 -- scripts/hessian_estimation.py:55-108
# ----------------------------------------------------------------------------------------------------------------------
def kfac_ggn(source_embeds, source_activations, target_embeds, num_classes, batch_size, logit_scale, logit_bias=0.0,
             likelihood="info_nce", dtype=np.float64, literal_batches=False):
    """A = sum_cb act^T act / sqrt(n), B = sum_cb sum_db H / sqrt(n) with the reference's dropped remainders.

    ``literal_batches`` walks the data batches of size ``batch_size`` one by one exactly like the reference loop
    (:71-97); otherwise the first floor(num_classes/batch_size)*batch_size sources of a class batch go through the
    collapsed form at once (identical sum: every source row's softmax is independent of the batching).
    """
    if likelihood not in ("info_nce", "siglip"):
        raise ValueError(f"Invalid likelihood: {likelihood}, must be one of ['info_nce', 'siglip'].")
    E = np.asarray(source_embeds, dtype)
    Aact = np.asarray(source_activations, dtype)
    T = np.asarray(target_embeds, dtype)
    ncb = len(T) // num_classes                                      # :55
    if ncb == 0:
        raise ValueError(f"To few datapoints for K-FAC approximation. Need at least {num_classes} datapoints.")
    A = 0
    B = 0
    for i in range(ncb):                                             # :62
        lo, hi = i * num_classes, (i + 1) * num_classes
        tgt, src, act = T[lo:hi], E[lo:hi], Aact[lo:hi]              # :67-69 (paired rows)
        ndb = len(src) // batch_size                                 # :71  (remainder dropped for B only)
        if literal_batches:
            for j in range(ndb):
                xb = src[j * batch_size:(j + 1) * batch_size]
                if likelihood == "info_nce":
                    B = B + infonce_ggn_naive(xb, tgt, logit_scale, dtype)
                else:
                    idx = np.arange(j * batch_size, (j + 1) * batch_size)
                    B = B + siglip_ggn_naive(xb, idx, tgt, logit_scale, logit_bias, dtype)
        elif ndb > 0:
            xb = src[: ndb * batch_size]
            if likelihood == "info_nce":
                B = B + infonce_ggn_collapsed(xb, tgt, logit_scale, dtype)
            else:
                B = B + siglip_ggn_collapsed(xb, tgt, logit_scale, logit_bias, dtype)
        if likelihood == "siglip":                                   # :101-104 ones column
            act = np.concatenate([act, np.ones_like(act[:, :1])], axis=1)
        A = A + act.T @ act                                          # :100
    n = ncb * num_classes
    return A / math.sqrt(n), B / math.sqrt(n)                        # :106-108


# ----------------------------------------------------------------------------------------------------------------------
# C1 -- bayesvlm/hessians.py:137-184
# ----------------------------------------------------------------------------------------------------------------------
def covariance(A, B, n, lmbda, dtype=np.float64):
    A = np.asarray(A, dtype)
    B = np.asarray(B, dtype)
    sn, sl = math.sqrt(n), math.sqrt(lmbda)
    A_inv = np.linalg.inv(A * dtype(sn) + dtype(sl) * np.eye(A.shape[0], dtype=dtype))
    B_inv = np.linalg.inv(B * dtype(sn) + dtype(sl) * np.eye(B.shape[0], dtype=dtype))
    return A_inv, B_inv


def log_marglik(A, B, n, lmbda, weight_norm_sq, n_params):
    """Objective of optimize_prior_precision (hessians.py:253-260): log_prior - (p logdet A_ + q logdet B_), no 1/2."""
    A = np.asarray(A, np.float64)
    B = np.asarray(B, np.float64)
    sn, sl = math.sqrt(n), math.sqrt(lmbda)
    ld_a = np.linalg.slogdet(A * sn + sl * np.eye(A.shape[0]))[1]
    ld_b = np.linalg.slogdet(B * sn + sl * np.eye(B.shape[0]))[1]
    log_prior = -0.5 * lmbda * weight_norm_sq + 0.5 * n_params * math.log(lmbda)
    return log_prior - (ld_a * A.shape[0] + ld_b * B.shape[0])


# ----------------------------------------------------------------------------------------------------------------------
# P1 / P2 / P3 -- bayesvlm/vlm.py:617-684, scripts/zeroshot.py:119-120
# ----------------------------------------------------------------------------------------------------------------------
def map_logits(src_embeds, tgt_embeds, logit_scale, logit_bias=0.0, dtype=np.float32):
    e = np.asarray(src_embeds, dtype)
    t = np.asarray(tgt_embeds, dtype)
    e = e / _rownorm(e)
    t = t / _rownorm(t)
    return e @ t.T * dtype(math.exp(float(logit_scale))) + dtype(logit_bias)      # :627


def predictive(src_embeds, src_acts, tgt_embeds, tgt_acts, A_inv_src, B_inv_src, A_inv_tgt, B_inv_tgt, logit_scale,
               src_bias=False, tgt_bias=False, dtype=np.float32):
    """Logit mean / variance of the Kronecker-Laplace predictive, operation by operation as vlm.py:630-684."""
    e = np.asarray(src_embeds, dtype)
    t = np.asarray(tgt_embeds, dtype)
    a_s = np.asarray(src_acts, dtype)
    a_t = np.asarray(tgt_acts, dtype)
    if src_bias:                                                     # :650-651
        a_s = np.concatenate([a_s, np.ones_like(a_s[:, :1])], axis=-1)
    if tgt_bias:                                                     # :653-654
        a_t = np.concatenate([a_t, np.ones_like(a_t[:, :1])], axis=-1)
    Ais = np.asarray(A_inv_src, dtype)
    Ait = np.asarray(A_inv_tgt, dtype)
    beta = np.asarray(B_inv_src, dtype).diagonal()                   # :659
    delta = np.asarray(B_inv_tgt, dtype).diagonal()                  # :660
    src_cov = ((a_s @ Ais) * a_s).sum(-1)[:, None] * beta            # :662
    tgt_cov = ((a_t @ Ait) * a_t).sum(-1)[:, None] * delta           # :663
    norm_s = e * e + src_cov                                         # :665
    En_s = norm_s.sum(-1, keepdims=True)
    norm_t = t * t + tgt_cov                                         # :667
    En_t = norm_t.sum(-1, keepdims=True)
    mean = (e / np.sqrt(En_s)) @ (t / np.sqrt(En_t)).T               # :671
    term1 = norm_s @ tgt_cov.T                                       # :674
    term2 = src_cov @ (t * t).T                                      # :675
    var = (term1 + term2) / (En_s * En_t.T)                          # :677
    s = dtype(math.exp(float(logit_scale)))
    return mean * s, var * (s * s)                                   # :681-684 (logit_bias NOT added)


def probit_softmax(mean, var, dtype=np.float32):
    mean = np.asarray(mean, dtype)
    var = np.asarray(var, dtype)
    kappa = 1 / np.sqrt(1.0 + dtype(math.pi / 8) * var)             # zeroshot.py:119
    return _softmax(kappa * mean, axis=-1)                           # zeroshot.py:120


def probit_softmax_method_quirk(mean, var, dtype=np.float32):
    """ProbabilisticLogits.softmax(num_samples=0) on 2-D var: uses diag of the N x C matrix (vlm.py:76-78)."""
    mean = np.asarray(mean, dtype)
    var = np.asarray(var, dtype)
    diag = np.diagonal(var, axis1=-2, axis2=-1)
    return _softmax(mean / np.sqrt(1 + dtype(math.pi / 8) * diag), axis=-1)


# ----------------------------------------------------------------------------------------------------------------------
# E0 / E1 / E2 -- bayesvlm/vlm.py:116-123, bayesvlm/epig.py:275-397
# ----------------------------------------------------------------------------------------------------------------------
def sample_probas(mean, var, eps):
    """eps [K, N, Cl] standard normal draws -> probabilities [N, K, Cl] fp32 (vlm.py:121-123)."""
    mean = np.asarray(mean, F32)
    std = np.sqrt(np.asarray(var, F32))
    samples = np.asarray(eps, F32) * std + mean
    return _softmax(np.transpose(samples, (1, 0, 2)), axis=2)


def _xlogy_f16(p16, device="cpu"):
    """torch.xlogy(p, p) on Half tensors.  torch's two backends round differently (both verified against torch 2.11,
    scripts/probe_epig_parity.py and tests/golden/torch_cuda_half_semantics.npz):
      device="cpu":  log(y) is rounded to Half BEFORE the multiply (c10 Half math: `std::log(Half) -> Half`);
      device="cuda": `x * std::log(y)` is evaluated in float in device code and rounded to Half ONCE."""
    p32 = p16.astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        lg = np.log(p32.astype(np.float64)).astype(F32)  # correctly rounded logf
        if device == "cpu":
            lg = lg.astype(F16).astype(F32)
        out = (p32 * lg).astype(F16)
    out[p16 == 0] = 0
    return out


def _div_scalar_f16(x16, k, device="cpu"):
    """Half tensor / python scalar: true fp32 division on the CPU, multiplication by the fp32 reciprocal (computed in
    double) on CUDA (div_true_kernel_cuda's CPU-scalar shortcut); the two agree except for rare last-bit cases."""
    if device == "cpu":
        return (x16.astype(F32) / F32(k)).astype(F16)
    return (x16.astype(F32) * F32(1.0 / float(k))).astype(F16)


def entropy_from_probs_f16(p16, device="cpu"):
    """-sum xlogy with fp32 accumulation rounded to fp16 (epig.py:292 on Half input)."""
    return (-(_xlogy_f16(p16, device).astype(F32).sum(-1).astype(F16))).astype(F16)


def marginal_entropy_f16(probs16, device="cpu"):
    """H[mean_K p] on fp16 probabilities [N, K, Cl] with the reference's rounding points (epig.py:306-308)."""
    assert probs16.ndim == 3
    k = probs16.shape[1]
    s = probs16.astype(F32).sum(axis=1)
    pbar = (s / F32(k)).astype(F16) if device == "cpu" else (s * F32(1.0 / k)).astype(F16)
    return entropy_from_probs_f16(pbar, device)


def epig_from_probs_f16(pool16, targ16, chunk_size=8192, device="cpu"):
    """EPIG scores with every fp16 rounding point of epig.py:342-397 (returns float32 [N_p]).

    `device` selects whose Half kernels are restated: the reference run on the CPU (the golden fixtures) or on CUDA (what
    the B200 kernels must reproduce; pinned by tests/golden/torch_cuda_half_semantics.npz, recorded on a B200)."""
    assert pool16.ndim == targ16.ndim == 3
    n_t, k, cl = targ16.shape
    h_pool = marginal_entropy_f16(pool16, device)                                       # :371
    h_targ = marginal_entropy_f16(targ16, device)
    ht = h_targ.astype(F32).sum()
    h_targ_mean = (ht / F32(n_t)).astype(F16) if device == "cpu" else F16(ht * F32(1.0 / n_t))  # :372
    pool = np.transpose(pool16, (0, 2, 1)).astype(F32)                                  # [N_p, Cl, K]
    targ = np.transpose(targ16, (1, 0, 2)).reshape(k, n_t * cl).astype(F32)             # [K, N_t*Cl]
    acc = np.zeros(pool.shape[0], F32)                                                  # :381
    for lo in range(0, n_t * cl, chunk_size):                                           # :383
        joint = (pool @ targ[:, lo:lo + chunk_size]).astype(F16)                        # :387 fp32 acc -> fp16
        joint = _div_scalar_f16(joint, k, device)                                       # :388
        xl = _xlogy_f16(joint, device)                                                  # :390
        s = xl.astype(F32).sum(axis=(-2, -1)).astype(F16)                               # :391 sum -> fp16
        h = _div_scalar_f16((-s).astype(F16), n_t, device)                              # :391 "/ N_t" in fp16
        acc += h.astype(F32)                                                            # :393 fp32 accumulate
    return (h_pool + h_targ_mean).astype(F32) - acc                                     # :395


def epig_from_probs_f32(pool, targ):
    """Noise-free definition of the same score (fp64 evaluation of fp16/fp32 probabilities)."""
    pool = np.asarray(pool, np.float64)
    targ = np.asarray(targ, np.float64)
    n_t, k, cl = targ.shape

    def H(p):
        with np.errstate(divide="ignore", invalid="ignore"):
            v = np.where(p > 0, p * np.log(p), 0.0)
        return -v.sum(-1)

    h_pool = H(pool.mean(1))
    h_targ = H(targ.mean(1)).mean()
    joint = np.einsum("pkc,tkd->ptcd", pool, targ) / k
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.where(joint > 0, joint * np.log(joint), 0.0)
    h_joint = -v.sum(axis=(2, 3)).mean(axis=1)
    return h_pool + h_targ - h_joint


# ---------------------------------------------------------------------------------------------------------------------
# Support-set search (reference bayesvlm/knn.py) -- SURVEY.md section 8(f) row 3
# ---------------------------------------------------------------------------------------------------------------------
def _knn_diag_cov(acts, A_inv, B_inv, dtype):
    a = np.asarray(acts, dtype)
    return ((a @ np.asarray(A_inv, dtype)) * a).sum(-1)[:, None] * np.asarray(B_inv, dtype).diagonal()   # knn.py:66-69


def knn_expected_cosine(test_embeds, test_acts, train_embeds, train_acts, A_inv, B_inv, dtype=np.float32):
    """[N_test, N_train] expected cosine similarity, operation by operation as knn.py:66-79."""
    mu_te, mu_tr = np.asarray(test_embeds, dtype), np.asarray(train_embeds, dtype)
    cov_te, cov_tr = _knn_diag_cov(test_acts, A_inv, B_inv, dtype), _knn_diag_cov(train_acts, A_inv, B_inv, dtype)
    en_tr = (mu_tr * mu_tr + cov_tr).sum(-1, keepdims=True)          # :71-72
    en_te = (mu_te * mu_te + cov_te).sum(-1, keepdims=True)          # :73-74
    return (mu_te / np.sqrt(en_te)) @ (mu_tr / np.sqrt(en_tr)).T     # :77-80


def knn_diagonal_wasserstein(mu1, mu2, cov1, cov2, dtype=np.float32):
    """knn.py:6-17.  The pairwise squared distance is formed from differences (the quantity ``cdist(...)**2`` stands for;
    torch's own evaluation goes through a matmul for large inputs and carries fp32 cancellation noise)."""
    mu1, mu2 = np.asarray(mu1, dtype), np.asarray(mu2, dtype)
    cov1, cov2 = np.asarray(cov1, dtype), np.asarray(cov2, dtype)
    l2 = ((mu1[:, None, :] - mu2[None, :, :]) ** 2).sum(-1)          # :8
    cross = 2 * (np.sqrt(cov1) @ np.sqrt(cov2).T)                    # :11
    return l2 + cov1.sum(-1)[:, None] + cov2.sum(-1)[None, :] - cross   # :14


def knn_neg_wasserstein(test_embeds, test_acts, train_embeds, train_acts, A_inv, B_inv, dtype=np.float32):
    """[N_test, N_train] similarity of find_similar_samples_wasserstein (knn.py:163-169)."""
    cov_te, cov_tr = _knn_diag_cov(test_acts, A_inv, B_inv, dtype), _knn_diag_cov(train_acts, A_inv, B_inv, dtype)
    return -knn_diagonal_wasserstein(test_embeds, train_embeds, cov_te, cov_tr, dtype)


def knn_support(sim, indices_test, values_test, k_nearest, buffersize=150):
    """Support-set selection from a similarity matrix, loop by loop as knn.py:86-135 (also :171-218).
    Returns {test_idx: dict(score, indices, similarities)} in insertion order."""
    sim = np.asarray(sim)
    n_test, n_train = sim.shape
    width = min(k_nearest + buffersize, n_train)
    order = np.argsort(-sim, axis=1, kind="stable")[:, :width]       # topk, sorted descending (:86)
    vals = np.take_along_axis(sim, order, axis=1)
    goal = k_nearest * n_test
    k_ = k_nearest
    while True:                                                      # :90-106
        flat = order[:, :k_].T.flatten()                             # :92
        if len(np.unique(flat)) >= goal:                             # :99
            while len(np.unique(flat)) > goal:                       # :23-26
                flat = flat[:-1]
            break
        if k_ >= width:
            raise ValueError("not enough distinct neighbours")      # (the reference would loop forever here)
        k_ += 1
    kept = set(np.unique(flat).tolist())                             # :108
    out = {}
    for i in range(n_test):                                          # :114-133
        ids, sims = [], []
        for idx, val in zip(order[i, :k_], vals[i, :k_]):
            if int(idx) in kept:
                ids.append(int(idx))
                sims.append(float(val))
        out[int(indices_test[i])] = dict(score=float(values_test[i]), indices=ids, similarities=sims)
    return out
