"""Torch-CPU port of the reference ALGORITHMS on the Laplace hot path -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference (MridulPandey17/BayesVLM) is a PyTorch program and cannot travel to the GPU box; this module restates
its hot-path functions with the same ATen operations and the same arithmetic cost (including the [B,C,D] broadcast
and the per-sample [B,D,D] Jacobian sandwich of hessians.py:10-48 that the CUDA pipeline eliminates), so that
``bench.py --impl reference`` and the ``cpu_baseline`` leg time what the reference's own CPU path would do on the
box's host cores.  Only ``tests/`` and ``bench.py`` import it; the product package never does.

Parity status: PINNED -- ``tests/test_oracle_golden.py`` checks every function here against the golden outputs of
the reference itself (``tests/golden/make_golden.py``).  Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import math

import torch


def _unit(x):
    nrm = x.norm(p=2, dim=-1, keepdim=True)
    return x / nrm, nrm


# ---------------------------------------------------------------------------------------------------------------------
# K2 -- bayesvlm/hessians.py:10-48
# ---------------------------------------------------------------------------------------------------------------------
def infonce_ggn(source, target, logit_scale):
    s = math.exp(float(logit_scale))
    xh, nx = _unit(source)
    yh, _ = _unit(target)
    p = torch.softmax(xh @ yh.T * s, dim=-1)                                   # :24-27
    weighted = yh.unsqueeze(0) * p.unsqueeze(-1)                               # [B,C,D] broadcast, :30
    second = weighted.transpose(1, 2) @ yh                                     # [B,D,D]
    mean = (yh.T @ p.unsqueeze(-1))                                            # [B,D,1], :33
    cov = second - mean @ mean.transpose(1, 2)                                 # :36,:46
    d = source.shape[-1]
    jac = torch.eye(d, dtype=source.dtype, device=source.device) / nx.unsqueeze(-1) \
        - source.unsqueeze(2) * source.unsqueeze(1) / (nx ** 3).unsqueeze(-1)  # :39-43
    return (jac @ cov @ jac.transpose(1, 2) * s ** 2).sum(dim=0)               # :46-48


# ---------------------------------------------------------------------------------------------------------------------
# K3 -- bayesvlm/hessians.py:50-117
# ---------------------------------------------------------------------------------------------------------------------
def siglip_ggn(x, indices, y, logit_scale, logit_bias, chunk_size_j=None):
    n_y, d_y = y.shape
    assert x.shape[1] == d_y, "The input and output dimensions must be the same"   # :77
    chunk = n_y if chunk_size_j is None else chunk_size_j
    xh, nx = _unit(x)
    yh, _ = _unit(y)
    s = math.exp(float(logit_scale))
    z = xh @ yh.T * s + float(logit_bias)                                      # :88
    sign = (2 * torch.eye(n_y, dtype=x.dtype, device=x.device) - 1)[indices]                    # :89-90
    sg = torch.sigmoid(z * sign)                                               # :93
    lam = s * s * sg * (1 - sg)                                                # :94
    d = x.shape[1]
    jac = torch.eye(d, dtype=x.dtype, device=x.device).unsqueeze(0) / nx.unsqueeze(-1) \
        - x.unsqueeze(2) * x.unsqueeze(1) / (nx.unsqueeze(-1) ** 3)            # :109-111
    total = 0
    for lo in range(0, n_y, chunk):                                            # :98
        yc = yh[lo:lo + chunk]
        outer = yc.unsqueeze(2) * yc.unsqueeze(1)                              # [chunk,D,D], :103
        hess = torch.einsum("bc,cde->bde", lam[:, lo:lo + chunk], outer)       # :106
        total = total + torch.einsum("bij,bjk,bkl->il", jac, hess, jac)        # :113
    return total


# ---------------------------------------------------------------------------------------------------------------------
# K0 -- scripts/hessian_estimation.py:26-109
# ---------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def kfac_ggn(source_embeds, source_activations, target_embeds, num_classes, batch_size, logit_scale, logit_bias=0.0,
             likelihood="info_nce", siglip_chunk_size_j=8000, max_data_batches=None):
    """The reference double loop.  ``max_data_batches`` truncates the data-batch loop of every class batch (bench.py's
    bounded sample: the cost per data batch is constant for fixed C and D); None = the full reference behaviour."""
    if likelihood not in ("info_nce", "siglip"):
        raise ValueError(f"Invalid likelihood: {likelihood}, must be one of ['info_nce', 'siglip'].")
    n_cb = len(target_embeds) // num_classes                                   # :55
    if n_cb == 0:
        raise ValueError(f"To few datapoints for K-FAC approximation. Need at least {num_classes} datapoints.")
    A = 0
    B = 0
    for i in range(n_cb):                                                      # :62
        lo, hi = i * num_classes, (i + 1) * num_classes
        tgt, src, act = target_embeds[lo:hi], source_embeds[lo:hi], source_activations[lo:hi]
        n_db = len(src) // batch_size                                          # :71
        if max_data_batches is not None:
            n_db = min(n_db, max_data_batches)
        for j in range(n_db):
            xb = src[j * batch_size:(j + 1) * batch_size]
            if likelihood == "info_nce":
                B = B + infonce_ggn(xb, tgt, logit_scale)                      # :80-84
            else:
                idx = torch.arange(j * batch_size, (j + 1) * batch_size)
                B = B + siglip_ggn(xb, idx, tgt, logit_scale, logit_bias, siglip_chunk_size_j)   # :85-94
        if likelihood == "siglip":
            act = torch.cat([act, torch.ones_like(act[:, :1])], dim=1)         # :103
        A = A + act.T @ act                                                    # :100,:104
    n = n_cb * num_classes
    return A / math.sqrt(n), B / math.sqrt(n)                                  # :106-108


# ---------------------------------------------------------------------------------------------------------------------
# P1 / P2 / P3 -- bayesvlm/vlm.py:630-684, scripts/zeroshot.py:119-120
# ---------------------------------------------------------------------------------------------------------------------
def predictive_grad(src_embeds, src_acts, tgt_embeds, tgt_acts, A_inv_src, B_inv_src, A_inv_tgt, B_inv_tgt, logit_scale,
                    src_bias=False, tgt_bias=False):
    """Same expression with autograd left on (the 1 x C step of the online loop, epig.py:214-227); logit_scale may be a tensor."""
    if src_bias:
        src_acts = torch.cat([src_acts, torch.ones_like(src_acts[:, :1])], dim=-1)     # :650-651
    if tgt_bias:
        tgt_acts = torch.cat([tgt_acts, torch.ones_like(tgt_acts[:, :1])], dim=-1)     # :653-654
    beta = B_inv_src.diagonal()                                                # :659
    delta = B_inv_tgt.diagonal()                                               # :660
    cov_s = torch.einsum("ij,jk,ik->i", src_acts, A_inv_src, src_acts).unsqueeze(1) * beta.unsqueeze(0)    # :662
    cov_t = torch.einsum("ij,jk,ik->i", tgt_acts, A_inv_tgt, tgt_acts).unsqueeze(1) * delta.unsqueeze(0)   # :663
    sq_s = src_embeds ** 2 + cov_s                                             # :665
    e_s = sq_s.sum(dim=-1, keepdim=True)
    sq_t = tgt_embeds ** 2 + cov_t                                             # :667
    e_t = sq_t.sum(dim=-1, keepdim=True)
    mean = (src_embeds / e_s.sqrt()) @ (tgt_embeds / e_t.sqrt()).T             # :671
    var = (sq_s @ cov_t.T + cov_s @ (tgt_embeds ** 2).T) / (e_s * e_t.T)       # :674-677
    s = logit_scale.exp() if isinstance(logit_scale, torch.Tensor) else math.exp(float(logit_scale))
    return mean * s, var * s * s                                               # :679-684 (no logit_bias)


@torch.no_grad()
def predictive(*args, **kwargs):
    return predictive_grad(*args, **kwargs)


@torch.no_grad()
def probit_softmax(mean, var):
    return torch.softmax(mean / torch.sqrt(1 + math.pi / 8 * var), dim=-1)     # zeroshot.py:119-120


# ---------------------------------------------------------------------------------------------------------------------
# E0 / E1 / E2 -- bayesvlm/vlm.py:116-123, bayesvlm/epig.py:275-397
# ---------------------------------------------------------------------------------------------------------------------
def sample_probas(mean, var, num_samples, seed):
    torch.manual_seed(seed)                                                    # vlm.py:113-114
    noise = torch.randn((num_samples,) + tuple(mean.shape), device=mean.device)  # vlm.py:121 (device's default generator)
    return torch.softmax((noise * var.sqrt() + mean).permute(1, 0, 2), dim=2)  # vlm.py:122-123


def _entropy(p):
    return -torch.sum(torch.xlogy(p, p), dim=-1)                               # epig.py:292


@torch.no_grad()
def epig_from_probs(probs_pool, probs_targ, chunk_size=8192):
    n_t, k, cl = probs_targ.shape
    h_pool = _entropy(probs_pool.mean(dim=1))                                  # :371
    h_targ = _entropy(probs_targ.mean(dim=1)).mean()                           # :372
    pool = probs_pool.permute(0, 2, 1)                                         # [N_p, Cl, K]
    targ = probs_targ.permute(1, 0, 2).reshape(k, n_t * cl)                    # [K, N_t * Cl]
    h_joint = torch.zeros(pool.shape[0], device=pool.device)                   # :381
    for lo in range(0, n_t * cl, chunk_size):                                  # :383
        joint = pool @ targ[:, lo:lo + chunk_size] / k                         # :387-388
        h_joint += -torch.sum(torch.xlogy(joint, joint), dim=(-2, -1)) / n_t   # :390-393
    return h_pool + h_targ - h_joint                                           # :395


@torch.no_grad()
def epig_from_logits(mean_p, var_p, mean_t, var_t, seed, num_samples, chunk_size=4096):
    out = []
    for lo in range(0, mean_p.shape[0], chunk_size):                           # epig.py:323
        pt = sample_probas(mean_t, var_t, num_samples, seed + lo).to(torch.float16)                          # :324
        pp = sample_probas(mean_p[lo:lo + chunk_size], var_p[lo:lo + chunk_size], num_samples, seed + lo).to(torch.float16)
        out.append(epig_from_probs(pp, pt, chunk_size=chunk_size).to(torch.float32))                         # :336
    return torch.cat(out, dim=0)


# ---------------------------------------------------------------------------------------------------------------------
# C1 + prior precision -- bayesvlm/hessians.py:170-201, 219-280 (the reference forms: inv / logdet every step)
# ---------------------------------------------------------------------------------------------------------------------
def compute_covariance(A, B, n, lmbda):
    sqrt_n, sqrt_l = torch.sqrt(n), torch.sqrt(lmbda)                          # :176-177
    A = A * sqrt_n + sqrt_l * torch.eye(A.size(0), device=A.device, dtype=A.dtype)     # :178
    B = B * sqrt_n + sqrt_l * torch.eye(B.size(0), device=B.device, dtype=B.dtype)     # :179
    return torch.linalg.inv(A), torch.linalg.inv(B)                            # :181-184


def compute_covariances(A_img, B_img, A_txt, B_txt, info):
    t = lambda key, like: torch.tensor(info[key], dtype=like.dtype, device=like.device)                  # :194-197
    return (compute_covariance(A_img, B_img, t("n_img", A_img), t("lambda_img", A_img)),
            compute_covariance(A_txt, B_txt, t("n_txt", A_txt), t("lambda_txt", A_txt)))


def optimize_prior_precision(weight, A, B, lmbda_init, n, lr, num_steps, device):
    norm_sq = (weight.detach() ** 2).sum()                                     # :267-268 (projection has one weight, no bias)
    n_par = weight.numel()                                                     # :270-271
    A, B = A.to(device), B.to(device)
    log_lmbda = torch.nn.Parameter(torch.tensor(lmbda_init, device=device, dtype=torch.float32).log())   # :241-243
    sqrt_n = torch.tensor(n, device=device, dtype=torch.float32).sqrt()        # :244
    opt = torch.optim.Adam([log_lmbda], lr=lr, maximize=True)                  # :246
    for _ in range(num_steps):
        opt.zero_grad()
        lmbda = log_lmbda.exp()
        sqrt_l = lmbda.sqrt()
        A_ = A * sqrt_n + sqrt_l * torch.eye(A.shape[0], device=device, dtype=A.dtype)  # :255
        B_ = B * sqrt_n + sqrt_l * torch.eye(B.shape[0], device=device, dtype=B.dtype)  # :256
        log_prior = -0.5 * lmbda * norm_sq + 0.5 * n_par * torch.log(lmbda)    # :273-274
        log_det = torch.logdet(A_) * A_.shape[0] + torch.logdet(B_) * B_.shape[0]       # :276-280 (quirk: p and q, no 1/2)
        (log_prior - log_det).backward()                                       # :260-262
        opt.step()
    return log_lmbda.exp()


# ---------------------------------------------------------------------------------------------------------------------
# online greedy EPIG loop -- bayesvlm/epig.py:44-273 ('random' pool subsampling; features given as plain tensors)
# ---------------------------------------------------------------------------------------------------------------------
def select_epig_online(label_e, label_a, pool_e, pool_a, targ_e, targ_a, pool_class_ids, weight, logit_scale, A_img, A_txt,
                       B_img, B_txt, cov_info, budget, lr, hessian_update_scale, device, num_samples, seed,
                       pool_max_size=None, target_max_size=None, chunk_size=4096):
    """Returns (selected pool indices, their scores, the score vector of every step, the pool subset those vectors index).  `weight` [D, d_in] is the bias-free
    image projection; residuals are zero (EncoderResult default, vlm.py:36-37)."""
    torch.manual_seed(seed)                                                    # :69
    cov_info = dict(cov_info)
    n_pool_all, n_targ_all = len(pool_e), len(targ_e)
    if pool_max_size is not None:
        pool_max_size = min(pool_max_size, n_pool_all)                         # :71-72
    if target_max_size is not None:
        target_max_size = min(target_max_size, n_targ_all)                     # :73-74
    weight = weight.detach().clone().to(device)                                # :76-77
    dev = lambda t: t.to(device)
    label_e, label_a, pool_e, pool_a, targ_e, targ_a = map(dev, (label_e, label_a, pool_e, pool_a, targ_e, targ_a))
    pool_class_ids = pool_class_ids.to(device)
    A_img, B_img, A_txt, B_txt = map(dev, (A_img, B_img, A_txt, B_txt))
    ls = torch.tensor(float(logit_scale), device=device)
    (Ai, Bi), (At, Bt) = compute_covariances(A_img, B_img, A_txt, B_txt, cov_info)     # :93-94
    idx_t = torch.randperm(n_targ_all)[:target_max_size] if (target_max_size is not None and target_max_size < n_targ_all) \
        else torch.arange(n_targ_all)                                          # :99-102
    idx_p = torch.randperm(n_pool_all)[:pool_max_size] if (pool_max_size is not None and pool_max_size < n_pool_all) \
        else torch.arange(n_pool_all)                                          # :104-108
    selected, scores_sel, all_scores = [], [], []
    for i in range(budget):                                                    # :166
        pe, pa, pc = pool_e[idx_p], pool_a[idx_p], pool_class_ids[idx_p]       # :170-174
        te, ta = targ_e[idx_t], targ_a[idx_t]                                  # :176-179
        mp, vp = predictive(pe, pa, label_e, label_a, Ai, Bi, At, Bt, ls)      # :181
        mt, vt = predictive(te, ta, label_e, label_a, Ai, Bi, At, Bt, ls)      # :182
        epig = epig_from_logits(mp, vp, mt, vt, seed=seed + i, num_samples=num_samples, chunk_size=chunk_size)   # :185-191
        all_scores.append(epig.clone())
        best = None
        for cand in torch.argsort(epig, descending=True):                      # :194-199
            if idx_p[cand].item() in selected:
                continue
            best = cand
            break
        best_a = pa[best].unsqueeze(0)                                         # :201
        best_c = pc[best].unsqueeze(0)                                         # :203
        selected.append(idx_p[best].item())                                    # :205
        scores_sel.append(epig[best].item())                                   # :206
        w = weight.clone().requires_grad_(True)                                # :209-212
        best_embed = best_a @ w.T                                              # :214 (+ zero residual)
        m1, _ = predictive_grad(best_embed, best_a, label_e, label_a, Ai, Bi, At, Bt, ls)                        # :221
        loss = torch.nn.functional.cross_entropy(input=m1, target=best_c)      # :223-226
        loss.backward()                                                        # :227
        with torch.no_grad():
            weight = weight - lr * w.grad                                      # :229-231
            pool_e = pool_a @ weight.T                                         # :234 update_embeddings (:15-42)
            targ_e = targ_a @ weight.T                                         # :235
        picked_e, picked_a = pe[best], pa[best]                                # :237-238 (embeds BEFORE the update)
        A_new = picked_a @ picked_a                                            # :240 `a @ a.T` on a 1-D a: the SCALAR |a|^2 (quirk)
        B_new = infonce_ggn(picked_e.unsqueeze(0), label_e, ls)                # :242-246
        n = 327_680 + i                                                        # :250
        s0, s1 = torch.sqrt(torch.tensor(n)), torch.sqrt(torch.tensor(n + 1))  # :251-252
        A_img = (s0 * A_img + A_new * hessian_update_scale) / s1               # :254
        B_img = (s0 * B_img + B_new * hessian_update_scale) / s1               # :255
        lam = optimize_prior_precision(weight, A_img, B_img, cov_info["lambda_img"], cov_info["n_img"], 1e-3, 20, device)  # :257-267
        cov_info["lambda_img"] = lam.item()                                    # :268
        (Ai, Bi), (At, Bt) = compute_covariances(A_img, B_img, A_txt, B_txt, cov_info)   # :270-271
    return selected, scores_sel, all_scores, idx_p
