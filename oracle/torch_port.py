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
@torch.no_grad()
def predictive(src_embeds, src_acts, tgt_embeds, tgt_acts, A_inv_src, B_inv_src, A_inv_tgt, B_inv_tgt, logit_scale,
               src_bias=False, tgt_bias=False):
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
    s = math.exp(float(logit_scale))
    return mean * s, var * s * s                                               # :679-684 (no logit_bias)


@torch.no_grad()
def probit_softmax(mean, var):
    return torch.softmax(mean / torch.sqrt(1 + math.pi / 8 * var), dim=-1)     # zeroshot.py:119-120


# ---------------------------------------------------------------------------------------------------------------------
# E0 / E1 / E2 -- bayesvlm/vlm.py:116-123, bayesvlm/epig.py:275-397
# ---------------------------------------------------------------------------------------------------------------------
def sample_probas(mean, var, num_samples, seed):
    torch.manual_seed(seed)                                                    # vlm.py:113-114
    noise = torch.randn(num_samples, mean.shape[0], mean.shape[1])             # vlm.py:121
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
