"""Support-set search: nearest training samples of a set of test samples under the Laplace posterior (reference
``bayesvlm/knn.py``).

Two similarity measures between Gaussian embeddings ``N(mu, alpha * diag(B_inv))`` (``alpha = a^T A_inv a``):

* expected cosine similarity ``mu_a . mu_b / sqrt(E|e_a|^2 E|e_b|^2)`` (knn.py:61-79) -- exactly the mean of the
  Kronecker-Laplace predictive with unit temperature and the SAME covariance on both sides, so it runs on the fused
  predictive kernels (`CLIP` module, split-fp16 mean GEMM: rankings need fp32-level accuracy);
* negative diagonal 2-Wasserstein distance (knn.py:6-20, 169).  With ``cov = alpha * beta`` the cross term
  ``2 sum_k sqrt(cov_a,k cov_b,k)`` collapses to ``2 sqrt(alpha_a alpha_b) sum(beta)``, hence
  ``W = |mu_a|^2 + |mu_b|^2 - 2 mu_a.mu_b + (sqrt(alpha_a) - sqrt(alpha_b))^2 sum(beta)``: one GEMM (the same
  kernels, un-normalised afterwards) plus the quadratic-form kernel for the alphas.

The selection of the support set out of the per-row top-k lists (knn.py:86-127, 175-218) is restated in vectorised
form; it returns the same ``OrderedDict`` the reference builds.  No CPU fallback: ``device`` must be a CUDA device.
"""
from collections import OrderedDict
from typing import Dict, List

import torch

from . import _lib
from ._lib import lib
from .hessians import KroneckerFactorizedCovariance
from .vlm import CLIP, EncoderResult, _FactorOperand


def diagonal_wasserstein_distance(mu1: torch.Tensor, mu2: torch.Tensor, cov1: torch.Tensor, cov2: torch.Tensor):
    """Squared 2-Wasserstein distance between diagonal Gaussians, all pairs: [N1, N2] (reference knn.py:6-17).
    Generic tensor expression (any device); the search functions below use the kernel path instead."""
    cross = cov1.sqrt() @ cov2.sqrt().t()
    return torch.cdist(mu1, mu2).square() + cov1.sum(-1)[:, None] + cov2.sum(-1)[None, :] - 2.0 * cross


def wdist2(mu1, mu2, cov1, cov2):
    """Alias used by the EPIG pool subsampling (reference knn.py:19-21)."""
    return diagonal_wasserstein_distance(mu1, mu2, cov1, cov2)


def extract_test_train_indices(text_idx_to_train_data: Dict) -> Dict[str, List[int]]:
    """Test indices and the de-duplicated union of their support indices (reference knn.py:28-39)."""
    test = [int(k) for k in text_idx_to_train_data]
    train = {int(i) for entry in text_idx_to_train_data.values() for i in entry["indices"]}
    return dict(test=test, train=list(train))


# ---------------------------------------------------------------------------------------------------------------------
# similarities on the device
# ---------------------------------------------------------------------------------------------------------------------
def _module(cov: KroneckerFactorizedCovariance, device, has_bias: bool) -> CLIP:
    """Unit-temperature similarity module with `cov` on both sides (split-fp16 mean GEMM)."""
    m = CLIP(logit_scale=0.0, device=device, precision="fp16x3")
    if has_bias:
        m.source_projection_has_bias = m.target_projection_has_bias = True
    m.set_covariances(cov, cov)
    return m


def _quadform(act: torch.Tensor, factor: _FactorOperand, has_bias: bool) -> torch.Tensor:
    """alpha_i = a_i^T A_inv a_i through the triangular quadratic-form GEMM (reference knn.py:68-69)."""
    act = _lib.rowmajor(_lib.require_cuda(act, "activations"))
    n, d = act.shape
    out = torch.empty(n, dtype=torch.float32, device=act.device)
    if n == 0:
        return out
    bias = 1 if has_bias else 0
    if d + bias != factor.dA:
        raise ValueError(f"activations have {d}(+{bias}) features but A_inv is {factor.dA}^2")
    ws = _lib.workspace(act.device, lib.bvlm_quadform_workspace_bytes(n, d, bias))
    _lib.run(act.device, "bvlm_quadform", _lib.ptr(act), n, d, act.stride(0), bias, _lib.ptr(factor.w16), factor.dA,
             factor.k_pad, factor.scale, _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(act.device))
    return out


@torch.no_grad()
def expected_cosine_similarity(test: EncoderResult, train: EncoderResult, cov: KroneckerFactorizedCovariance,
                               has_bias: bool = False) -> torch.Tensor:
    """[N_test, N_train] expected cosine similarity under the posterior (reference knn.py:61-79)."""
    return _module(cov, test.embeds.device, has_bias)(test, train).mean


@torch.no_grad()
def negative_wasserstein_similarity(test: EncoderResult, train: EncoderResult, cov: KroneckerFactorizedCovariance,
                                    has_bias: bool = False) -> torch.Tensor:
    """[N_test, N_train] negative diagonal 2-Wasserstein distance (reference knn.py:163-169)."""
    m = _module(cov, test.embeds.device, has_bias)
    src, _, (sum_beta, _, _) = m._sides()
    sim = m(test, train).mean  # mu_a.mu_b / sqrt(E_a E_b)
    a_te, a_tr = _quadform(test.activations, src.factor, has_bias), _quadform(train.activations, src.factor, has_bias)
    n_te, n_tr = test.embeds.square().sum(-1), train.embeds.square().sum(-1)
    root_te, root_tr = (n_te + a_te * sum_beta).sqrt(), (n_tr + a_tr * sum_beta).sqrt()
    sim.mul_(root_te[:, None]).mul_(root_tr[None, :]).mul_(2.0)  # 2 mu_a.mu_b, in place
    sim.sub_(n_te[:, None]).sub_(n_tr[None, :])
    spread = a_te.sqrt()[:, None] - a_tr.sqrt()[None, :]
    return sim.addcmul_(spread, spread, value=-sum_beta)


# ---------------------------------------------------------------------------------------------------------------------
# support-set selection
# ---------------------------------------------------------------------------------------------------------------------
def _support_from_topk(top_idx: torch.Tensor, top_val: torch.Tensor, indices_test, values_test, k_nearest: int):
    """Reference knn.py:86-127: grow the per-row neighbour count k until the neighbour lists hold at least
    ``k_nearest * N_test`` distinct training samples; keep the distinct samples met first in rank-major order (all first
    neighbours, then all second neighbours, ...) up to that count; report, per test sample, those of its first k
    neighbours that were kept.

    Deviation: when even the whole top-k buffer cannot supply enough distinct samples the reference loops forever
    (``k_`` grows past the buffer width, knn.py:102-104); this raises ``ValueError`` instead."""
    n_test, width = top_idx.shape
    goal = k_nearest * n_test
    k = None
    for cand in range(k_nearest, width + 1):
        if torch.unique(top_idx[:, :cand]).numel() >= goal:
            k = cand
            break
    if k is None:
        raise ValueError(f"the top-{width} neighbour lists hold fewer than {goal} distinct training samples; "
                         "increase buffersize or lower k_nearest")
    flat = top_idx[:, :k].t().reshape(-1)
    # longest prefix of `flat` with at most `goal` distinct values == drop trailing elements while there are more
    order = torch.argsort(flat, stable=True)
    sorted_vals = flat[order]
    first_of_run = torch.ones_like(sorted_vals, dtype=torch.bool)
    first_of_run[1:] = sorted_vals[1:] != sorted_vals[:-1]
    is_first = torch.zeros_like(first_of_run)
    is_first[order[first_of_run]] = True  # stable sort: the first element of a run is the first occurrence
    kept_prefix = int((torch.cumsum(is_first.long(), 0) <= goal).sum())
    kept = torch.unique(flat[:kept_prefix])
    member = torch.isin(top_idx[:, :k], kept)

    idx_host, val_host, member_host = top_idx[:, :k].cpu(), top_val[:, :k].cpu(), member.cpu()
    test_ids = torch.as_tensor(indices_test).cpu().tolist()
    test_scores = torch.as_tensor(values_test).cpu().tolist()
    out = OrderedDict()
    for i in range(n_test):
        sel = member_host[i]
        out[test_ids[i]] = dict(score=test_scores[i], indices=idx_host[i][sel].tolist(),
                                similarities=val_host[i][sel].tolist())
    return out


def _find(kind: str, train: EncoderResult, test: EncoderResult, indices_test, values_test, k_nearest: int,
          source_covariance: KroneckerFactorizedCovariance, device, buffersize: int):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("the support-set search runs on CUDA kernels; there is no CPU fallback")
    idx = torch.as_tensor(indices_test)
    sub = EncoderResult(embeds=test.embeds[idx.to(test.embeds.device)].to(dev),
                        activations=test.activations[idx.to(test.activations.device)].to(dev))
    tr = EncoderResult(embeds=train.embeds.to(dev), activations=train.activations.to(dev))
    cov = KroneckerFactorizedCovariance(A_inv=source_covariance.A_inv.to(dev), B_inv=source_covariance.B_inv.to(dev))
    sim = (expected_cosine_similarity if kind == "cosine" else negative_wasserstein_similarity)(sub, tr, cov)
    top = sim.topk(min(k_nearest + buffersize, len(tr)), dim=1)
    return _support_from_topk(top.indices, top.values, indices_test, values_test, k_nearest)


def find_similar_samples_cosine(train: EncoderResult, test: EncoderResult, indices_test: torch.Tensor,
                                values_test: torch.Tensor, k_nearest: int, source_covariance, device: str,
                                buffersize: int = 150):
    """k nearest training samples of ``test[indices_test]`` by expected cosine similarity (reference knn.py:41-137).
    Returns ``OrderedDict[test_idx] = dict(score, indices, similarities)``."""
    return _find("cosine", train, test, indices_test, values_test, k_nearest, source_covariance, device, buffersize)


def find_similar_samples_wasserstein(train: EncoderResult, test: EncoderResult, indices_test: torch.Tensor,
                                     values_test: torch.Tensor, k_nearest: int, source_covariance, device: str,
                                     buffersize: int = 150):
    """k nearest training samples by diagonal 2-Wasserstein distance (reference knn.py:139-220)."""
    return _find("wasserstein", train, test, indices_test, values_test, k_nearest, source_covariance, device, buffersize)
