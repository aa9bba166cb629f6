"""EPIG acquisition on B200 kernels; mirrors the reference's ``bayesvlm/epig.py``.

The scoring functions (reference :275-397) keep their names and rounding behaviour: for fp16 CUDA probabilities the
marginal entropies run as a warp-shuffle reduction kernel and the joint-entropy term as a tcgen05 GEMM whose epilogue
does ``/K -> xlogy -> sum -> /N_t`` with the reference's fp16 rounding points, so the [N_p, Cl, chunk] joint tile is
never materialised (``csrc/epig.cu``).  The online greedy loop (reference :44-273) keeps its torch control flow.
"""
from __future__ import annotations

import copy
from typing import Literal, Optional

import torch

from . import _lib
from ._lib import lib
from .hessians import (FactorSpectrum, compute_covariances, compute_hessian_analytic_InfoNCE, covariance_from_spectra,
                       optimize_prior_precision)
from .vlm import CLIP, EncoderResult, ProbabilisticLogits

_JOINT_TILE_N = 256  # column tile of the joint-entropy kernel; chunk_size must be a multiple of it


def _kernel_path_ok(probs: torch.Tensor) -> bool:
    return probs.is_cuda and probs.dtype == torch.float16


def _prepare_fits(k: int, cl: int, from_noise: bool) -> bool:
    """The fused sample / permute / marginal-entropy kernel stages a block's tiles in shared memory (csrc/epig.cu)."""
    return bool(lib.bvlm_epig_prepare_supported(k, cl, 1 if from_noise else 0))


def _joint_fused_ok(k: int, cl: int, chunk_size: int) -> bool:
    return chunk_size % _JOINT_TILE_N == 0 and chunk_size > 0


def prepare_from_probs(probs: torch.Tensor, want_operand: bool = True, want_entropy: bool = True):
    """fp16 CUDA probabilities [N, K, Cl] -> (operand [N, Cl, Kp] fp16 for the joint-entropy GEMM, marginal entropies [N]
    fp16) in one pass (the permute of reference epig.py:374-376 + marginal_entropy_from_probs :294-311)."""
    n, k, cl = probs.shape
    probs = probs.contiguous()
    dev = probs.device
    oper = torch.empty((n, cl, int(lib.bvlm_epig_operand_k(k))), dtype=torch.float16, device=dev) if want_operand else None
    marg = torch.empty(n, dtype=torch.float16, device=dev) if want_entropy else None
    _lib.run(dev, "bvlm_epig_prepare_from_probs", _lib.ptr(probs), n, k, cl, _lib.ptr(oper), _lib.ptr(marg),
             _lib.stream_ptr(dev))
    return oper, marg


def prepare_from_noise(mean: torch.Tensor, var: torch.Tensor, eps: torch.Tensor, want_probs: bool = False,
                       want_operand: bool = True, want_entropy: bool = True):
    """E0 + E1 + operand layout fused: noise eps [K, N, Cl] (torch.randn order, reference vlm.py:121) and logits
    mean / var [N, Cl] -> (probs [N, K, Cl] fp16 | None, operand [N, Cl, Kp] fp16 | None, marginal entropies [N] fp16 | None)."""
    from .vlm import _check_noise_shapes

    k, n, cl = _check_noise_shapes(mean, var, eps)
    dev = mean.device
    f16 = dict(dtype=torch.float16, device=dev)
    probs = torch.empty((n, k, cl), **f16) if want_probs else None
    oper = torch.empty((n, cl, int(lib.bvlm_epig_operand_k(k))), **f16) if want_operand else None
    marg = torch.empty(n, **f16) if want_entropy else None
    _lib.run(dev, "bvlm_epig_prepare_from_noise", _lib.ptr(mean.contiguous()), _lib.ptr(var.contiguous()),
             _lib.ptr(eps.contiguous()), n, k, cl, _lib.ptr(probs), _lib.ptr(oper), _lib.ptr(marg), _lib.stream_ptr(dev))
    return probs, oper, marg


def prepare_pair_from_noise(mean_a, var_a, eps_a, mean_b, var_b, eps_b):
    """``prepare_from_noise`` for two sample sets sharing [K, Cl] (the target set and one pool chunk of reference
    epig.py:326-333) in ONE kernel launch; returns (operand_a, entropy_a, operand_b, entropy_b)."""
    from .vlm import _check_noise_shapes

    k, na, cl = _check_noise_shapes(mean_a, var_a, eps_a)
    kb, nb, clb = _check_noise_shapes(mean_b, var_b, eps_b)
    if (k, cl) != (kb, clb) or mean_a.device != mean_b.device:
        raise ValueError("both sample sets must share [K, Cl] and the device")
    dev = mean_a.device
    f16 = dict(dtype=torch.float16, device=dev)
    kp = int(lib.bvlm_epig_operand_k(k))
    oper_a, oper_b = torch.empty((na, cl, kp), **f16), torch.empty((nb, cl, kp), **f16)
    marg_a, marg_b = torch.empty(na, **f16), torch.empty(nb, **f16)
    _lib.run(dev, "bvlm_epig_prepare_pair_from_noise", _lib.ptr(mean_a.contiguous()), _lib.ptr(var_a.contiguous()),
             _lib.ptr(eps_a.contiguous()), na, _lib.ptr(oper_a), _lib.ptr(marg_a), _lib.ptr(mean_b.contiguous()),
             _lib.ptr(var_b.contiguous()), _lib.ptr(eps_b.contiguous()), nb, _lib.ptr(oper_b), _lib.ptr(marg_b), k, cl,
             _lib.stream_ptr(dev))
    return oper_a, marg_a, oper_b, marg_b


def entropy_from_probs(probs: torch.Tensor) -> torch.Tensor:
    """H[p] = -sum_y p log p with 0 log 0 = 0 (reference epig.py:275-292)."""
    return -torch.sum(torch.xlogy(probs, probs), dim=-1)


def marginal_entropy_from_probs(probs: torch.Tensor) -> torch.Tensor:
    """H[E_theta p(y|x,theta)] for probs [N, K, Cl] -> [N] (reference epig.py:294-311)."""
    assert probs.ndim == 3
    if not probs.is_cuda:
        raise RuntimeError("bayesvlm_b200.epig runs on CUDA tensors only (no CPU fallback)")
    if _kernel_path_ok(probs) and _prepare_fits(probs.shape[1], probs.shape[2], False):
        return prepare_from_probs(probs, want_operand=False)[1]
    return entropy_from_probs(torch.mean(probs, dim=1))


def joint_entropy_from_operands(oper_pool: torch.Tensor, oper_targ: torch.Tensor, k: int, chunk_size: int) -> torch.Tensor:
    """E_t H[p(y, y_t | x, x_t)] from the permuted operands [N, Cl, Kp]: fp32 accumulation over column chunks of the
    flattened (t, c) axis with the reference's fp16 rounding points inside a chunk (epig.py:376-393)."""
    n_p, cl, _ = oper_pool.shape
    n_t = oper_targ.shape[0]
    dev = oper_pool.device
    out = torch.empty(n_p, dtype=torch.float32, device=dev)
    ws = _lib.workspace(dev, lib.bvlm_epig_joint_operands_workspace_bytes(n_p, n_t, cl, int(chunk_size)), tag="epig_joint")
    _lib.run(dev, "bvlm_epig_joint_entropy_operands", _lib.ptr(oper_pool), n_p, _lib.ptr(oper_targ), n_t, k, cl,
             int(chunk_size), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    return out


def joint_entropy_from_probs(probs_pool: torch.Tensor, probs_targ: torch.Tensor, chunk_size: int) -> torch.Tensor:
    """Same term from fp16 probabilities [N, K, Cl] (permutes internally)."""
    n_p, k, cl = probs_pool.shape
    if probs_targ.shape[1] != k or probs_targ.shape[2] != cl:
        raise ValueError("pool and target probabilities must share [K, Cl]")
    oper_p, _ = prepare_from_probs(probs_pool, want_entropy=False)
    oper_t, _ = prepare_from_probs(probs_targ, want_entropy=False)
    return joint_entropy_from_operands(oper_p, oper_t, k, chunk_size)


def _joint_entropy_torch(probs_pool, probs_targ, chunk_size):
    """Generic-dtype device expression of the same term (used for fp32 'noise-free' scores and odd shapes)."""
    n_t, k, cl = probs_targ.shape
    pool = probs_pool.permute(0, 2, 1)
    targ = probs_targ.permute(1, 0, 2).reshape(k, n_t * cl)
    acc = torch.zeros(pool.shape[0], device=pool.device)
    for lo in range(0, n_t * cl, chunk_size):
        joint = pool @ targ[:, lo:lo + chunk_size] / k
        acc += -torch.sum(torch.xlogy(joint, joint), dim=(-2, -1)) / n_t
    return acc


@torch.no_grad()
def epig_from_probs_using_matmul(probs_pool: torch.Tensor, probs_targ: torch.Tensor, chunk_size: int = 8192):
    """EPIG(x) = H[p(y|x)] + E_t H[p(y_t|x_t)] - E_t H[p(y,y_t|x,x_t)]  (reference epig.py:342-397).

    probs_pool [N_p, K, Cl], probs_targ [N_t, K, Cl] -> [N_p].
    """
    assert probs_pool.ndim == probs_targ.ndim == 3
    if not (probs_pool.is_cuda and probs_targ.is_cuda):
        raise RuntimeError("bayesvlm_b200.epig runs on CUDA tensors only (no CPU fallback)")
    k, cl = probs_targ.shape[1], probs_targ.shape[2]
    if probs_pool.shape[1] != k or probs_pool.shape[2] != cl:
        raise ValueError("pool and target probabilities must share [K, Cl]")
    fused = (_kernel_path_ok(probs_pool) and _kernel_path_ok(probs_targ) and _joint_fused_ok(k, cl, chunk_size) and
             _prepare_fits(k, cl, False))
    if fused:
        oper_p, entropy_pool = prepare_from_probs(probs_pool)
        oper_t, entropy_targ = prepare_from_probs(probs_targ)
        entropy_joint = joint_entropy_from_operands(oper_p, oper_t, k, chunk_size)
        return entropy_pool + torch.mean(entropy_targ) - entropy_joint
    entropy_pool = marginal_entropy_from_probs(probs_pool)
    entropy_targ_mean = torch.mean(marginal_entropy_from_probs(probs_targ))
    return entropy_pool + entropy_targ_mean - _joint_entropy_torch(probs_pool, probs_targ, chunk_size)


@torch.no_grad()
def epig_from_logits_using_matmul(logits_pool: ProbabilisticLogits, logits_targ: ProbabilisticLogits, seed: int,
                                  num_samples: int, chunk_size: int = 4096) -> torch.Tensor:
    """Pool rows in chunks of ``chunk_size``; each chunk re-draws target AND pool samples under ``seed + row_offset``
    (both calls re-seed torch's generator with the same value), casts to fp16 and scores (reference epig.py:313-340).

    The noise comes from ``torch.randn`` on the logits' device exactly as in the reference (``vlm.py:121``); sampling,
    fp16 cast, the permutes and the marginal entropies are one kernel per side, the joint term one tcgen05 GEMM."""
    pieces = []
    n_pool = logits_pool.mean.shape[0]
    cl = logits_pool.mean.shape[-1]
    fused = (logits_pool.var.ndim == 2 and logits_targ.var.ndim == 2 and logits_pool.mean.is_cuda and
             logits_pool.mean.dtype == torch.float32 and logits_targ.mean.dtype == torch.float32 and
             _joint_fused_ok(num_samples, cl, chunk_size) and _prepare_fits(num_samples, cl, True))
    for lo in range(0, n_pool, chunk_size):
        if not fused:
            probs_targ = logits_targ.sample_probas_f16(num_samples, seed=seed + lo)
            chunk = ProbabilisticLogits(mean=logits_pool.mean[lo:lo + chunk_size], var=logits_pool.var[lo:lo + chunk_size])
            probs_pool = chunk.sample_probas_f16(num_samples, seed=seed + lo)
            pieces.append(epig_from_probs_using_matmul(probs_pool, probs_targ, chunk_size=chunk_size).to(torch.float32))
            continue
        dev = logits_targ.mean.device
        torch.manual_seed(seed + lo)
        eps_t = torch.randn((num_samples,) + tuple(logits_targ.mean.shape), device=dev)
        mean_p, var_p = logits_pool.mean[lo:lo + chunk_size], logits_pool.var[lo:lo + chunk_size]
        torch.manual_seed(seed + lo)
        eps_p = torch.randn((num_samples,) + tuple(mean_p.shape), device=mean_p.device)
        oper_t, ent_t, oper_p, ent_p = prepare_pair_from_noise(logits_targ.mean, logits_targ.var, eps_t, mean_p, var_p, eps_p)
        del eps_t, eps_p
        joint = joint_entropy_from_operands(oper_p, oper_t, num_samples, chunk_size)
        pieces.append((ent_p + torch.mean(ent_t) - joint).to(torch.float32))
    return torch.cat(pieces, dim=0)


def update_embeddings(projection: torch.nn.Module, outputs: EncoderResult, device: str = "cuda",
                      keep_on_device: bool = False) -> EncoderResult:
    """Re-project all activations with the current projection weights (reference epig.py:15-42), as one device GEMM
    instead of a CPU DataLoader round trip.  Returns CPU tensors like the reference unless ``keep_on_device``."""
    act = outputs.activations.to(device)
    res = outputs.residuals.to(device)
    with torch.no_grad():
        embeds = projection(act) + res
    new = EncoderResult(embeds=embeds, activations=act, residuals=res)
    return new if keep_on_device else new.to("cpu")


def select_epig_online(label_features: EncoderResult, pool_features: EncoderResult, target_features: EncoderResult,
                       pool_class_ids: torch.Tensor, image_projection: torch.nn.Linear, clip: CLIP, A_img: torch.Tensor,
                       A_txt: torch.Tensor, B_img: torch.Tensor, B_txt: torch.Tensor, cov_info: dict, budget: int,
                       lr: float, hessian_update_scale: float, device: torch.device, num_samples: int, seed: int,
                       pool_max_size: Optional[int] = None, target_max_size: Optional[int] = None,
                       chunk_size: int = 4096, pool_subsampling: Literal["random", "knn"] = "random",
                       k_nearest_neighbors: int = 1, proj_has_bias=False):
    """Greedy online EPIG selection with a rank-one posterior update per pick (reference epig.py:44-273)."""
    torch.manual_seed(seed)
    n_pool_all, n_targ_all = len(pool_features.embeds), len(target_features.embeds)
    if pool_max_size is not None:
        pool_max_size = min(pool_max_size, n_pool_all)
    if target_max_size is not None:
        target_max_size = min(target_max_size, n_targ_all)

    image_projection = copy.deepcopy(image_projection).to(device).train()
    pool_features = pool_features.to(device)
    target_features = target_features.to(device)
    label_features = label_features.to(device)
    pool_class_ids = pool_class_ids.to(device)
    A_img, B_img, A_txt, B_txt = (t.to(device) for t in (A_img, B_img, A_txt, B_txt))

    clip = clip.to(device).eval()
    for p in clip.parameters():
        p.requires_grad = False
    cov_img, cov_txt = compute_covariances(A_img, B_img, A_txt, B_txt, cov_info)
    clip.set_covariances(cov_img, cov_txt)

    if target_max_size is not None and target_max_size < n_targ_all:
        idx_targ = torch.randperm(n_targ_all)[:target_max_size]
    else:
        idx_targ = torch.arange(n_targ_all)

    if pool_subsampling == "random":
        if pool_max_size is not None and pool_max_size < n_pool_all:
            idx_pool = torch.randperm(n_pool_all)[:pool_max_size]
        else:
            idx_pool = torch.arange(n_pool_all)
    elif pool_subsampling in ("knn_cosine", "knn_wasserstein"):
        from .knn import expected_cosine_similarity, negative_wasserstein_similarity

        test_sub = EncoderResult(embeds=target_features.embeds[idx_targ.to(device)],
                                 activations=target_features.activations[idx_targ.to(device)])
        similarity = expected_cosine_similarity if pool_subsampling == "knn_cosine" else negative_wasserstein_similarity
        sim = similarity(test_sub, pool_features, cov_img, has_bias=proj_has_bias)  # [targets, pool] on the kernels
        nearest = torch.argsort(sim, descending=True, dim=1)
        idx_pool = nearest[:, :k_nearest_neighbors].flatten().unique().cpu()
        if len(idx_pool) < budget:
            raise ValueError(f"Could not find enough samples in the pool. Found {len(idx_pool)}, expected at least {budget}.")
    else:
        raise ValueError(f"Unknown subsampling method: {pool_subsampling}")

    selected_indices, epig_scores = [], []
    for step in range(budget):
        pool_sub = EncoderResult(embeds=pool_features.embeds[idx_pool], activations=pool_features.activations[idx_pool])
        pool_ids_sub = pool_class_ids[idx_pool]
        targ_sub = EncoderResult(embeds=target_features.embeds[idx_targ], activations=target_features.activations[idx_targ])

        logits_pool = clip(pool_sub.to(device), label_features.to(device)).detach()
        logits_targ = clip(targ_sub.to(device), label_features.to(device)).detach()
        scores = epig_from_logits_using_matmul(logits_pool, logits_targ, num_samples=num_samples, chunk_size=chunk_size,
                                               seed=seed + step)

        best = None
        for cand in torch.argsort(scores, descending=True):
            if idx_pool[cand].item() in selected_indices:
                print(f"Skipping {cand} as it has already been selected.")
                continue
            best = cand
            break

        best_act = pool_sub.activations[best].unsqueeze(0)
        best_res = pool_sub.residuals[best].unsqueeze(0)
        best_cls = pool_ids_sub[best].unsqueeze(0)
        selected_indices.append(idx_pool[best].item())
        epig_scores.append(scores[best].item())

        # one SGD step on the projection weight with the CE of the mean logits of the picked sample
        for p in image_projection.parameters():
            p.requires_grad = True
        image_projection.zero_grad()
        best_embed = image_projection(best_act) + best_res
        best_logits = clip(EncoderResult(embeds=best_embed, activations=best_act, residuals=best_res), label_features)
        loss = torch.nn.functional.cross_entropy(input=best_logits.mean, target=best_cls)
        loss.backward()
        with torch.no_grad():
            image_projection.weight.data -= lr * image_projection.weight.grad
            image_projection.weight.grad.zero_()

        pool_features = update_embeddings(image_projection, pool_features, device, keep_on_device=True)
        target_features = update_embeddings(image_projection, target_features, device, keep_on_device=True)

        picked_embed = pool_sub.embeds[best]
        picked_act = pool_sub.activations[best]
        # reference quirk (epig.py:240): a 1-D activation makes `a @ a.T` the SCALAR |a|^2, broadcast over A_img
        A_new = torch.dot(picked_act, picked_act)
        B_new = compute_hessian_analytic_InfoNCE(source_embeds=picked_embed.unsqueeze(0).to(device),
                                                 target_embeds=label_features.embeds.to(device),
                                                 logit_scale=clip.logit_scale.data.to(device))
        n_prev = 327_680 + step  # reference hard-codes the size of the initial estimate (epig.py:250)
        s0, s1 = torch.sqrt(torch.tensor(n_prev)), torch.sqrt(torch.tensor(n_prev + 1))
        A_img = (s0 * A_img + A_new * hessian_update_scale) / s1
        B_img = (s0 * B_img + B_new * hessian_update_scale) / s1

        # ONE eigendecomposition per updated image factor feeds the 20 Adam steps on lambda AND the covariance of the
        # optimised lambda; the text-side covariance never changes inside the loop (reference epig.py:255-263 re-inverts all
        # four factors and re-factorises A_img / B_img twenty times per pick).
        spectra = (FactorSpectrum.of(A_img, device), FactorSpectrum.of(B_img, device))
        lmbda_img = optimize_prior_precision(projection=image_projection, A=A_img, B=B_img,
                                             lmbda_init=cov_info["lambda_img"], n=cov_info["n_img"], lr=1e-3,
                                             num_steps=20, device=device, retain_graph=True, spectra=spectra)
        cov_info["lambda_img"] = lmbda_img.item()
        cov_img = covariance_from_spectra(spectra[0], spectra[1], cov_info["n_img"], cov_info["lambda_img"], dtype=A_img.dtype)
        clip.set_covariances(cov_img, cov_txt)

    return selected_indices, epig_scores
