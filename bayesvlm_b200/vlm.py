"""Probabilistic-logit interface of BayesVLM on B200 kernels.

Mirrors the hot-path part of the reference's ``bayesvlm/vlm.py`` (EncoderResult :27-61, ProbabilisticLogits :63-204,
CLIP :567-710, SIGLIP :712-728) with the same names, arguments and quirks; the Kronecker-Laplace predictive
(`CLIP._compute_probabilistic_logits_smith`, reference :630-684) runs as one fused tcgen05 GEMM whose epilogue emits the
logit mean and the rank-2 variance, fed by a triangular quadratic-form GEMM (reference :662-663) -- see
``csrc/predictive.cu``.  The HF encoder wrappers of the reference file are out of scope (they only produce the inputs).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Union

import torch

from . import _lib
from ._lib import lib
from .hessians import KroneckerFactorizedCovariance

# mean-logit GEMM: "fp16x3" = hi/lo split, three fp16 tensor-core products (~2^-22); "fp16+fp8" = fp16 hi.hi plus the two
# error-compensation products in FP8 at twice the rate (~2^-15, 2/3 of the cost); "fp16" = single pass (~2^-11)
_PRECISIONS = {"fp16x3": _lib.PREC_X3, "fp16+fp8": _lib.PREC_X2F8, "fp16": _lib.PREC_X1, "x3": _lib.PREC_X3,
               "x2f8": _lib.PREC_X2F8, "x1": _lib.PREC_X1}
_MIN_D_REDUCED_PRECISION = 128  # embedding widths below this always take the hi/lo split mean GEMM


class EncoderResult:
    """Container of (embeds [N,D], activations [N,d_in], residuals [N,D]); reference vlm.py:27-61."""

    def __init__(self, embeds, activations, residuals=None):
        self.embeds = embeds
        self.activations = activations
        self.residuals = torch.zeros_like(embeds) if residuals is None else residuals

    def clone(self):
        return EncoderResult(self.embeds.clone(), self.activations.clone(), self.residuals.clone())

    def to(self, device):
        self.embeds = self.embeds.to(device)
        self.activations = self.activations.to(device)
        self.residuals = self.residuals.to(device)
        return self

    def __len__(self):
        return len(self.embeds)

    def __getitem__(self, idx):
        if isinstance(idx, (list, torch.Tensor)):
            return EncoderResult(self.embeds[idx], self.activations[idx], self.residuals[idx])
        return self.embeds[idx], self.activations[idx], self.residuals[idx]

    def __repr__(self):
        return (f"EncoderResult(embeds={tuple(self.embeds.shape)}, activations={tuple(self.activations.shape)}, "
                f"device={self.embeds.device})")


def probit_softmax(mean: torch.Tensor, var: torch.Tensor) -> torch.Tensor:
    """Canonical element-wise probit softmax, ``softmax(mean / sqrt(1 + pi/8 var))`` (scripts/zeroshot.py:119-120)."""
    _lib.require_cuda(mean, "mean")
    _lib.require_cuda(var, "var")
    if mean.shape != var.shape or mean.dim() != 2:
        raise ValueError("probit_softmax expects mean and var of identical [N, C] shape")
    mean = mean.contiguous()
    var = var.contiguous()
    out = torch.empty_like(mean)
    n, c = mean.shape
    _lib.run(mean.device, "bvlm_probit_softmax", _lib.ptr(mean), _lib.ptr(var), n, c, c, _lib.ptr(out),
             _lib.stream_ptr(mean.device))
    return out


def _check_noise_shapes(mean: torch.Tensor, var: torch.Tensor, eps: torch.Tensor):
    """Validate the operands of the sampling kernel (it indexes raw pointers): fp32 CUDA tensors on one device with
    mean / var [N, Cl] and eps [K, N, Cl]; returns (K, N, Cl)."""
    _lib.require_cuda(mean, "mean")
    _lib.require_cuda(var, "var")
    _lib.require_cuda(eps, "eps")
    if eps.dim() != 3 or mean.dim() != 2 or tuple(mean.shape) != tuple(var.shape) or tuple(eps.shape[1:]) != tuple(mean.shape):
        raise ValueError(f"expected mean/var [N, Cl] and eps [K, N, Cl]; got {tuple(mean.shape)}, {tuple(var.shape)}, "
                         f"{tuple(eps.shape)}")
    if not (mean.device == var.device == eps.device):
        raise ValueError("mean, var and eps must live on the same device")
    k, n, cl = eps.shape
    return k, n, cl


def sample_probas_from_noise(mean: torch.Tensor, var: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """E0: fp16 ``softmax(mean + eps * sqrt(var))`` in the [N, K, Cl] layout from noise eps [K, N, Cl]."""
    k, n, cl = _check_noise_shapes(mean, var, eps)
    out = torch.empty((n, k, cl), dtype=torch.float16, device=mean.device)
    _lib.run(mean.device, "bvlm_epig_sample_probs", _lib.ptr(mean.contiguous()), _lib.ptr(var.contiguous()),
             _lib.ptr(eps.contiguous()), n, k, cl, _lib.ptr(out), _lib.stream_ptr(mean.device))
    return out


_MC_NOISE_BYTES = 1 << 30  # noise of one kernel launch: G draws of [N, C] fp32


def _mc_kernel_ok(mean: torch.Tensor, var: torch.Tensor, dim: int) -> bool:
    """The fused Monte-Carlo kernel handles fp32 CUDA [N, C <= 1024] logits reduced over the class axis."""
    return (mean.is_cuda and var.is_cuda and mean.dim() == 2 and mean.shape == var.shape and dim in (-1, 1) and
            mean.dtype == torch.float32 and var.dtype == torch.float32 and 0 < mean.shape[1] <= 1024 and mean.shape[0] > 0 and
            not (mean.requires_grad or var.requires_grad))


def _mc_accumulate(mean: torch.Tensor, var: torch.Tensor, num_samples: int, want_probs: bool = False,
                   want_entropy: bool = False):
    """sum_g softmax(mean + eps_g sqrt(var)) and / or sum_g H[softmax(...)] over ``num_samples`` draws (reference vlm.py:86-89,
    :145-149).  Every draw is its own ``torch.randn(var.shape)`` call on the default generator, like the reference's loop, so
    a shared seed reproduces its noise on the same device; G draws at a time are consumed by ONE kernel launch."""
    mean, var = mean.contiguous(), var.contiguous()
    n, c = mean.shape
    dev = mean.device
    acc_p = torch.zeros_like(mean) if want_probs else None
    acc_h = torch.zeros(n, dtype=torch.float32, device=dev) if want_entropy else None
    group = max(1, min(num_samples, _MC_NOISE_BYTES // (4 * n * c)))
    noise = torch.empty((group, n, c), dtype=torch.float32, device=dev)
    done = 0
    while done < num_samples:
        g = min(group, num_samples - done)
        for i in range(g):
            torch.randn((n, c), device=dev, out=noise[i])
        _lib.run(dev, "bvlm_mc_softmax_accumulate", _lib.ptr(mean), _lib.ptr(var), _lib.ptr(noise), n, c, g, _lib.ptr(acc_p),
                 _lib.ptr(acc_h), _lib.stream_ptr(dev))
        done += g
    return acc_p, acc_h


@dataclass
class ProbabilisticLogits:
    """Gaussian over logits: mean [N,C], var [N,C] (or full covariances [N,C,C]); reference vlm.py:63-204."""

    mean: torch.Tensor
    var: torch.Tensor

    # -- quirk kept on purpose: with 2-D ``var`` the reference takes ``var.diagonal(dim1=-2, dim2=-1)`` (vlm.py:76),
    #    i.e. the length-min(N,C) diagonal of the N x C matrix, not the per-pair variance.  `probit()` below is the
    #    canonical element-wise form the zero-shot script uses.
    def softmax(self, dim=-1, num_samples=400, chunk_size=10000, seed=None):
        if seed is not None:
            torch.manual_seed(seed)
        if num_samples == 0:
            diag = self.var.diagonal(dim1=-2, dim2=-1)
            return torch.softmax(self.mean / torch.sqrt(1 + torch.pi / 8 * diag), dim=dim)
        if self.var.ndim == 2:
            if _mc_kernel_ok(self.mean, self.var, dim):
                return _mc_accumulate(self.mean, self.var, num_samples, want_probs=True)[0] / num_samples
            std = self.var.sqrt()
            acc = torch.zeros_like(self.mean)
            for _ in range(num_samples):
                noise = torch.randn(std.shape, device=std.device) * std
                acc += torch.softmax(self.mean + noise, dim=dim)
            return acc / num_samples
        if self.var.ndim == 3:
            pieces = []
            n_chunks = math.ceil(self.mean.shape[0] / chunk_size)
            for mu, cov in zip(torch.chunk(self.mean, n_chunks, dim=0), torch.chunk(self.var, n_chunks, dim=0)):
                mvn = torch.distributions.MultivariateNormal(mu, covariance_matrix=cov)
                part = 0
                for _ in range(num_samples):
                    part = part + torch.softmax(mvn.sample(), dim=dim)
                pieces.append(part)
            return torch.cat(pieces, dim=0) / num_samples
        return torch.zeros_like(self.mean) / num_samples

    def probit(self) -> torch.Tensor:
        """Element-wise probit-adjusted softmax (scripts/zeroshot.py:119-120) as a CUDA row kernel."""
        return probit_softmax(self.mean, self.var)

    def sample_probas(self, num_samples: int, seed=None):
        """[N, num_samples, C] class probabilities from MC logit samples; reference vlm.py:105-139.

        Noise is drawn with ``torch.randn`` from the default generator of the tensors' device in [K, N, C] order,
        exactly like the reference, so a shared ``torch.manual_seed`` reproduces its draws on the same device.
        """
        if seed is not None:
            torch.manual_seed(seed)
        if self.var.ndim == 2:
            noise = torch.randn((num_samples,) + tuple(self.mean.shape), device=self.mean.device)
            draws = noise * self.var.sqrt() + self.mean  # fp32 result; the fused fp16 path is sample_probas_f16
            return torch.softmax(draws.permute(1, 0, 2), dim=2)
        if self.var.ndim == 3:
            mvn = torch.distributions.MultivariateNormal(self.mean, covariance_matrix=self.var)
            draws = torch.cat([mvn.sample((1,)) for _ in range(num_samples)], dim=0)
            return torch.softmax(draws.permute(1, 0, 2), dim=2)
        raise ValueError("Invalid variance tensor shape.")

    def sample_probas_f16(self, num_samples: int, seed=None) -> torch.Tensor:
        """Same draws as :meth:`sample_probas` but returns the fp16 tensor EPIG consumes (epig.py:324,334)."""
        if seed is not None:
            torch.manual_seed(seed)
        if self.var.ndim != 2:
            return self.sample_probas(num_samples).to(torch.float16)
        noise = torch.randn((num_samples,) + tuple(self.mean.shape), device=self.mean.device)
        return sample_probas_from_noise(self.mean, self.var, noise)

    def expected_aleatoric_entropy(self, num_samples=400, dim=-1):
        total = 0
        if self.var.ndim == 2 and num_samples > 0 and _mc_kernel_ok(self.mean, self.var, dim):
            return _mc_accumulate(self.mean, self.var, num_samples, want_entropy=True)[1] / num_samples
        if self.var.ndim == 2:
            std = self.var.sqrt()
            for _ in range(num_samples):
                p = torch.softmax(self.mean + torch.randn(self.var.shape, device=self.var.device) * std, dim=dim)
                total = total - (p * p.log()).sum(dim=dim)
        elif self.var.ndim == 3:
            mvn = torch.distributions.MultivariateNormal(self.mean, covariance_matrix=self.var)
            for _ in range(num_samples):
                p = torch.softmax(mvn.sample(), dim=dim)
                total = total - (p * p.log()).sum(dim=dim)
        return total / num_samples

    def cross_entropy(self, target, num_samples=400, reduction="sum"):
        ce = torch.nn.functional.cross_entropy
        if num_samples == 0:
            diag = self.var.diagonal(dim1=-2, dim2=-1)
            return ce(self.mean / torch.sqrt(1 + torch.pi / 8 * diag), target, reduction=reduction)
        total = 0
        if self.var.ndim == 2:
            diag_std = self.var.diagonal(dim1=-2, dim2=-1).sqrt()  # reference quirk, vlm.py:186
            for _ in range(num_samples):
                noise = torch.randn(self.var.shape, device=self.var.device) * diag_std
                total = total + ce(self.mean + noise, target, reduction=reduction)
        elif self.var.ndim == 3:
            mvn = torch.distributions.MultivariateNormal(self.mean, covariance_matrix=self.var)
            for _ in range(num_samples):
                total = total + ce(mvn.sample(), target, reduction=reduction)
        return total / num_samples

    def __getitem__(self, idx):
        return ProbabilisticLogits(mean=self.mean[idx], var=self.var[idx])

    def to(self, device):
        self.mean = self.mean.to(device)
        self.var = self.var.to(device)
        return self

    def detach(self):
        return ProbabilisticLogits(mean=self.mean.detach(), var=self.var.detach())

    def clone(self):
        return ProbabilisticLogits(mean=self.mean.clone(), var=self.var.clone())


# ----------------------------------------------------------------------------------------------------------------------
# covariance-side state prepared once per set_covariances()
# ----------------------------------------------------------------------------------------------------------------------
class _FactorOperand:
    """A_inv = W^T W with W lower triangular, stored as the fp16 GEMM operand the quadratic-form kernel reads."""

    def __init__(self, a_inv: torch.Tensor):
        _lib.require_cuda(a_inv, "A_inv")
        d = a_inv.shape[0]
        sym = a_inv.double()
        sym = 0.5 * (sym + sym.T)
        flipped = torch.flip(sym, dims=(0, 1))
        jitter = 0.0
        for attempt in range(6):
            try:
                chol = torch.linalg.cholesky(flipped + jitter * torch.eye(d, dtype=sym.dtype, device=sym.device))
                break
            except Exception:  # not numerically PD: add relative jitter and retry
                jitter = (10.0 ** (attempt - 9)) * float(sym.diagonal().abs().mean())
        else:
            raise RuntimeError("A_inv is not positive definite; cannot factor the covariance")
        # flip(chol) is upper triangular U with A_inv = U U^T, hence W = U^T is lower triangular and A_inv = W^T W
        w = torch.flip(chol, dims=(0, 1)).T.contiguous().float()
        wmax = float(w.abs().max())
        self.scale = 1.0 if wmax <= 0 else float(2.0 ** (9 - math.ceil(math.log2(wmax))))
        self.dA = d
        self.k_pad = int(lib.bvlm_padded_k(d))
        self.w16 = torch.empty((d, self.k_pad), dtype=torch.float16, device=a_inv.device)
        _lib.run(a_inv.device, "bvlm_factor_prepare", _lib.ptr(w), d, d, self.scale, _lib.ptr(self.w16), self.k_pad,
                 _lib.stream_ptr(a_inv.device))
        self._keep = w  # the conversion is asynchronous: keep the source alive with the operand


class _SideState:
    def __init__(self, cov: KroneckerFactorizedCovariance):
        self.factor = _FactorOperand(cov.A_inv)
        self.diag_b = cov.B_inv.diagonal().contiguous().float()


class CLIP(torch.nn.Module):
    """Similarity module with a Kronecker-factored Laplace posterior over both projection layers (vlm.py:567-710)."""

    source_projection_has_bias = False
    target_projection_has_bias = False

    def __init__(self, logit_scale: float, logit_bias: float = 0, source_covariance=None, target_covariance=None,
                 device: Optional[str] = None, precision: str = "fp16+fp8"):
        super().__init__()
        self.logit_scale = torch.nn.Parameter(torch.ones([], device=device) * logit_scale)
        self.logit_bias = torch.nn.Parameter(torch.ones([], device=device) * logit_bias)
        self.source_covariance = source_covariance
        self.target_covariance = target_covariance
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.precision = precision
        self._cov_version = 0
        self._side_cache = None    # (version, device, _SideState src, _SideState tgt, scalars)
        self._target_cache = None  # (key, T16, colA, colB)

    @property
    def device(self):
        return self.logit_scale.data.device

    def _logit_scale_args(self, dev: torch.device):
        """(host value, device pointer) of the log-space logit scale for the kernels.  When the parameter lives on the
        kernels' device the kernel reads it there (no host sync per call, and in-place `.data` updates -- the idiom of
        reference epig.py:230 -- are always seen); otherwise the host value is read now."""
        p = self.logit_scale.data
        if p.is_cuda and p.device == dev and p.dtype == torch.float32:
            return 0.0, _lib.ptr(p)
        return float(p), _lib.ptr(None)

    def set_covariances(self, source_covariance=None, target_covariance=None):
        def _own(cov):
            if cov is None:
                return None
            return KroneckerFactorizedCovariance(A_inv=cov.A_inv.clone().to(self.device),
                                                 B_inv=cov.B_inv.clone().to(self.device))

        self.source_covariance = _own(source_covariance)
        self.target_covariance = _own(target_covariance)
        self._cov_version += 1
        self._side_cache = None
        self._target_cache = None

    @classmethod
    def from_huggingface(cls, model_name: str, device: Optional[str] = None):
        from transformers import CLIPModel  # needs network/cache; outside the hot path

        ref = CLIPModel.from_pretrained(model_name)
        model = cls(logit_scale=ref.logit_scale.item())
        return model.to(device) if device is not None else model

    # ------------------------------------------------------------------------------------------------------------
    def _compute_logits(self, source_embeds: torch.Tensor, target_embeds: torch.Tensor):
        """Deterministic (MAP) logits ``exp(scale) * cos + bias``; differentiable torch expression (vlm.py:617-628)."""
        src = torch.nn.functional.normalize(source_embeds, p=2, dim=-1, eps=0.0)
        tgt = torch.nn.functional.normalize(target_embeds, p=2, dim=-1, eps=0.0)
        return src @ tgt.t() * self.logit_scale.exp() + self.logit_bias

    def _sides(self):
        dev = self.device
        cache = self._side_cache
        if cache is not None and cache[0] == self._cov_version and cache[1] == dev:
            return cache[2], cache[3], cache[4]
        if self.source_covariance is None or self.target_covariance is None:
            raise RuntimeError("set_covariances() must be called before the probabilistic forward")
        if self.source_covariance.A_inv.device != dev:
            self.source_covariance.to(dev)
            self.target_covariance.to(dev)
        src = _SideState(self.source_covariance)
        tgt = _SideState(self.target_covariance)
        scalars = torch.stack([src.diag_b.sum(), tgt.diag_b.sum(), (src.diag_b * tgt.diag_b).sum()]).tolist()
        self._side_cache = (self._cov_version, dev, src, tgt, scalars)
        return src, tgt, scalars

    def _target_side(self, target: EncoderResult, prec: int):
        src, tgt, (sum_beta, sum_delta, kappa) = self._sides()
        emb, act = target.embeds, target.activations
        # the cache keeps the target tensors alive and compares identity + version (a data_ptr alone could be recycled)
        key = (self._cov_version, prec, emb._version, act._version)
        tc = self._target_cache
        if tc is not None and tc[0] == key and tc[1] is emb and tc[2] is act:
            ready, made_on = tc[7], tc[8]
            cur = torch.cuda.current_stream(tc[3].device)
            if cur.cuda_stream != made_on:
                # prepared (asynchronously) on another stream: order this stream behind it and tell the caching allocator
                # that the operands are in use here too
                cur.wait_event(ready)
                for t in tc[3:7]:
                    if t is not None:
                        t.record_stream(cur)
            return tc[3:7]
        emb = _lib.rowmajor(_lib.require_cuda(emb.detach(), "target embeds"))
        act = _lib.rowmajor(_lib.require_cuda(act.detach(), "target activations"))
        c, d = emb.shape
        d_act = act.shape[1]
        bias = 1 if self.target_projection_has_bias else 0
        if d_act + bias != tgt.factor.dA:
            raise ValueError(f"target activations have {d_act}(+{bias}) features but A_inv is {tgt.factor.dA}^2")
        seg = int(lib.bvlm_padded_k(d))
        t16 = torch.empty((c, seg * (2 if prec == _lib.PREC_X3 else 1)), dtype=torch.float16, device=emb.device)
        t8 = (torch.empty((c, int(lib.bvlm_predictive_t8_cols(d))), dtype=torch.uint8, device=emb.device)
              if prec == _lib.PREC_X2F8 else None)
        c_pad = int(lib.bvlm_padded_cols(c))  # the epilogue reads whole 256-column tiles
        col_a = torch.empty(c_pad, dtype=torch.float32, device=emb.device)
        col_b = torch.empty(c_pad, dtype=torch.float32, device=emb.device)
        ws_bytes = lib.bvlm_predictive_target_workspace_bytes(c, d, d_act, bias)
        ws = _lib.workspace(emb.device, ws_bytes)
        _lib.run(
            emb.device, "bvlm_predictive_target_prepare", _lib.ptr(emb), c, d, emb.stride(0), _lib.ptr(act), d_act, act.stride(0), bias, _lib.ptr(tgt.factor.w16),
            tgt.factor.dA, tgt.factor.k_pad, tgt.factor.scale, _lib.ptr(src.diag_b), sum_delta, kappa, prec,
            _lib.ptr(t16), _lib.ptr(t8), _lib.ptr(col_a), _lib.ptr(col_b), _lib.ptr(ws), ws.numel(),
            _lib.stream_ptr(emb.device))
        cur = torch.cuda.current_stream(emb.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        self._target_cache = (key, target.embeds, target.activations, t16, t8, col_a, col_b, ready, cur.cuda_stream)
        return t16, t8, col_a, col_b

    def _smith_torch(self, source_results: EncoderResult, target_results: EncoderResult):
        """Differentiable on-device torch expression of the predictive, used only when an input requires grad
        (the 1 x C case of the online EPIG loop, reference epig.py:214-227)."""
        a_src, a_tgt = source_results.activations, target_results.activations
        if self.source_projection_has_bias:
            a_src = torch.cat([a_src, torch.ones_like(a_src[:, :1])], dim=-1)
        if self.target_projection_has_bias:
            a_tgt = torch.cat([a_tgt, torch.ones_like(a_tgt[:, :1])], dim=-1)
        e, t = source_results.embeds, target_results.embeds
        beta = self.source_covariance.B_inv.diagonal()
        delta = self.target_covariance.B_inv.diagonal()
        alpha = ((a_src @ self.source_covariance.A_inv) * a_src).sum(-1, keepdim=True)
        gamma = ((a_tgt @ self.target_covariance.A_inv) * a_tgt).sum(-1, keepdim=True)
        cov_e, cov_t = alpha * beta, gamma * delta
        sq_e, sq_t = e.square() + cov_e, t.square() + cov_t
        n_e, n_t = sq_e.sum(-1, keepdim=True), sq_t.sum(-1, keepdim=True)
        mean = (e / n_e.sqrt()) @ (t / n_t.sqrt()).t()
        var = (sq_e @ cov_t.t() + cov_e @ t.square().t()) / (n_e * n_t.t())
        s = self.logit_scale.exp()
        return ProbabilisticLogits(mean=mean * s, var=var * s.square())

    def _compute_probabilistic_logits_smith(self, source_results: EncoderResult, target_results: EncoderResult,
                                            compute_covariance: bool = False, return_probs: bool = False):
        """Expected value and variance of the cosine similarity between two Gaussian embeddings (vlm.py:630-684).

        NOTE (reference behaviour, kept): ``logit_bias`` is not added to the probabilistic mean (vlm.py:681-684).
        """
        if compute_covariance:
            raise NotImplementedError("Only the variances are supported for now.")
        needs_grad = torch.is_grad_enabled() and any(
            t.requires_grad for t in (source_results.embeds, source_results.activations, target_results.embeds,
                                      target_results.activations))
        if needs_grad:
            return self._smith_torch(source_results, target_results)

        emb = _lib.rowmajor(_lib.require_cuda(source_results.embeds.detach(), "source embeds"))
        act = _lib.rowmajor(_lib.require_cuda(source_results.activations.detach(), "source activations"))
        n = emb.shape[0]
        c = target_results.embeds.shape[0]
        mean = torch.empty((n, c), dtype=torch.float32, device=emb.device)
        var = torch.empty((n, c), dtype=torch.float32, device=emb.device)
        probs = torch.empty((n, c), dtype=torch.float32, device=emb.device) if return_probs else None
        self._smith_into(emb, act, target_results, mean, var, probs)
        out = ProbabilisticLogits(mean=mean, var=var)
        if return_probs:
            return out, probs
        return out

    def _smith_into(self, emb: torch.Tensor, act: torch.Tensor, target_results: EncoderResult, mean: torch.Tensor,
                    var: torch.Tensor, probs: Optional[torch.Tensor] = None):
        """Enqueue the predictive kernels for CUDA `emb` [n, D] / `act` [n, d_in] into preallocated `mean` / `var`
        ([n, C] fp32 views with unit column stride) on the current stream."""
        prec = _PRECISIONS[self.precision]
        n, d = emb.shape
        if d < _MIN_D_REDUCED_PRECISION:
            # the ~2^-15 / ~2^-11 products rely on their rounding errors averaging over the embedding dimension (error ~ s 2^-p /
            # sqrt(D)); below 128 dimensions nothing averages (scripts/fuzz_parity.py: up to 2x the logit tolerance at D = 3 .. 21)
            # and the third tensor-core pass costs nothing at that size
            prec = _lib.PREC_X3
        src, tgt, (sum_beta, sum_delta, kappa) = self._sides()
        t16, t8, col_a, col_b = self._target_side(target_results, prec)
        c = target_results.embeds.shape[0]
        if target_results.embeds.shape[1] != d:
            raise ValueError("source and target embeddings must share the embedding dimension")
        d_act = act.shape[1]
        bias = 1 if self.source_projection_has_bias else 0
        if d_act + bias != src.factor.dA:
            raise ValueError(f"source activations have {d_act}(+{bias}) features but A_inv is {src.factor.dA}^2")
        if n == 0:
            return
        if mean.stride(0) != var.stride(0) or (probs is not None and probs.stride(0) != mean.stride(0)):
            raise ValueError("mean / var / probs must share their row pitch")
        ls_host, ls_dev = self._logit_scale_args(emb.device)

        ws = _lib.workspace(emb.device, lib.bvlm_predictive_workspace_bytes(n, d, d_act, bias, prec))
        _lib.run(
            emb.device, "bvlm_predictive",
            _lib.ptr(emb), n, d, emb.stride(0), _lib.ptr(act), d_act, act.stride(0), bias, _lib.ptr(src.factor.w16),
            src.factor.dA, src.factor.k_pad, src.factor.scale, _lib.ptr(tgt.diag_b), sum_beta,
            ls_host, ls_dev, _lib.ptr(t16), _lib.ptr(t8), _lib.ptr(col_a), _lib.ptr(col_b), c, prec,
            _lib.ptr(mean), _lib.ptr(var), _lib.ptr(probs), mean.stride(0), _lib.ptr(ws), ws.numel(),
            _lib.stream_ptr(emb.device))

    def forward(self, source_embeds: Union[torch.Tensor, EncoderResult], target_embeds: Union[torch.Tensor, EncoderResult],
                map_estimate: bool = False):
        """[#source, #target] logits; ``EncoderResult`` inputs give a :class:`ProbabilisticLogits` (vlm.py:686-710)."""
        if isinstance(source_embeds, EncoderResult) and isinstance(target_embeds, EncoderResult):
            if map_estimate:
                logits = self._compute_logits(source_embeds.embeds, target_embeds.embeds)
                return ProbabilisticLogits(mean=logits, var=torch.zeros_like(logits))
            return self._compute_probabilistic_logits_smith(source_embeds, target_embeds)
        return self._compute_logits(source_embeds, target_embeds)

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def predict_host(self, image_outputs: EncoderResult, text_outputs: EncoderResult, batch_size: int = 2048,
                     return_probs: bool = False, out_pinned: bool = True, out=None):
        """End-to-end predictive for HOST-resident features (the `make_predictions` data flow, precompute.py:18-65).

        A three-stream pipeline: image batches are copied host -> device on a copy stream into double-buffered staging
        tensors, the kernels run on a compute stream, and mean / var go device -> host on a third stream, so PCIe in,
        the tensor cores and PCIe out overlap.  Text-side quantities are computed once, not per batch as the reference
        does (vlm.py:663).  `out=(mean, var[, probs])` lets a serving loop reuse its own (pinned) host buffers; otherwise
        fresh host tensors are allocated (pinned when `out_pinned`)."""
        from .hostmem import pinned_empty  # page-locked on the GPU's NUMA node

        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("predict_host needs the module on a CUDA device")
        n, c = len(image_outputs), len(text_outputs)
        d, d_act = image_outputs.embeds.shape[1], image_outputs.activations.shape[1]
        if out is not None:
            mean, var = out[0], out[1]
            probs = out[2] if return_probs else None
            if tuple(mean.shape) != (n, c) or tuple(var.shape) != (n, c):
                raise ValueError("out buffers must be [N, C]")
        else:
            mk = (lambda: pinned_empty((n, c), device=dev)) if out_pinned else (lambda: torch.empty((n, c)))
            mean, var = mk(), mk()
            probs = mk() if return_probs else None
        bs = max(1, min(batch_size, n))
        bounce = any(t.device.type == "cpu" and not t.is_pinned() for t in (image_outputs.embeds, image_outputs.activations))
        key = (dev, bs, c, d, d_act, return_probs, bounce)
        pipe = getattr(self, "_host_pipe", None)
        if pipe is None or pipe["key"] != key:
            f32 = dict(dtype=torch.float32, device=dev)
            pipe = {"key": key, "s_in": torch.cuda.Stream(dev), "s_cmp": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
                    "bufs": [dict(emb=torch.empty((bs, d), **f32), act=torch.empty((bs, d_act), **f32),
                                  mean=torch.empty((bs, c), **f32), var=torch.empty((bs, c), **f32),
                                  probs=torch.empty((bs, c), **f32) if return_probs else None,
                                  h_emb=pinned_empty((bs, d), device=dev) if bounce else None,
                                  h_act=pinned_empty((bs, d_act), device=dev) if bounce else None,
                                  computed=None, drained=None, loaded=None) for _ in range(2)]}
            self._host_pipe = pipe
        s_in, s_cmp, s_out = pipe["s_in"], pipe["s_cmp"], pipe["s_out"]
        cur = torch.cuda.current_stream(dev)
        for st_ in (s_in, s_cmp, s_out):
            st_.wait_stream(cur)
        with torch.cuda.stream(s_cmp):
            text_dev = EncoderResult(text_outputs.embeds.to(dev, non_blocking=True),
                                     text_outputs.activations.to(dev, non_blocking=True))
        for i, lo in enumerate(range(0, n, bs)):
            hi = min(n, lo + bs)
            m = hi - lo
            b = pipe["bufs"][i % 2]
            src_emb, src_act = image_outputs.embeds[lo:hi], image_outputs.activations[lo:hi]
            if bounce:  # pageable inputs: torch's host copy into a pinned bounce buffer, then a truly asynchronous transfer
                if b["loaded"] is not None:
                    b["loaded"].synchronize()  # the transfer of two batches ago has left the bounce buffer
                b["h_emb"][:m].copy_(src_emb)
                b["h_act"][:m].copy_(src_act)
                src_emb, src_act = b["h_emb"][:m], b["h_act"][:m]
            with torch.cuda.stream(s_in):
                if b["computed"] is not None:
                    s_in.wait_event(b["computed"])  # the kernels of two batches ago have consumed this staging buffer
                b["emb"][:m].copy_(src_emb, non_blocking=True)
                b["act"][:m].copy_(src_act, non_blocking=True)
                loaded = torch.cuda.Event()
                loaded.record(s_in)
                b["loaded"] = loaded
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(loaded)
                if b["drained"] is not None:
                    s_cmp.wait_event(b["drained"])  # the previous results in this buffer are on the host
                self._smith_into(b["emb"][:m], b["act"][:m], text_dev, b["mean"][:m], b["var"][:m],
                                 b["probs"][:m] if return_probs else None)
                b["computed"] = torch.cuda.Event()
                b["computed"].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(b["computed"])
                mean[lo:hi].copy_(b["mean"][:m], non_blocking=True)
                var[lo:hi].copy_(b["var"][:m], non_blocking=True)
                if return_probs:
                    probs[lo:hi].copy_(b["probs"][:m], non_blocking=True)
                b["drained"] = torch.cuda.Event()
                b["drained"].record(s_out)
        s_out.synchronize()
        cur.wait_stream(s_cmp)
        for b in pipe["bufs"]:
            b["computed"] = b["drained"] = b["loaded"] = None
        res = ProbabilisticLogits(mean=mean, var=var)
        return (res, probs) if return_probs else res


class SIGLIP(CLIP):
    """Same predictive with bias-augmented activations on both projection layers (vlm.py:712-714)."""

    source_projection_has_bias = True
    target_projection_has_bias = True

    @classmethod
    def from_huggingface(cls, model_name: str, device: Optional[str] = None):
        from transformers import SiglipModel

        ref = SiglipModel.from_pretrained(model_name)
        model = cls(logit_scale=ref.logit_scale.item(), logit_bias=ref.logit_bias.item())
        return model.to(device) if device is not None else model
