"""KFAC / GGN estimation entry points of BayesVLM on B200 kernels.

Mirrors the reference's ``bayesvlm/hessians.py`` (names, arguments, error behaviour) plus ``kfac_ggn`` from
``scripts/hessian_estimation.py:26-109``.  The two analytic GGN functions and the A-factor SYRK run as tcgen05 GEMM
pipelines (``csrc/kfac.cu``); covariance assembly / inversion and the prior-precision optimiser stay in torch (they are
O(d^3) on <= 3073-dimensional matrices and not on the data path).
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass
from pathlib import Path
from typing import Literal, Optional, Tuple

import torch

from . import _lib
from ._lib import lib


# ----------------------------------------------------------------------------------------------------------------------
# K2 / K3: analytic GGN of the contrastive losses w.r.t. the source embeddings
# ----------------------------------------------------------------------------------------------------------------------
# Source batches below this size evaluate the logits with the hi/lo split (~fp32): a single fp16 pass perturbs every
# logit by ~ s * 2^-11 / sqrt(D), which only averages out over thousands of sources (the KFAC class batches).
GGN_SPLIT_LOGITS_BELOW = 8192


def _ggn(source: torch.Tensor, target: torch.Tensor, logit_scale, logit_bias, siglip: bool,
         out: Optional[torch.Tensor] = None, accumulate: bool = False, precision: Optional[str] = None) -> torch.Tensor:
    source = _lib.rowmajor(_lib.require_cuda(source, "source embeddings"))
    target = _lib.rowmajor(_lib.require_cuda(target, "target embeddings"))
    if source.dim() != 2 or target.dim() != 2:
        raise ValueError("embeddings must be 2-D [rows, D]")
    b, d = source.shape
    c = target.shape[0]
    dev = source.device
    if out is None:
        out = torch.zeros((d, d), dtype=torch.float32, device=dev)
        accumulate = False
    if b == 0 or c == 0:
        if not accumulate:
            out.zero_()
        return out
    if precision is None:
        prec = _lib.PREC_X3 if b < GGN_SPLIT_LOGITS_BELOW else _lib.PREC_X1
    else:
        prec = {"fp16": _lib.PREC_X1, "fp16x3": _lib.PREC_X3}[precision]
    ws = _lib.workspace(dev, lib.bvlm_ggn_workspace_bytes(b, c, d, prec), tag="ggn")
    ls = float(logit_scale)
    if siglip:
        _lib.run(dev, "bvlm_ggn_siglip", _lib.ptr(source), b, source.stride(0), _lib.ptr(target), c, target.stride(0), d, ls,
                 float(logit_bias), prec, _lib.ptr(out), out.stride(0), int(accumulate), _lib.ptr(ws), ws.numel(),
                 _lib.stream_ptr(dev))
    else:
        _lib.run(dev, "bvlm_ggn_infonce", _lib.ptr(source), b, source.stride(0), _lib.ptr(target), c, target.stride(0), d, ls,
                 prec, _lib.ptr(out), out.stride(0), int(accumulate), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    return out


def compute_hessian_analytic_InfoNCE(source_embeds: torch.Tensor, target_embeds: torch.Tensor, logit_scale: torch.Tensor):
    """sum_b s^2 J_b Yh^T (diag p_b - p_b p_b^T) Yh J_b^T with p_b = softmax(s xh_b Yh^T); reference hessians.py:10-48.

    Inputs are un-normalised [B, D] / [C, D] fp32 CUDA tensors, ``logit_scale`` is in log space; returns [D, D] fp32.
    """
    return _ggn(source_embeds, target_embeds, logit_scale, 0.0, siglip=False)


def compute_hessian_analytic_SigLIP(x_batch: torch.Tensor, indices_batch: torch.Tensor, y: torch.Tensor,
                                    logit_scale: torch.Tensor, logit_bias: torch.Tensor, chunk_size_j: int = None):
    """Hessian of the SigLIP loss w.r.t. x summed over the batch; reference hessians.py:50-117.

    ``indices_batch`` only flips the sign inside sigma(.)(1 - sigma(.)), an even function, so it does not enter the
    result; ``chunk_size_j`` is the reference's memory workaround (the sum is chunk invariant) and is ignored.
    """
    assert x_batch.shape[1] == y.shape[1], "The input and output dimensions must be the same"
    del indices_batch, chunk_size_j
    return _ggn(x_batch, y, logit_scale, logit_bias, siglip=True)


def syrk_accumulate(acts: torch.Tensor, out: Optional[torch.Tensor] = None, append_one: bool = False,
                    alpha: float = 1.0, accumulate: bool = False) -> torch.Tensor:
    """K1: ``out (+)= alpha * [acts 1?]^T [acts 1?]`` (scripts/hessian_estimation.py:99-104); fp16 operands with exact per-feature
    power-of-two scaling, fp32 accumulation."""
    acts = _lib.rowmajor(_lib.require_cuda(acts, "activations"))
    n, d = acts.shape
    d_a = d + (1 if append_one else 0)
    own = out is None
    if own:  # (row pitch a multiple of 16 bytes: split-K partials go through TMA bulk reductions; returned contiguous)
        out = torch.zeros((d_a, (d_a + 3) // 4 * 4), dtype=torch.float32, device=acts.device)[:, :d_a]
        accumulate = False
    if n == 0:
        if not accumulate:
            out.zero_()
        return out.contiguous() if own else out
    ws = _lib.workspace(acts.device, lib.bvlm_syrk_workspace_bytes(n, d, int(append_one), _lib.PREC_X1), tag="syrk")
    _lib.run(acts.device, "bvlm_syrk_f32acc", _lib.ptr(acts), n, d, acts.stride(0), int(append_one), _lib.PREC_X1,
             _lib.ptr(out), out.stride(0), float(alpha), int(accumulate), _lib.ptr(ws), ws.numel(),
             _lib.stream_ptr(acts.device))
    return out.contiguous() if own else out


# ----------------------------------------------------------------------------------------------------------------------
# K0: the KFAC accumulation loop
# ----------------------------------------------------------------------------------------------------------------------
def class_batch_schedule(num_class_batches: int, rank: int, world_size: int):
    """Class batches owned by ``rank``: a CONTIGUOUS block (sizes differ by at most one).  Each class batch carries its own
    paired targets, so nothing is replicated and no data-path collective is needed until the final all-reduce; contiguous
    blocks let a rank run its K1 SYRK over all of its rows in one launch (A is a plain sum over rows)."""
    q, r = divmod(num_class_batches, world_size)
    lo = rank * q + min(rank, r)
    return list(range(lo, lo + q + (1 if rank < r else 0)))


def reduce_factors(A: torch.Tensor, B: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """One all-reduce(sum) over the concatenated [A || B] buffer of a modality (no-op without torch.distributed)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return A, B
    flat = torch.cat([A.reshape(-1), B.reshape(-1)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat[: A.numel()].view_as(A), flat[A.numel():].view_as(B)


_KFAC_STAGING = {}


def _kfac_staging(dev: torch.device, num_classes: int, inputs):
    """Two sets of device staging buffers (+ pinned bounce buffers for pageable inputs) per (device, shapes), a copy stream,
    and the events that order their reuse; kept across calls."""
    from .hostmem import pinned_empty

    need_bounce = tuple(t.device.type == "cpu" and not t.is_pinned() for t in inputs)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), num_classes,
           tuple(t.shape[1] for t in inputs), need_bounce)
    st = _KFAC_STAGING.get(key)
    if st is None:
        _KFAC_STAGING.clear()  # one configuration at a time: the buffers are hundreds of MB
        st = {"stream": torch.cuda.Stream(dev),
              "device": [[torch.empty((num_classes, t.shape[1]), dtype=torch.float32, device=dev) for t in inputs]
                         for _ in range(2)],
              "bounce": [[pinned_empty((num_classes, t.shape[1]), torch.float32, dev) if nb else None
                          for t, nb in zip(inputs, need_bounce)] for _ in range(2)]}
        _KFAC_STAGING[key] = st
    st["loaded"], st["consumed"] = [None, None], [None, None]
    return st


@torch.no_grad()
def kfac_ggn(vlm, num_classes: int, batch_size: int, source_embeds: torch.Tensor, source_activations: torch.Tensor,
             target_embeds: torch.Tensor, device: str, likelihood: Literal["info_nce", "siglip"],
             siglip_chunk_size_j: int = 8000, group=None, distributed: bool = False
             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """K-FAC (last layer) of the GGN of ``-log p(target | source)``; reference scripts/hessian_estimation.py:26-109.

    Reference behaviour that is reproduced: the class-batch remainder is dropped (:55); inside a class batch the
    data-batch remainder is dropped for B only (:71 vs :100); targets of a class batch are the paired rows (:67);
    SigLIP appends a ones column before ``act^T act`` (:103-104); both factors are divided by sqrt(n) (:106-108);
    A is returned on ``device`` and B on the CPU (:84,:97,:100).

    New: the whole class batch is processed by one kernel pipeline (the per-row softmax makes the result independent of
    the reference's data batching).  ``distributed=True`` (explicit opt-in; the default is the reference's single-process
    semantics whatever process group happens to be initialised) shards the class batches over the ranks of ``group`` in
    contiguous blocks and sums the factors with ONE all-reduce of [A || B].  CONTRACT: every rank must then call with the
    IDENTICAL full dataset (each rank reads only its own class batches); ranks holding different shards, or a call from a
    single rank, are not supported in this mode.
    """
    del siglip_chunk_size_j
    if likelihood not in ("info_nce", "siglip"):
        raise ValueError(f"Invalid likelihood: {likelihood}, must be one of ['info_nce', 'siglip'].")
    num_class_batches = len(target_embeds) // num_classes
    if num_class_batches == 0:
        raise ValueError(f"To few datapoints for K-FAC approximation. Need at least {num_classes} datapoints.")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("kfac_ggn runs on CUDA (sm_100a) only; there is no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    import torch.distributed as dist

    use_dist = bool(distributed)
    if use_dist and not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("kfac_ggn(distributed=True) needs an initialised torch.distributed process group")
    rank = dist.get_rank(group) if use_dist else 0
    world = dist.get_world_size(group) if use_dist else 1

    siglip = likelihood == "siglip"
    logit_scale = float(vlm.logit_scale.detach())
    logit_bias = float(vlm.logit_bias.detach())
    d_in = source_activations.shape[1] + (1 if siglip else 0)
    d_emb = source_embeds.shape[1]
    with torch.cuda.device(dev):
        # row pitch padded to a multiple of 16 bytes (the bias-augmented SigLIP factor is 769 / 3073 wide): the SYRK then adds
        # its split-K partial tiles with TMA bulk reductions instead of per-element atomics; the result below is contiguous
        A = torch.zeros((d_in, (d_in + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :d_in]
        B = torch.zeros((d_emb, d_emb), dtype=torch.float32, device=dev)

    # Host-resident inputs (the reference's calling convention) are staged one class batch AHEAD on a copy stream, so that
    # the PCIe transfer of batch i+1 (235 MB at config 2, ~4.3 ms) overlaps the kernels of batch i (~4.6 ms).  Two sets of
    # persistent device staging buffers are reused across calls (fresh per-batch allocations on the copy stream make the
    # caching allocator fall back to cudaMalloc while their cross-stream frees are pending: measured 4x slower).
    # Pageable host tensors additionally go through pinned bounce buffers filled by torch's (multi-threaded) host copy
    # while the GPU works on the previous batch -- the driver's own pageable path is a synchronous, single-threaded copy.
    inputs = (target_embeds, source_embeds, source_activations)
    staged = any(t.device.type == "cpu" for t in inputs)
    main_stream = torch.cuda.current_stream(dev)
    stage = _kfac_staging(dev, num_classes, inputs) if staged else None
    # K1 (A += act^T act) is independent of the GGN pipeline: it runs on a side stream and fills the SMs the persistent GGN
    # kernels leave idle at their tails.  Device-resident activations of the rank's contiguous block go through ONE launch.
    syrk_stream = _kfac_side_stream(dev)

    def fetch(i, slot):
        lo, hi = i * num_classes, (i + 1) * num_classes
        if not staged:
            return tuple(t[lo:hi].to(dev, dtype=torch.float32) for t in inputs) + (None,)
        sources = []
        for k, t in enumerate(inputs):
            host = stage["bounce"][slot][k]
            if t.device.type == "cpu" and not t.is_pinned():
                if stage["loaded"][slot] is not None:
                    stage["loaded"][slot].synchronize()  # the previous transfer out of this bounce buffer has finished
                host.copy_(t[lo:hi])
                sources.append(host)
            else:
                sources.append(t[lo:hi])
        with torch.cuda.stream(stage["stream"]):
            if stage["consumed"][slot] is not None:
                stage["stream"].wait_event(stage["consumed"][slot])  # the kernels of two batches ago are done with this set
            # targets and source embeddings first: the GGN pipeline starts as soon as they have landed, the activations
            # (only the side-stream SYRK reads them) follow under its first kernels
            events = []
            for k, (dst, src_t) in enumerate(zip(stage["device"][slot], sources)):
                dst.copy_(src_t, non_blocking=True)
                if k >= 1:
                    ev = torch.cuda.Event()
                    ev.record(stage["stream"])
                    events.append(ev)
        stage["loaded"][slot] = events[-1]
        return tuple(stage["device"][slot]) + (tuple(events),)

    schedule = list(class_batch_schedule(num_class_batches, rank, world))
    syrk_stream.wait_stream(main_stream)
    whole_block = (not staged and len(schedule) > 0 and source_activations.device == dev and
                   source_activations.dtype == torch.float32)
    if whole_block:
        with torch.cuda.stream(syrk_stream):
            syrk_accumulate(source_activations[schedule[0] * num_classes:(schedule[-1] + 1) * num_classes], out=A,
                            append_one=siglip, accumulate=True)
    if staged:
        stage["stream"].wait_stream(main_stream)
        if stage.get("last_call_done") is not None:  # a previous call (possibly on another stream) may still read the buffers
            stage["stream"].wait_event(stage["last_call_done"])
    pending = fetch(schedule[0], 0) if schedule else None
    for k, _ in enumerate(schedule):
        tgt, src, act, ready = pending
        pending = fetch(schedule[k + 1], (k + 1) % 2) if k + 1 < len(schedule) else None
        if ready is not None:  # staged inputs: (embeddings landed, activations landed)
            main_stream.wait_event(ready[0])
        if not whole_block:
            syrk_stream.wait_stream(main_stream)  # (orders the staging-buffer reuse of two batches ago as well)
            if ready is not None:
                syrk_stream.wait_event(ready[1])
            with torch.cuda.stream(syrk_stream):
                syrk_accumulate(act, out=A, append_one=siglip, accumulate=True)
        used = (num_classes // batch_size) * batch_size  # data-batch remainder never reaches B
        if used > 0:
            _ggn(src[:used], tgt, logit_scale, logit_bias, siglip=siglip, out=B, accumulate=True)
        if staged:
            main_stream.wait_stream(syrk_stream)  # the staging set is free once BOTH consumers are done with it
            done = torch.cuda.Event()
            done.record(main_stream)
            stage["consumed"][k % 2] = done
    main_stream.wait_stream(syrk_stream)
    if staged:
        stage["last_call_done"] = torch.cuda.Event()
        stage["last_call_done"].record(main_stream)

    if use_dist and world > 1:
        A, B = reduce_factors(A, B, group)
    n = num_class_batches * num_classes
    A = A / math.sqrt(n)
    B = B / math.sqrt(n)
    return A, B.cpu()


_KFAC_SIDE_STREAMS = {}


def _kfac_side_stream(dev: torch.device):
    st = _KFAC_SIDE_STREAMS.get(dev.index)
    if st is None:
        st = _KFAC_SIDE_STREAMS[dev.index] = torch.cuda.Stream(dev)
    return st


# ----------------------------------------------------------------------------------------------------------------------
# C1: covariance plumbing (torch; reference hessians.py:120-217)
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class KroneckerFactorizedCovariance:
    A_inv: torch.Tensor
    B_inv: torch.Tensor

    def clone(self):
        return KroneckerFactorizedCovariance(A_inv=self.A_inv.clone(), B_inv=self.B_inv.clone())

    def to(self, device):
        self.A_inv = self.A_inv.to(device)
        self.B_inv = self.B_inv.to(device)
        return self


def _regularised_inverse(F: torch.Tensor, sqrt_n, sqrt_lmbda) -> torch.Tensor:
    eye = torch.eye(F.size(0), device=F.device, dtype=F.dtype)
    return torch.linalg.inv(F * sqrt_n + sqrt_lmbda * eye)


def _compute_covariance(A: torch.Tensor, B: torch.Tensor, n: torch.Tensor, lmbda: torch.Tensor):
    """Both Kronecker factors get sqrt(n) and sqrt(lambda) (reference hessians.py:170-184)."""
    sqrt_n, sqrt_l = torch.sqrt(n), torch.sqrt(lmbda)
    return KroneckerFactorizedCovariance(A_inv=_regularised_inverse(A, sqrt_n, sqrt_l),
                                         B_inv=_regularised_inverse(B, sqrt_n, sqrt_l))


def compute_covariances(A_img: torch.Tensor, B_img: torch.Tensor, A_txt: torch.Tensor, B_txt: torch.Tensor, info: dict):
    def scalar(key, like):
        return torch.tensor(info[key], dtype=like.dtype, device=like.device)

    cov_img = _compute_covariance(A_img, B_img, scalar("n_img", A_img), scalar("lambda_img", A_img))
    cov_txt = _compute_covariance(A_txt, B_txt, scalar("n_txt", A_txt), scalar("lambda_txt", A_txt))
    return cov_img, cov_txt


@dataclass
class FactorSpectrum:
    """Eigendecomposition ``F = V diag(e) V^T`` of one (symmetrised) Kronecker factor, fp64 on the factor's device.

    ONE ``syevd`` per factor serves both consumers of the online loop: the prior-precision optimiser needs
    ``logdet(sqrt(n) F + sqrt(lambda) I) = sum_k log(sqrt(n) e_k + sqrt(lambda))`` for many lambdas, and the covariance needs
    ``(sqrt(n) F + sqrt(lambda) I)^-1 = V diag(1 / (sqrt(n) e_k + sqrt(lambda))) V^T`` for the optimised one -- a d x d GEMM
    instead of a fresh LU factorisation + inverse (reference hessians.py:170-184 and :219-265 factorise 2 + 2 * num_steps times).
    """

    evals: torch.Tensor
    evecs: torch.Tensor

    @staticmethod
    def of(F: torch.Tensor, device=None) -> "FactorSpectrum":
        F64 = F.detach().to(device if device is not None else F.device).double()
        evals, evecs = torch.linalg.eigh(0.5 * (F64 + F64.T))
        return FactorSpectrum(evals=evals, evecs=evecs)

    def logdet_regularised(self, sqrt_n, sqrt_l) -> torch.Tensor:
        return torch.log(self.evals * sqrt_n + sqrt_l).sum()

    def inverse_regularised(self, sqrt_n: float, sqrt_l: float, dtype=torch.float32) -> torch.Tensor:
        d = 1.0 / (self.evals * sqrt_n + sqrt_l)
        return ((self.evecs * d) @ self.evecs.T).to(dtype)


def covariance_from_spectra(spec_A: FactorSpectrum, spec_B: FactorSpectrum, n: float, lmbda: float,
                            dtype=torch.float32) -> KroneckerFactorizedCovariance:
    """``_compute_covariance`` (reference hessians.py:170-184) from precomputed factor spectra."""
    sqrt_n, sqrt_l = math.sqrt(float(n)), math.sqrt(float(lmbda))
    return KroneckerFactorizedCovariance(A_inv=spec_A.inverse_regularised(sqrt_n, sqrt_l, dtype),
                                         B_inv=spec_B.inverse_regularised(sqrt_n, sqrt_l, dtype))


def load_hessians(la_dir: str, tag: Literal["img", "txt"], return_info: bool = False):
    A = torch.load(Path(la_dir) / f"A_{tag}_analytic.pt", map_location="cpu")
    B = torch.load(Path(la_dir) / f"B_{tag}_analytic.pt", map_location="cpu")
    if not return_info:
        return A, B
    with open(Path(la_dir) / "prior_precision_analytic.json") as f:
        info = json.load(f)
    return A, B, info


def load_covariances(la_dir: str, return_info: bool = False):
    A_img, B_img, info = load_hessians(la_dir, "img", return_info=True)
    A_txt, B_txt = load_hessians(la_dir, "txt")
    covs = []
    for A, B, tag in ((A_img, B_img, "img"), (A_txt, B_txt, "txt")):
        sn, sl = math.sqrt(info[f"n_{tag}"]), math.sqrt(info[f"lambda_{tag}"])
        covs.append(KroneckerFactorizedCovariance(A_inv=_regularised_inverse(A, sn, sl),
                                                  B_inv=_regularised_inverse(B, sn, sl)))
    if return_info:
        return covs[0], covs[1], info
    return covs[0], covs[1]


# ----------------------------------------------------------------------------------------------------------------------
# prior precision (reference hessians.py:219-280)
# ----------------------------------------------------------------------------------------------------------------------
def l2_norm_squared(module: torch.nn.Module):
    return sum((p ** 2).sum() for p in module.parameters())


def num_params(module: torch.nn.Module):
    return sum(p.numel() for p in module.parameters())


def compute_log_prior(l2_norm_squared: torch.Tensor, num_params: int, lmbda: float):
    return -0.5 * lmbda * l2_norm_squared + 0.5 * num_params * torch.log(lmbda)


def compute_log_det_kfac(A: torch.Tensor, B: torch.Tensor):
    """NOTE reference quirk: logdet(A) * p + logdet(B) * q with p = A.shape[0], q = B.shape[0] (hessians.py:276-280)."""
    return torch.logdet(A) * A.shape[0] + torch.logdet(B) * B.shape[0]


def optimize_prior_precision(projection: torch.nn.Module, A: torch.Tensor, B: torch.Tensor, lmbda_init: float, n: float,
                             lr: float, num_steps: int, device: str, retain_graph: bool = False, verbose: bool = False,
                             spectra: Optional[Tuple["FactorSpectrum", "FactorSpectrum"]] = None):
    """Adam ascent on log(lambda) of ``log_prior - logdet_kfac`` (no 1/2 on the log-det: reference hessians.py:260).

    The reference re-factorises both d x d matrices every step; here each factor is diagonalised once
    (``logdet(F sqrt(n) + sqrt(lambda) I) = sum_k log(sqrt(n) e_k + sqrt(lambda))``), which makes a step O(d).
    With ``spectra=(FactorSpectrum.of(A), FactorSpectrum.of(B))`` no factorisation happens here at all, and the same
    spectra give the covariance of the optimised lambda through ``covariance_from_spectra``.
    """
    del retain_graph, verbose
    for p in projection.parameters():
        p.requires_grad = False
    norm_sq = l2_norm_squared(projection).detach().to(device)
    n_par = num_params(projection)

    def spectrum(F):
        F64 = F.to(device).double()
        return torch.linalg.eigvalsh(0.5 * (F64 + F64.T))

    # ``spectra``: eigendecompositions the caller already holds (and will reuse for the covariance) -- no factorisation here
    eig_a, eig_b = (spectra[0].evals, spectra[1].evals) if spectra is not None else (spectrum(A), spectrum(B))
    # The objective is a scalar function of log(lambda) of the two spectra: its gradient is closed form, so the Adam steps
    # (torch.optim.Adam defaults, maximize=True, an fp32 parameter like the reference's) run on the host in microseconds
    # instead of ~0.5 ms of tiny device launches each (one small device -> host copy of the eigenvalues):
    #   d/dlog(l) [ -l |w|^2 / 2 + P log(l) / 2 ]                   = -l |w|^2 / 2 + P / 2
    #   d/dlog(l) [ p sum_k log(sqrt(n) e_k + sqrt(l)) + q sum_k .. ] = sqrt(l) / 2 * [ p sum_k 1 / (sqrt(n) e_k + sqrt(l)) + q .. ]
    import numpy as np

    ea = (eig_a.double() * math.sqrt(n)).cpu().numpy()
    eb = (eig_b.double() * math.sqrt(n)).cpu().numpy()
    w2 = float(norm_sq)
    p_dim, q_dim = A.shape[0], B.shape[0]
    f32 = np.float32
    x = f32(math.log(f32(lmbda_init)))
    m, v = f32(0.0), f32(0.0)
    b1, b2, eps = 0.9, 0.999, 1e-8
    for t in range(1, num_steps + 1):
        lam = float(np.exp(x, dtype=np.float32))
        sl = math.sqrt(lam)
        grad = (-0.5 * lam * w2 + 0.5 * n_par) - 0.5 * sl * (p_dim * float(np.sum(1.0 / (ea + sl))) + q_dim * float(np.sum(1.0 / (eb + sl))))
        g = f32(-grad)  # maximize=True: Adam descends on the negated gradient
        m = f32(b1) * m + f32(1.0 - b1) * g
        v = f32(b2) * v + f32(1.0 - b2) * g * g
        step = f32(lr / (1.0 - b1 ** t))
        denom = f32(np.sqrt(v) / f32(math.sqrt(1.0 - b2 ** t))) + f32(eps)
        x = f32(x - step * (m / denom))
    return torch.tensor(float(np.exp(x, dtype=np.float32)), dtype=torch.float32, device=device)
