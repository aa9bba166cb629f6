"""K0/K1/K2/K3 parity: CUDA path (Python mirror -> ctypes -> C ABI) against the reference-pinned oracle and the
golden outputs of the reference's own kfac_ggn / compute_hessian_analytic_* (tests/golden/make_golden.py).

Tolerances (BASELINE north_star: 1e-3 relative for factors, fp32 accumulate):
    ||F - F_ref||_F <= 1e-3 ||F_ref||_F   and   max|F - F_ref| <= 1e-3 max|F_ref|
"""
import math

import numpy as np
import pytest
import torch

from oracle import laplace_oracle as O

pytestmark = pytest.mark.gpu
LS = math.log(100.0)


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _check_factor(F, ref, tol=1e-3):
    F = F.detach().double().cpu().numpy()
    ref = np.asarray(ref, np.float64)
    assert F.shape == ref.shape
    assert np.isfinite(F).all()
    fro = np.linalg.norm(F - ref) / np.linalg.norm(ref)
    mx = np.abs(F - ref).max() / np.abs(ref).max()
    assert fro <= tol, f"frobenius rel err {fro:.3g}"
    assert mx <= tol, f"max abs err / max|ref| {mx:.3g}"


def _paired(gen, n, d):
    z = torch.randn(n, d, generator=gen)
    return z + 1.5 * torch.randn(n, d, generator=gen), z + 1.5 * torch.randn(n, d, generator=gen)


# ----------------------------------------------------------------------------------------------------------------------
# K2 / K3
# ----------------------------------------------------------------------------------------------------------------------
def test_infonce_golden(golden):
    from bayesvlm_b200.hessians import compute_hessian_analytic_InfoNCE

    H = compute_hessian_analytic_InfoNCE(_cuda(golden["ggn_X"]), _cuda(golden["ggn_Y"]), torch.tensor(LS))
    _check_factor(H, golden["ggn_infonce_H64"])
    _check_factor(H, golden["ggn_infonce_H"])


def test_siglip_golden(golden):
    from bayesvlm_b200.hessians import compute_hessian_analytic_SigLIP

    s, b = golden["ggn_siglip_params"]
    X = _cuda(golden["ggn_X"])
    idx = torch.arange(X.shape[0], device="cuda")
    H = compute_hessian_analytic_SigLIP(X, idx, _cuda(golden["ggn_Y"]), torch.tensor(float(s)), torch.tensor(float(b)),
                                        chunk_size_j=16)
    _check_factor(H, golden["ggn_siglip_H64"])
    with pytest.raises(AssertionError):
        compute_hessian_analytic_SigLIP(X, idx, _cuda(golden["ggn_Y"])[:, :-1].contiguous(), torch.tensor(1.0),
                                        torch.tensor(0.0))


@pytest.mark.parametrize("shape", [(5, 300, 64), (129, 257, 96), (640, 1024, 512), (1000, 2048, 768), (384, 4096, 1024)])
@pytest.mark.parametrize("siglip", [False, True])
def test_ggn_vs_oracle(shape, siglip):
    from bayesvlm_b200.hessians import compute_hessian_analytic_InfoNCE, compute_hessian_analytic_SigLIP

    B, C, D = shape
    gen = torch.Generator().manual_seed(B * 31 + C + D + int(siglip))
    Xa, Ya = _paired(gen, max(B, C), D)
    X, Y = Xa[:B].contiguous(), Ya[:C].contiguous()
    if siglip:
        ls, lb = 4.765, -12.93
        H = compute_hessian_analytic_SigLIP(X.cuda(), torch.arange(B).cuda(), Y.cuda(), torch.tensor(ls), torch.tensor(lb))
        ref = O.siglip_ggn_collapsed(X.numpy(), Y.numpy(), ls, lb)
    else:
        H = compute_hessian_analytic_InfoNCE(X.cuda(), Y.cuda(), torch.tensor(LS))
        ref = O.infonce_ggn_collapsed(X.numpy(), Y.numpy(), LS)
    _check_factor(H, ref)
    Hn = H.double().cpu().numpy()
    assert np.abs(Hn - Hn.T).max() <= 2e-4 * np.abs(Hn).max()


@pytest.mark.parametrize("siglip", [False, True])
def test_ggn_class_batch_properties(siglip):
    """Full reference class batch (B = C = 32768, D = 512): additivity over source rows (the GGN is a plain sum over
    sources: H(X1 u X2, Y) = H(X1, Y) + H(X2, Y)) and agreement with an fp32 torch evaluation of the collapsed form
    on a row subset (the oracle is too slow at this size)."""
    from bayesvlm_b200.hessians import _ggn

    n, D = 32768, 512
    gen = torch.Generator().manual_seed(77 + int(siglip))
    X, Y = _paired(gen, n, D)
    X, Y = X.cuda(), Y.cuda()
    ls, lb = (4.765, -12.93) if siglip else (LS, 0.0)
    H = _ggn(X, Y, ls, lb, siglip).clone()
    assert torch.isfinite(H).all()
    cut = 12345
    H1 = _ggn(X[:cut], Y, ls, lb, siglip).clone()
    H12 = _ggn(X[cut:], Y, ls, lb, siglip, out=H1, accumulate=True)
    rel = ((H12 - H).norm() / H.norm()).item()
    assert rel <= 2e-4, rel
    rows = torch.randperm(n, generator=gen)[:256]
    Hs = _ggn(X[rows.cuda()].contiguous(), Y, ls, lb, siglip)
    ref = (O.siglip_ggn_collapsed(X[rows.cuda()].cpu().numpy(), Y.cpu().numpy(), ls, lb) if siglip else
           O.infonce_ggn_collapsed(X[rows.cuda()].cpu().numpy(), Y.cpu().numpy(), ls))
    _check_factor(Hs, ref)


def test_ggn_edge_cases():
    from bayesvlm_b200.hessians import compute_hessian_analytic_InfoNCE

    gen = torch.Generator().manual_seed(5)
    X, Y = torch.randn(3, 40, generator=gen), torch.randn(17, 40, generator=gen)
    H = compute_hessian_analytic_InfoNCE(X[:0].cuda(), Y.cuda(), torch.tensor(LS))
    assert H.shape == (40, 40) and (H == 0).all()
    H1 = compute_hessian_analytic_InfoNCE(X[:1].cuda(), Y.cuda(), torch.tensor(LS))
    _check_factor(H1, O.infonce_ggn_naive(X[:1].numpy(), Y.numpy(), LS))
    # single target: softmax is one-hot -> zero curvature (up to rounding of a zero matrix)
    H0 = compute_hessian_analytic_InfoNCE(X.cuda(), Y[:1].cuda(), torch.tensor(LS))
    assert H0.abs().max().item() <= 1e-3 * H1.abs().max().item()
    with pytest.raises(RuntimeError):
        compute_hessian_analytic_InfoNCE(X, Y, torch.tensor(LS))  # CPU tensors: no fallback
    # strided rows
    Xs = torch.randn(3, 80, generator=gen).cuda()[:, :40]
    _check_factor(compute_hessian_analytic_InfoNCE(Xs, Y.cuda(), torch.tensor(LS)),
                  O.infonce_ggn_collapsed(Xs.cpu().numpy(), Y.numpy(), LS))


# ----------------------------------------------------------------------------------------------------------------------
# K1
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(64, 24, False), (64, 24, True), (1000, 769, False), (4096, 768, True), (32768, 512, False),
                                   (32768, 1280, False), (8192, 3072, True)])
def test_syrk_vs_fp64(shape):
    from bayesvlm_b200.hessians import syrk_accumulate

    n, d, one = shape
    gen = torch.Generator().manual_seed(n + d)
    X = torch.randn(n, d, generator=gen)
    A = syrk_accumulate(X.cuda(), append_one=one)
    Xa = torch.cat([X, torch.ones(n, 1)], 1) if one else X
    ref = (Xa.double().T @ Xa.double()).numpy()
    _check_factor(A, ref)
    assert torch.equal(A, A.T)
    # accumulate on top
    A2 = syrk_accumulate(X.cuda(), out=A.clone(), append_one=one, accumulate=True, alpha=0.5)
    _check_factor(A2, 1.5 * ref)


def test_syrk_accumulation_bias_is_bounded_for_long_row_blocks():
    """A rank's whole contiguous block goes through ONE launch (config 2: 2^20 rows).  The tensor cores' fp32 accumulator
    truncates, which biases sums of squares by ~ -5e-9 per row of an accumulation chain (scripts/syrk_bias.py: -5.5e-4 at 2^20
    rows before the chain length was bounded); the split-K schedule must keep it at the level of the operand rounding."""
    from bayesvlm_b200.hessians import syrk_accumulate

    n, d = 1 << 20, 64
    gen = torch.Generator(device="cuda").manual_seed(7)
    X = torch.randn(n, d, generator=gen, device="cuda")
    A = syrk_accumulate(X).double()
    ref = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    for lo in range(0, n, 1 << 18):
        xb = X[lo:lo + (1 << 18)].double()
        ref += xb.T @ xb
    assert float((A - ref).norm() / ref.norm()) <= 1e-4
    assert float(((A.diagonal() - ref.diagonal()) / ref.diagonal()).abs().max()) <= 1e-4


# ----------------------------------------------------------------------------------------------------------------------
# K0
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("likelihood", ["info_nce", "siglip"])
def test_kfac_ggn_golden(golden, likelihood):
    """The reference kfac_ggn loop (scripts/hessian_estimation.py:26-109) incl. its dropped remainders."""
    from bayesvlm_b200.hessians import kfac_ggn
    from bayesvlm_b200.vlm import CLIP, SIGLIP

    ncls, bs = (int(v) for v in golden["kfac_cfg"])
    if likelihood == "info_nce":
        vlm, tag = CLIP(logit_scale=LS, device="cuda"), "infonce"
    else:
        s, b = golden["ggn_siglip_params"]
        vlm, tag = SIGLIP(logit_scale=float(s), logit_bias=float(b), device="cuda"), "siglip"
    src_e, src_a, tgt_e = (torch.from_numpy(golden[k]) for k in ("kfac_emb_s", "kfac_act_s", "kfac_emb_t"))
    A, B = kfac_ggn(vlm, ncls, bs, src_e, src_a, tgt_e, "cuda", likelihood)
    assert B.device.type == "cpu" and A.device.type == "cuda"  # reference placement (:84,:97,:100)
    _check_factor(A, golden[f"kfac_{tag}_A"])
    _check_factor(B, golden[f"kfac_{tag}_B"])
    with pytest.raises(ValueError):
        kfac_ggn(vlm, 1000, bs, src_e, src_a, tgt_e, "cuda", likelihood)
    with pytest.raises(ValueError):
        kfac_ggn(vlm, ncls, bs, src_e, src_a, tgt_e, "cuda", "bogus")


def test_kfac_ggn_medium_vs_oracle():
    from bayesvlm_b200.hessians import kfac_ggn
    from bayesvlm_b200.vlm import CLIP

    gen = torch.Generator().manual_seed(2002)
    n, ncls, bs, D, d_in = 2 * 1024 + 100, 1024, 5, 512, 768
    src_e, tgt_e = _paired(gen, n, D)
    src_a = torch.randn(n, d_in, generator=gen)
    A, B = kfac_ggn(CLIP(logit_scale=LS, device="cuda"), ncls, bs, src_e, src_a, tgt_e, "cuda", "info_nce")
    Ar, Br = O.kfac_ggn(src_e.numpy(), src_a.numpy(), tgt_e.numpy(), ncls, bs, LS)
    _check_factor(A, Ar)
    _check_factor(B, Br)


def test_kfac_ggn_input_placement_is_irrelevant():
    """Pageable host, pinned host (staged one class batch ahead on a copy stream) and device-resident inputs run the same
    kernels on the same data: the same factors (up to the summation order of the split-K atomics)."""
    from bayesvlm_b200.hessians import kfac_ggn
    from bayesvlm_b200.vlm import CLIP

    gen = torch.Generator().manual_seed(2003)
    n, ncls, bs, D, d_in = 5 * 512 + 37, 512, 5, 128, 160
    src_e, tgt_e = _paired(gen, n, D)
    src_a = torch.randn(n, d_in, generator=gen)
    vlm = CLIP(logit_scale=LS, device="cuda")
    ref = kfac_ggn(vlm, ncls, bs, src_e.cuda(), src_a.cuda(), tgt_e.cuda(), "cuda", "info_nce")
    for place in (lambda t: t, lambda t: t.pin_memory()):
        A, B = kfac_ggn(vlm, ncls, bs, place(src_e), place(src_a), place(tgt_e), "cuda", "info_nce")
        # run-to-run differences of the split-K / column-sum atomics reach 3e-6 of the largest entry
        assert float((A - ref[0]).abs().max()) <= 2e-5 * float(ref[0].abs().max())
        assert float((B - ref[1]).abs().max()) <= 2e-5 * float(ref[1].abs().max())
