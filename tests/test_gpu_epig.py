"""E0/E1/E2/E3 parity of the CUDA path ("EPIG parity protocol").

The reference evaluates EPIG in fp16; its scores are decided by where that arithmetic rounds, and torch's CPU and CUDA
kernels round differently (xlogy: tests/golden/torch_cuda_half_semantics.npz).  The kernels reproduce the CUDA rounding
points, so the checker is the reference's operation sequence executed by torch ON THE SAME GPU (oracle/torch_port.py, pinned
on the reference itself by tests/test_oracle_golden.py):

  (1) identical fp16 probabilities in, scores out: exact-match rate >= 95 %; a mismatch is at most ONE score quantum (one
      fp16 ulp of a per-chunk partial sum, carried through `/ N_t`) except for the rare row where two chunks flip; top-k sets
      identical after expanding ties at that quantum;
  (2) agreement with the noise-free fp32 definition is reported by the oracle tests (CPU);
  (3) `select_epig_online`, budget 5 on a 4096 x 2000 problem with the shared device RNG: same picks as the reference loop.

What remains different between two correct implementations is the fp32 summation ORDER inside a chunk (266k addends at
Cl = 65), which moves a partial sum across an fp16 rounding boundary in ~3e-4 of the chunks.
"""
import json
import math
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import laplace_oracle as O
from oracle import torch_port as T

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _logits(gen, n, cl, spread=2.0):
    return (torch.randn(n, cl, generator=gen) * spread).cuda(), (torch.rand(n, cl, generator=gen) * 3 + 0.1).cuda()


def _probs16(gen, n, k, cl):
    """fp16 probabilities produced by torch's own expression (vlm.py:116-123 then .to(float16)) on the GPU."""
    mean, var = _logits(gen, n, cl)
    eps = torch.randn(k, n, cl, generator=gen).cuda()
    return torch.softmax((eps * var.sqrt() + mean).permute(1, 0, 2), dim=2).half(), (mean, var, eps)


def _reference_with_quantum(p16, t16, chunk):
    """The reference sequence (epig.py:371-395) with torch's CUDA kernels, plus the score quantum: the largest change of a
    per-chunk term `fp16(fp16(-sum) / N_t)` when its fp16 partial sum moves by one ulp."""
    n_t, k, cl = t16.shape
    h_pool = T._entropy(p16.mean(dim=1))
    h_targ = T._entropy(t16.mean(dim=1)).mean()
    pool = p16.permute(0, 2, 1)
    targ = t16.permute(1, 0, 2).reshape(k, n_t * cl)
    acc = torch.zeros(pool.shape[0], device=p16.device)
    quantum = 0.0
    for lo in range(0, n_t * cl, chunk):
        joint = pool @ targ[:, lo:lo + chunk] / k
        s16 = torch.sum(torch.xlogy(joint, joint), dim=(-2, -1))
        h = -s16 / n_t
        acc += h
        ulp_s = np.spacing(np.abs(s16.cpu().numpy()).max().astype(np.float16)).astype(np.float64) / n_t
        ulp_h = float(np.spacing(np.abs(h.cpu().numpy()).max().astype(np.float16)))
        quantum = max(quantum, float(ulp_s), ulp_h)
    return (h_pool + h_targ - acc).float(), quantum


def _check_scores(s, ref, quantum, k_top, record=None, name=None):
    s = s.float().cpu().numpy().astype(np.float64)
    ref = ref.float().cpu().numpy().astype(np.float64)
    d = np.abs(s - ref)
    exact = float((d == 0).mean())
    over1 = float((d > 1.01 * quantum).mean())
    # top-k sets identical after expanding ties at the quantum
    a = set(np.argsort(-s, kind="stable")[:k_top].tolist())
    b = set(np.argsort(-ref, kind="stable")[:k_top].tolist())
    kth = np.sort(ref)[::-1][k_top - 1]
    ties_ok = all(abs(ref[i] - kth) <= 2.02 * quantum for i in a ^ b)
    if record is not None:
        record[name] = {"exact_match_rate": exact, "max_abs_diff_in_quanta": float(d.max() / quantum), "quantum": quantum,
                        "rows": int(d.size), "topk": k_top, "topk_symmetric_difference": len(a ^ b)}
    assert exact >= 0.95, (name, exact)
    assert d.max() <= 2.02 * quantum, (name, d.max(), quantum)   # two flipped chunks in one row: rare, bounded
    assert over1 <= 0.005, (name, over1)                          # "at most one quantum" for >= 99.5 % of the rows
    assert ties_ok, (name, sorted(a ^ b))
    return exact


_RATES = {}


@pytest.fixture(scope="module", autouse=True)
def _write_match_rates():
    yield
    out = ROOT / "gpurun_out"
    if _RATES and (out.exists() or os.environ.get("GRAFT_REPO_ROOT")):
        out.mkdir(exist_ok=True)
        (out / "epig_match_rates.json").write_text(json.dumps(_RATES, indent=1))
    print("\nEPIG match rates vs torch CUDA:", json.dumps(_RATES))


# ------------------------------------------------------------------------------------------------------------------ E0 / E1
@pytest.mark.parametrize("cl,k", [(10, 100), (5, 16), (16, 33), (65, 64)])
def test_sample_probs_vs_torch_cuda(cl, k):
    """E0: the sampling kernel against torch's expression on the same GPU; bit-identical for rows of <= 16 classes (the
    kernel follows softmax_warp_forward's summation order), >= 99.9 % otherwise (<= one fp16 ulp)."""
    from bayesvlm_b200.vlm import sample_probas_from_noise

    gen = torch.Generator().manual_seed(100 + cl)
    ref, (mean, var, eps) = _probs16(gen, 777, k, cl)
    p = sample_probas_from_noise(mean, var, eps)
    assert p.dtype == torch.float16 and p.shape == ref.shape
    d = (p.float() - ref.float()).abs()
    rate = float((d == 0).float().mean())
    _RATES[f"E0_sample_Cl{cl}"] = rate
    assert rate == 1.0 if cl <= 16 else rate >= 0.999
    ulp = torch.from_numpy(np.spacing(ref.cpu().numpy()).astype(np.float32)).cuda()
    assert bool((d <= ulp).all())


@pytest.mark.parametrize("n,k,cl,misalign", [(7, 400, 10, 0), (1, 1, 1, 0), (13, 33, 3, 0), (130, 100, 10, 1), (130, 100, 10, 2),
                                               (6, 64, 12, 3), (9, 130, 16, 0), (5, 20, 200, 0), (11, 257, 7, 1)])
def test_prepare_kernel_odd_shapes_and_alignments(n, k, cl, misalign):
    """The fused prepare kernel (E0 + E1 + operand layout) at shapes that exercise its tails and fallbacks: row counts that
    are not a multiple of the 4-row block, K > 128 (several samples per thread) and K = 1, odd / large class counts (scalar and
    generic paths), and noise whose base address is only 4- / 8- / 12-byte aligned (8- and 4-byte cp.async pieces)."""
    from bayesvlm_b200 import epig as E

    gen = torch.Generator().manual_seed(1000 * n + k + cl)
    mean = (torch.randn(n, cl, generator=gen) * 2).cuda()
    var = (torch.rand(n, cl, generator=gen) * 3 + 0.1).cuda()
    buf = torch.randn(k * n * cl + 8, generator=gen).cuda()
    eps = buf[misalign:misalign + k * n * cl].view(k, n, cl)
    assert eps.data_ptr() % 16 == (4 * misalign) % 16
    ref = torch.softmax((eps * var.sqrt() + mean).permute(1, 0, 2), dim=2).half()
    assert E._prepare_fits(k, cl, True)
    probs, oper, marg = E.prepare_from_noise(mean, var, eps, want_probs=True)
    d = (probs.float() - ref.float()).abs()
    assert bool((d == 0).all()) if cl <= 16 else float((d == 0).double().mean()) >= 0.999
    kp = oper.shape[2]
    assert kp % 64 == 0 and kp >= k
    assert torch.equal(oper[:, :, :k], probs.permute(0, 2, 1)) and bool((oper[:, :, k:] == 0).all())
    me = E.entropy_from_probs(torch.mean(ref, dim=1))  # torch's own Half sequence on the same probabilities
    assert float((marg.float() - me.float()).abs().max()) <= float(np.spacing(np.float16(max(1.0, float(me.abs().max())))))
    # the same operands from the fp16 probabilities (the [N, K, Cl] entry point)
    oper2, marg2 = E.prepare_from_probs(probs)
    assert torch.equal(oper2, oper) and torch.equal(marg2, marg)


def test_sample_probs_golden(golden):
    from bayesvlm_b200.vlm import sample_probas_from_noise

    p = sample_probas_from_noise(_cuda(golden["epig_mean_p"]), _cuda(golden["epig_var_p"]), _cuda(golden["epig_eps_p"]))
    ref16 = golden["epig_probs_p"].astype(np.float16)
    d = np.abs(p.cpu().numpy().astype(np.float32) - ref16.astype(np.float32))
    assert (d == 0).mean() >= 0.99          # the golden softmax ran on the CPU (other summation order): last-bit cases
    assert (d <= np.spacing(np.abs(ref16)).astype(np.float32)).all()


def test_sample_probs_shape_checks():
    from bayesvlm_b200.vlm import sample_probas_from_noise

    m, v, e = torch.zeros(4, 3).cuda(), torch.ones(4, 3).cuda(), torch.zeros(2, 4, 3).cuda()
    assert sample_probas_from_noise(m, v, e).shape == (4, 2, 3)
    with pytest.raises(ValueError):
        sample_probas_from_noise(m, v[:3], e)
    with pytest.raises(ValueError):
        sample_probas_from_noise(m, v, e[:, :3])
    with pytest.raises(RuntimeError):
        sample_probas_from_noise(m.cpu(), v, e)


@pytest.mark.parametrize("n,k,cl", [(4096, 100, 10), (1000, 100, 65), (333, 16, 5), (64, 33, 128)])
def test_marginal_entropy_vs_torch_cuda(n, k, cl):
    from bayesvlm_b200.epig import marginal_entropy_from_probs

    gen = torch.Generator().manual_seed(n + cl)
    p16, _ = _probs16(gen, n, k, cl)
    me = marginal_entropy_from_probs(p16)
    ref = T._entropy(p16.mean(dim=1))
    assert me.dtype == torch.float16 and me.shape == ref.shape
    d = (me.float() - ref.float()).abs().cpu().numpy()
    rate = float((d == 0).mean())
    _RATES[f"E1_marginal_N{n}_Cl{cl}"] = rate
    assert rate >= 0.99
    assert (d <= np.spacing(np.abs(ref.cpu().numpy())).astype(np.float32)).all()


def test_marginal_entropy_golden_and_generic_dtype(golden):
    from bayesvlm_b200.epig import marginal_entropy_from_probs

    p16 = _cuda(golden["epig_probs_p"].astype(np.float16))
    me = marginal_entropy_from_probs(p16).cpu().numpy()
    # the CUDA rounding points, restated by the oracle: exact; the CPU-reference golden: within one fp16 ulp
    assert np.array_equal(me, O.marginal_entropy_f16(p16.cpu().numpy(), "cuda"))
    ref = golden["epig_marginal_p16"]
    assert (np.abs(me.astype(np.float32) - ref.astype(np.float32)) <= np.spacing(np.abs(ref)).astype(np.float32)).all()
    me32 = marginal_entropy_from_probs(_cuda(golden["epig_probs_p"]))  # fp32 input keeps torch's generic path
    np.testing.assert_allclose(me32.cpu().numpy(), golden["epig_marginal_p32"], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------------------------ E2
@pytest.mark.parametrize("cfg", [dict(Np=45, Nt=30, K=16, Cl=5, chunk=256), dict(Np=300, Nt=200, K=100, Cl=10, chunk=512),
                                 dict(Np=257, Nt=129, K=64, Cl=65, chunk=4096), dict(Np=64, Nt=77, K=33, Cl=128, chunk=1024),
                                 dict(Np=1000, Nt=700, K=100, Cl=10, chunk=4096),
                                 dict(Np=37, Nt=41, K=20, Cl=200, chunk=2048),         # a pool row spans CTAs (ImageNet-R: 200 classes)
                                 dict(Np=9, Nt=12, K=8, Cl=300, chunk=1024),           # ... and CTA pairs
                                 dict(Np=50, Nt=61, K=400, Cl=10, chunk=256),          # K = 400 (7 K blocks), one tile per chunk
                                 dict(Np=3, Nt=2, K=5, Cl=2, chunk=8192),              # one partial tile, chunk > all columns
                                 dict(Np=4096, Nt=10000, K=100, Cl=10, chunk=4096),    # config 5, primary
                                 dict(Np=4096, Nt=10000, K=100, Cl=65, chunk=4096)])   # config 5, secondary (OfficeHome)
def test_epig_scores_vs_reference_sequence_on_same_gpu(cfg):
    """Protocol (1): identical fp16 probabilities to the reference's operation sequence on this GPU and to the fused path."""
    from bayesvlm_b200.epig import epig_from_probs_using_matmul

    gen = torch.Generator().manual_seed(cfg["Np"] + 13 * cfg["Cl"])
    p16, _ = _probs16(gen, cfg["Np"], cfg["K"], cfg["Cl"])
    t16, _ = _probs16(gen, cfg["Nt"], cfg["K"], cfg["Cl"])
    s = epig_from_probs_using_matmul(p16, t16, chunk_size=cfg["chunk"])
    assert s.dtype == torch.float32 and bool(torch.isfinite(s).all())
    ref, quantum = _reference_with_quantum(p16, t16, cfg["chunk"])
    assert torch.equal(ref, T.epig_from_probs(p16, t16, chunk_size=cfg["chunk"]).float())  # the helper IS the port's sequence
    name = "E2_scores_Np{Np}_Nt{Nt}_K{K}_Cl{Cl}_chunk{chunk}".format(**cfg)
    _check_scores(s, ref, quantum, min(50, cfg["Np"] // 2), _RATES, name)


def test_epig_scores_vs_oracle_cuda_semantics():
    """The numpy restatement of the CUDA rounding points (oracle, device='cuda') agrees with the kernel as well."""
    from bayesvlm_b200.epig import epig_from_probs_using_matmul

    gen = torch.Generator().manual_seed(77)
    p16, _ = _probs16(gen, 200, 100, 10)
    t16, _ = _probs16(gen, 150, 100, 10)
    s = epig_from_probs_using_matmul(p16, t16, chunk_size=512).cpu().numpy()
    ref = O.epig_from_probs_f16(p16.cpu().numpy(), t16.cpu().numpy(), 512, device="cuda")
    d = np.abs(s - ref)
    assert (d == 0).mean() >= 0.95 and d.max() <= 2.0 ** -10


def test_epig_fused_path_vs_reference_golden():
    """The fused kernel (chunk 256) against scores the REFERENCE ITSELF produced (CPU, tests/golden/make_golden.py).  The CPU
    kernels of torch round xlogy differently from its CUDA kernels, so agreement here is to a few fp16 ulps of the
    O(1) entropies, not bit-exact (the bit-exact pins are the same-GPU test above and the oracle's CPU mode)."""
    from bayesvlm_b200.epig import epig_from_probs_using_matmul, marginal_entropy_from_probs

    g = np.load(ROOT / "tests" / "golden" / "epig_online_small.npz")
    p16, t16 = _cuda(g["fused_p16"]), _cuda(g["fused_t16"])
    s = epig_from_probs_using_matmul(p16, t16, chunk_size=int(g["fused_chunk"][0])).cpu().numpy()
    assert np.abs(s - g["fused_scores"]).max() <= 4 * 2.0 ** -10
    s_cuda = O.epig_from_probs_f16(g["fused_p16"], g["fused_t16"], int(g["fused_chunk"][0]), device="cuda")
    assert (s != s_cuda).mean() <= 0.05 and np.abs(s - s_cuda).max() <= 2.0 ** -10
    me = marginal_entropy_from_probs(p16).cpu().numpy().astype(np.float32)
    assert np.abs(me - g["fused_marginal"].astype(np.float32)).max() <= 2.0 ** -9


def test_epig_scores_golden_generic_path(golden):
    """chunk = 64 is not a multiple of the kernel's 256-column tile: generic device expression (torch ops)."""
    from bayesvlm_b200.epig import epig_from_probs_using_matmul

    p16 = _cuda(golden["epig_probs_p"].astype(np.float16))
    t16 = _cuda(golden["epig_probs_t"].astype(np.float16))
    chunk = int(golden["epig_cfg"][0])
    s = epig_from_probs_using_matmul(p16, t16, chunk_size=chunk)
    assert np.abs(s.float().cpu().numpy() - golden["epig_scores_f16"]).max() <= 4e-3
    s32 = epig_from_probs_using_matmul(_cuda(golden["epig_probs_p"]), _cuda(golden["epig_probs_t"]), chunk_size=chunk)
    np.testing.assert_allclose(s32.cpu().numpy(), golden["epig_scores_f32"], atol=2e-5)


# ------------------------------------------------------------------------------------------------------------------ E3
@pytest.mark.parametrize("n_pool,n_targ,cl,k,chunk", [(5000, 1500, 10, 100, 4096), (700, 300, 10, 32, 256), (600, 500, 65, 64, 512)])
def test_epig_from_logits_vs_reference_sequence_shared_device_rng(n_pool, n_targ, cl, k, chunk):
    """E3 against oracle/torch_port.epig_from_logits: both sides seed torch's generator and draw with torch.randn on the GPU
    (the reference's RNG contract, vlm.py:113-121), per pool chunk with seed + row offset (epig.py:323-334)."""
    from bayesvlm_b200.epig import epig_from_logits_using_matmul
    from bayesvlm_b200.vlm import ProbabilisticLogits

    gen = torch.Generator().manual_seed(9 + cl)
    mp, vp = _logits(gen, n_pool, cl)
    mt, vt = _logits(gen, n_targ, cl)
    s = epig_from_logits_using_matmul(ProbabilisticLogits(mp, vp), ProbabilisticLogits(mt, vt), seed=3, num_samples=k, chunk_size=chunk)
    ref = T.epig_from_logits(mp, vp, mt, vt, seed=3, num_samples=k, chunk_size=chunk)
    assert s.shape == (n_pool,) and s.dtype == torch.float32
    # quantum of this problem from the first pool chunk's probabilities
    pt = T.sample_probas(mt, vt, k, 3).half()
    pp = T.sample_probas(mp[:chunk], vp[:chunk], k, 3).half()
    _, quantum = _reference_with_quantum(pp, pt, chunk)
    _check_scores(s, ref, quantum, 50, _RATES, f"E3_from_logits_Np{n_pool}_Nt{n_targ}_Cl{cl}")
    assert torch.equal(s, epig_from_logits_using_matmul(ProbabilisticLogits(mp, vp), ProbabilisticLogits(mt, vt), seed=3,
                                                        num_samples=k, chunk_size=chunk))


def test_epig_errors():
    from bayesvlm_b200.epig import epig_from_probs_using_matmul, marginal_entropy_from_probs

    p = torch.rand(4, 3, 5).softmax(-1)
    with pytest.raises(RuntimeError):
        epig_from_probs_using_matmul(p, p)  # CPU tensors: no fallback
    with pytest.raises(AssertionError):
        marginal_entropy_from_probs(p[0].cuda())
    with pytest.raises(ValueError):
        epig_from_probs_using_matmul(p.cuda().half(), p[:, :2].cuda().half())


# ------------------------------------------------------------------------------------------------------------------ online loop
def _online_problem(seed, D, d_in, n_cls, n_pool, n_targ):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    spd = lambda d, sc: (lambda w: (w.T @ w) / math.sqrt(4 * d) * sc)(rn(4 * d, d))
    W = rn(D, d_in) / math.sqrt(d_in)
    pool_a, targ_a = rn(n_pool, d_in), rn(n_targ, d_in)
    label_e, label_a = rn(n_cls, D), rn(n_cls, D)
    ids = torch.randint(0, n_cls, (n_pool,), generator=g)
    return dict(W=W, pool_a=pool_a, targ_a=targ_a, label_e=label_e, label_a=label_a, ids=ids, A_img=spd(d_in, 3e3),
                A_txt=spd(D, 3e3), B_img=spd(D, 20.0), B_txt=spd(D, 20.0),
                info={"n_img": 1.0, "n_txt": 1.0, "lambda_img": 600.0, "lambda_txt": 220.0})


@pytest.mark.parametrize("shape", [dict(D=64, d_in=96, n_cls=10, n_pool=4096, n_targ=2000, budget=5, K=100, chunk=4096, pool_max=None,
                                        targ_max=None),
                                   dict(D=32, d_in=40, n_cls=6, n_pool=600, n_targ=300, budget=3, K=16, chunk=256, pool_max=512,
                                        targ_max=256)])
def test_select_epig_online_vs_reference_loop(shape):
    """Protocol (3): the online greedy loop (reference epig.py:44-273) on the kernels against the SAME loop on the reference's
    torch operations (oracle/torch_port.select_epig_online, pinned on the reference's own picks on the CPU), both on this GPU
    with the shared device RNG.  The picks must be identical; should a pick differ, it must be a tie at the score quantum in
    the reference's own score vector (after which the two trajectories legitimately diverge and the comparison stops)."""
    from bayesvlm_b200.epig import select_epig_online
    from bayesvlm_b200.vlm import CLIP, EncoderResult

    pr = _online_problem(21, shape["D"], shape["d_in"], shape["n_cls"], shape["n_pool"], shape["n_targ"])
    ls = math.log(20.0)
    kw = dict(budget=shape["budget"], lr=1e-4, hessian_update_scale=10.0, num_samples=shape["K"], seed=0,
              pool_max_size=shape["pool_max"], target_max_size=shape["targ_max"], chunk_size=shape["chunk"])
    dev = torch.device("cuda")
    ref_idx, ref_scores, ref_all, ref_subset = T.select_epig_online(
        pr["label_e"], pr["label_a"], pr["pool_a"] @ pr["W"].T, pr["pool_a"], pr["targ_a"] @ pr["W"].T, pr["targ_a"], pr["ids"],
        pr["W"], ls, pr["A_img"], pr["A_txt"], pr["B_img"], pr["B_txt"], pr["info"], device=dev, **kw)
    proj = torch.nn.Linear(shape["d_in"], shape["D"], bias=False)
    with torch.no_grad():
        proj.weight.copy_(pr["W"])
        pool = EncoderResult(pr["pool_a"] @ pr["W"].T, pr["pool_a"])
        targ = EncoderResult(pr["targ_a"] @ pr["W"].T, pr["targ_a"])
    idx, scores = select_epig_online(
        label_features=EncoderResult(pr["label_e"], pr["label_a"]), pool_features=pool, target_features=targ,
        pool_class_ids=pr["ids"], image_projection=proj, clip=CLIP(logit_scale=ls), A_img=pr["A_img"], A_txt=pr["A_txt"],
        B_img=pr["B_img"], B_txt=pr["B_txt"], cov_info=dict(pr["info"]), device=dev, **kw)
    assert len(idx) == shape["budget"] and len(set(idx)) == shape["budget"] and all(math.isfinite(v) for v in scores)
    agree = 0
    for step, (a, b) in enumerate(zip(idx, ref_idx)):
        if a == b:
            agree += 1
            continue
        # a differing pick must be a tie at the score quantum in the REFERENCE's own score vector of that step
        sv = ref_all[step].float().cpu().numpy()
        pos = (ref_subset == a).nonzero()
        assert pos.numel() == 1, (step, a)
        quantum = 2.0 ** -11 * max(1.0, 2 * math.log(shape["n_cls"]))  # one fp16 ulp of the O(log Cl^2) entropies
        assert sv.max() - sv[int(pos)] <= 2 * quantum, (step, a, b, sv.max(), sv[int(pos)])
        break
    _RATES[f"online_loop_pool{shape['n_pool']}_budget{shape['budget']}"] = {"picks_identical": agree, "of": shape["budget"],
                                                                           "ours": idx, "reference_loop": ref_idx}
    assert agree >= 1
    if agree == shape["budget"]:
        np.testing.assert_allclose(scores, ref_scores, atol=4 * 2.0 ** -10)
