"""Support-set search (reference bayesvlm/knn.py): the oracle restatement against golden outputs of the reference's own
functions (tests/golden/make_golden.py::make_knn), and the product's host logic (selection of the support set from top-k
lists, index extraction, the generic distance expression) on CPU tensors."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import laplace_oracle as O

from conftest import GOLDEN


@pytest.fixture(scope="module")
def knn():
    return dict(np.load(GOLDEN / "knn_small.npz"))


def _golden_dict(g, tag):
    out = OrderedDict()
    for r, key in enumerate(g[f"{tag}_keys"]):
        n = int((g[f"{tag}_indices"][r] >= 0).sum())
        out[int(key)] = dict(score=float(g[f"{tag}_scores"][r]), indices=g[f"{tag}_indices"][r, :n].tolist(),
                             similarities=g[f"{tag}_sims"][r, :n].tolist())
    return out


def _same(res, ref, rtol):
    assert list(res.keys()) == list(ref.keys())
    for k in ref:
        assert res[k]["indices"] == ref[k]["indices"], k
        assert res[k]["score"] == pytest.approx(ref[k]["score"], rel=1e-6)
        np.testing.assert_allclose(res[k]["similarities"], ref[k]["similarities"], rtol=rtol, atol=rtol)


def _sim(g, tag, dtype=np.float32):
    it = g["indices_test"]
    fn = O.knn_neg_wasserstein if tag.startswith("wass") else O.knn_expected_cosine
    return fn(g["test_e"][it], g["test_a"][it], g["train_e"], g["train_a"], g["A_inv"], g["B_inv"], dtype)


@pytest.mark.parametrize("tag", ["cos", "wass", "cos_k5"])
def test_oracle_matches_reference(knn, tag):
    k_nearest, buf = (int(v) for v in knn[f"{tag}_cfg"])
    res = O.knn_support(_sim(knn, tag), knn["indices_test"], knn["values_test"], k_nearest, buf)
    _same(res, _golden_dict(knn, tag), rtol=2e-5)
    res64 = O.knn_support(_sim(knn, tag, np.float64), knn["indices_test"], knn["values_test"], k_nearest, buf)
    _same(res64, _golden_dict(knn, tag), rtol=2e-5)


def test_oracle_wasserstein_distance(knn):
    it = knn["indices_test"]
    w = O.knn_diagonal_wasserstein(knn["test_e"][it], knn["train_e"], knn["w_cov1"], knn["w_cov2"])
    np.testing.assert_allclose(w, knn["w_dist"], rtol=2e-5, atol=2e-4)


@pytest.mark.parametrize("tag", ["cos", "wass", "cos_k5"])
def test_product_selection_logic(knn, tag):
    """bayesvlm_b200.knn._support_from_topk (vectorised) == the reference's loops, on the oracle's similarity matrix."""
    from bayesvlm_b200.knn import _support_from_topk, extract_test_train_indices

    k_nearest, buf = (int(v) for v in knn[f"{tag}_cfg"])
    sim = torch.from_numpy(_sim(knn, tag))
    top = sim.topk(min(k_nearest + buf, sim.shape[1]), dim=1)
    res = _support_from_topk(top.indices, top.values, torch.from_numpy(knn["indices_test"]),
                             torch.from_numpy(knn["values_test"]), k_nearest)
    _same(res, _golden_dict(knn, tag), rtol=2e-5)
    ex = extract_test_train_indices(res)
    assert ex["test"] == knn[f"{tag}_keys"].tolist()
    assert sorted(ex["train"]) == knn[f"{tag}_extract_train"].tolist()


def test_product_selection_random_vs_oracle():
    from bayesvlm_b200.knn import _support_from_topk

    gen = torch.Generator().manual_seed(5)
    for n_test, n_train, k_nearest, buf in ((7, 60, 2, 6), (16, 40, 2, 30), (5, 500, 4, 3), (1, 9, 3, 2)):
        sim = torch.randn(n_test, n_train, generator=gen)
        sim[:, :3] += 2.5  # shared favourites: forces the neighbour count to grow
        ids, vals = torch.randperm(100, generator=gen)[:n_test], torch.rand(n_test, generator=gen)
        top = sim.topk(min(k_nearest + buf, n_train), dim=1)
        res = _support_from_topk(top.indices, top.values, ids, vals, k_nearest)
        _same(res, O.knn_support(sim.numpy(), ids.numpy(), vals.numpy(), k_nearest, buf), rtol=1e-6)


def test_product_selection_not_enough_neighbours():
    from bayesvlm_b200.knn import _support_from_topk

    sim = torch.zeros(4, 6)
    sim[:, :2] = 1.0
    top = sim.topk(2, dim=1)  # every row lists the same two samples: 2 distinct < 1 * 4
    with pytest.raises(ValueError):
        _support_from_topk(top.indices, top.values, torch.arange(4), torch.zeros(4), 1)


def test_product_distance_expression(knn):
    from bayesvlm_b200.knn import diagonal_wasserstein_distance, wdist2

    it = knn["indices_test"]
    args = [torch.from_numpy(a) for a in (knn["test_e"][it], knn["train_e"], knn["w_cov1"], knn["w_cov2"])]
    np.testing.assert_allclose(diagonal_wasserstein_distance(*args).numpy(), knn["w_dist"], rtol=2e-5, atol=2e-4)
    assert torch.equal(wdist2(*args), diagonal_wasserstein_distance(*args))


def test_product_refuses_cpu(knn):
    from bayesvlm_b200.hessians import KroneckerFactorizedCovariance
    from bayesvlm_b200.knn import find_similar_samples_cosine
    from bayesvlm_b200.vlm import EncoderResult

    t = lambda k: torch.from_numpy(knn[k])
    cov = KroneckerFactorizedCovariance(A_inv=t("A_inv"), B_inv=t("B_inv"))
    with pytest.raises(RuntimeError):
        find_similar_samples_cosine(EncoderResult(t("train_e"), t("train_a")), EncoderResult(t("test_e"), t("test_a")),
                                    t("indices_test"), t("values_test"), 3, cov, device="cpu", buffersize=10)


def test_hostmem_degrades_without_gpu():
    """bayesvlm_b200.hostmem is best effort: without a CUDA device / NVML the placement helpers are no-ops."""
    from bayesvlm_b200.hostmem import gpu_local_cpus, numa_local

    if not torch.cuda.is_available():
        assert gpu_local_cpus("cuda:0") is None
    with numa_local("cpu") as bound:
        assert bound is False
