"""bayesvlm_b200.selection on CUDA tensors (the scores are row reductions of the predictive's [N, C] outputs, which live on
the device): deterministic variants against the golden outputs of the reference's selection.py; Monte-Carlo variants are
reproducible per device and statistically consistent with the CPU golden (different generator stream on CUDA)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_selection_on_device():
    from bayesvlm_b200 import selection as S
    from bayesvlm_b200.vlm import ProbabilisticLogits

    g = dict(np.load(GOLDEN / "selection_small.npz"))
    pl = ProbabilisticLogits(mean=torch.from_numpy(g["mean"]).cuda(), var=torch.from_numpy(g["var"]).cuda())
    ids = torch.from_numpy(g["class_ids"]).cuda()
    for ev in ("map_alea", "comb"):
        np.testing.assert_allclose(S._entropy(pl.mean, pl.var, ev).cpu().numpy(), g[f"entropy_{ev}"], rtol=2e-5, atol=1e-6)
    idx, val = S.select_topk(pl, 9, "entropy", "map_alea", ignore_percentage=0.1, return_values=True)
    assert idx.cpu().tolist() == g["topk_entropy_idx"].tolist()
    assert S.select_topk_classbalanced(pl, ids, 10, "entropy", "map_alea").cpu().tolist() == g["topk_cb_entropy"].tolist()
    assert S.select_topk_classbalanced(pl, ids, 10, "var").cpu().tolist() == g["topk_cb_var"].tolist()
    a = S._entropy(pl.mean, pl.var, "comb_covar", num_samples=400, seed=3)
    b = S._entropy(pl.mean, pl.var, "comb_covar", num_samples=400, seed=3)
    assert torch.equal(a, b)
    assert np.abs(a.cpu().numpy() - g["entropy_comb_covar"]).mean() < 0.1  # 25-sample CPU estimate vs 400-sample CUDA one
    assert S.select_random(pl, 12, seed=8).tolist() == g["random"].tolist()  # host permutation: device independent
