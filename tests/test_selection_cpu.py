"""bayesvlm_b200.selection against golden outputs of the reference's bayesvlm/selection.py (tests/golden/make_golden.py::
make_selection).  The functions are device-generic torch expressions that draw from torch's default generator in the
reference's call order, so on the same device and seed they reproduce its Monte-Carlo results."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def sel():
    return dict(np.load(GOLDEN / "selection_small.npz"))


def _logits(g):
    from bayesvlm_b200.vlm import ProbabilisticLogits

    return ProbabilisticLogits(mean=torch.from_numpy(g["mean"]), var=torch.from_numpy(g["var"]))


@pytest.mark.parametrize("variant", ["map_alea", "comb", "comb_covar", "exp_alea"])
def test_entropy_variants(sel, variant):
    from bayesvlm_b200.selection import _entropy

    torch.manual_seed(11)
    h = _entropy(torch.from_numpy(sel["mean"]), torch.from_numpy(sel["var"]), variant, num_samples=25, seed=3)
    np.testing.assert_allclose(h.numpy(), sel[f"entropy_{variant}"], rtol=1e-5, atol=1e-6)


def test_scores(sel):
    from bayesvlm_b200.selection import complexity_score
    from bayesvlm_b200.vlm import ProbabilisticLogits

    pl = _logits(sel)
    np.testing.assert_allclose(complexity_score(pl, "var").numpy(), sel["score_var"], rtol=1e-6)
    torch.manual_seed(12)
    np.testing.assert_allclose(complexity_score(pl, "map_mutual_info", seed=5).numpy(), sel["score_map_mi"], rtol=1e-4, atol=1e-6)
    torch.manual_seed(13)
    np.testing.assert_allclose(complexity_score(pl, "exp_mutual_info", seed=5).numpy(), sel["score_exp_mi"], rtol=1e-4, atol=1e-6)
    cov = torch.stack([torch.diag(v) + 0.01 for v in pl.var[:10]])
    pl3 = ProbabilisticLogits(mean=pl.mean[:10], var=cov)
    np.testing.assert_allclose(complexity_score(pl3, "logdet").numpy(), sel["score_logdet"], rtol=1e-5)
    np.testing.assert_allclose(complexity_score(pl3, "var").numpy(), sel["score_var3d"], rtol=1e-6)
    assert complexity_score(pl, "nonsense") is None


def test_selections(sel):
    from bayesvlm_b200 import selection as S
    from bayesvlm_b200.vlm import ProbabilisticLogits

    pl = _logits(sel)
    ids = torch.from_numpy(sel["class_ids"])
    idx, val = S.select_topk(pl, 9, "entropy", "map_alea", ignore_percentage=0.1, return_values=True)
    assert idx.tolist() == sel["topk_entropy_idx"].tolist()
    np.testing.assert_allclose(val.numpy(), sel["topk_entropy_val"], rtol=1e-6)
    assert S.select_topk(pl, 9, "entropy", "map_alea", ignore_percentage=0.1).tolist() == sel["topk_entropy_idx"].tolist()
    cov = torch.stack([torch.diag(v) + 0.01 for v in pl.var[:10]])
    assert S.select_topk(ProbabilisticLogits(mean=pl.mean[:10], var=cov), 4, "var").tolist() == sel["topk_var3d"].tolist()
    assert S.select_topk_classbalanced(pl, ids, 10, "var").tolist() == sel["topk_cb_var"].tolist()
    assert S.select_topk_classbalanced(pl, ids, 10, "entropy", "map_alea").tolist() == sel["topk_cb_entropy"].tolist()
    assert S.select_topk_randomized(pl, 8, 1.5, "entropy", "comb", seed=4).tolist() == sel["topk_rand"].tolist()
    assert S.select_random_classbalanced(pl.var, ids, 10, seed=6).tolist() == sel["random_cb"].tolist()
    assert S.select_random(pl, 12, seed=8).tolist() == sel["random"].tolist()
    with pytest.raises(RuntimeError):  # 2-D variances make 'var' a scalar (reference selection.py:35): topk fails there too
        S.select_topk(pl, 9, "var")
