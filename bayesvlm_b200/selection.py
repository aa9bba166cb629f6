"""Acquisition scores and subset selection over predictive logits (reference ``bayesvlm/selection.py``; SURVEY.md
section 8(f) row 4).  Device-generic torch expressions over :class:`ProbabilisticLogits`: the scores are row reductions of
``[N, C]`` tensors that already live on the GPU after the predictive kernels; the Monte-Carlo variants draw from torch's
default generator exactly like the reference (same call order), so a shared seed reproduces its draws on the same device.

Not mirrored: ``create_subset_json`` (reference selection.py:147-173) passes ``entropy_variant='alea'``, which none of its
own branches accepts, and fails before returning.
"""
from typing import Literal, Optional

import torch

from .vlm import ProbabilisticLogits

EntropyVariant = Literal["map_alea", "exp_alea", "comb", "comb_covar"]
ScoreVariant = Literal["var", "logdet", "entropy", "map_mutual_info", "exp_mutual_info"]


def _shannon(probas: torch.Tensor) -> torch.Tensor:
    return -(probas * probas.log()).sum(dim=1)


def _entropy(logits_mean: torch.Tensor, logits_var: torch.Tensor, variant: EntropyVariant, num_samples: int = 1000,
             seed: Optional[int] = None) -> torch.Tensor:
    """Predictive entropy per row under one of four treatments of the logit uncertainty (reference selection.py:7-26):
    MAP softmax, expected aleatoric entropy (MC), probit-adjusted softmax (the method's 2-D quirk included), MC softmax."""
    logits = ProbabilisticLogits(mean=logits_mean, var=logits_var)
    if variant == "exp_alea":
        return logits.expected_aleatoric_entropy(num_samples=num_samples)  # (the reference does not seed this one)
    if variant == "map_alea":
        return _shannon(torch.softmax(logits.mean, dim=1))
    if variant == "comb":
        return _shannon(logits.softmax(num_samples=0, seed=seed))
    if variant == "comb_covar":
        return _shannon(logits.softmax(num_samples=num_samples, seed=seed))
    raise UnboundLocalError(f"unknown entropy variant {variant!r}")  # the reference fails the same way (:26)


def complexity_score(prob_logits: ProbabilisticLogits, variant: ScoreVariant, entropy_variant: Optional[EntropyVariant] = None,
                     seed: Optional[int] = None) -> torch.Tensor:
    """Per-sample acquisition score (reference selection.py:28-50); ``None`` for an unknown variant, like the reference."""
    if variant == "var":
        return prob_logits.var.diagonal(dim1=-2, dim2=-1).sum(dim=-1)
    if variant == "logdet":
        return prob_logits.var.logdet()
    if variant == "entropy":
        return _entropy(prob_logits.mean, prob_logits.var, entropy_variant, seed=seed)
    if variant in ("exp_mutual_info", "map_mutual_info"):
        total = _entropy(prob_logits.mean, prob_logits.var, "comb_covar", seed=seed)
        aleatoric = _entropy(prob_logits.mean, prob_logits.var, "exp_alea" if variant == "exp_mutual_info" else "map_alea",
                             seed=seed)
        return total - aleatoric
    return None


def select_topk(prob_logits: ProbabilisticLogits, k: int, variant: ScoreVariant,
                entropy_variant: Optional[EntropyVariant] = None, ignore_percentage: float = 0.0, return_values: bool = False,
                seed: Optional[int] = None):
    """Indices of the k highest scores after skipping the top ``ignore_percentage`` fraction (reference selection.py:52-76)."""
    n_rows = prob_logits.mean.shape[0]
    offset = int(n_rows * ignore_percentage) if ignore_percentage > 0.0 else 0
    top = complexity_score(prob_logits, variant, entropy_variant, seed=seed).topk(min(k + offset, n_rows))
    if return_values:
        return top.indices[offset:], top.values[offset:]
    return top.indices[offset:]


def _per_class_quota(class_ids: torch.Tensor, k: int):
    classes = class_ids.unique(sorted=True)
    base, extra = divmod(k, len(classes))
    return [(c, base + (1 if i < extra else 0)) for i, c in enumerate(classes)]


def select_topk_classbalanced(prob_logits: ProbabilisticLogits, class_ids: torch.Tensor, k: int,
                              variant: Literal["var", "entropy"], entropy_variant=None) -> torch.Tensor:
    """k // #classes picks per class, the first k % #classes classes get one more (reference selection.py:78-104).
    As in the reference, the returned indices are positions INSIDE each class's subset, concatenated."""
    picks = []
    for c, n in _per_class_quota(class_ids, k):
        mask = class_ids == c
        if variant == "var":
            picks.append(prob_logits.var[mask].sum(dim=1).topk(n).indices)
        elif variant == "entropy":
            picks.append(_entropy(prob_logits.mean[mask], prob_logits.var[mask], entropy_variant).topk(n).indices)
    return torch.cat(picks)


def select_topk_randomized(prob_logits: ProbabilisticLogits, k: int, temp: float, variant: ScoreVariant,
                           entropy_variant: Optional[EntropyVariant] = None, seed: int = 0) -> torch.Tensor:
    """k draws (with replacement) from softmax(temp * standardised score) (reference selection.py:106-121)."""
    score = complexity_score(prob_logits, variant, entropy_variant)
    torch.manual_seed(seed)
    score = (score - score.mean()) / score.std()
    return torch.distributions.Categorical(probs=torch.softmax(score * temp, dim=0)).sample((k,))


def select_random_classbalanced(logits_var: torch.Tensor, class_ids: torch.Tensor, k: int, seed: int) -> torch.Tensor:
    """Uniformly random class-balanced subset (reference selection.py:125-142); `logits_var` is unused there too."""
    del logits_var
    torch.manual_seed(seed)
    picks = []
    for c, n in _per_class_quota(class_ids, k):
        members = torch.where(class_ids == c)[0]
        picks.append(members[torch.randperm(len(members))[:n]])
    return torch.cat(picks)


def select_random(prob_logits: ProbabilisticLogits, k: int, seed: Optional[int]) -> torch.Tensor:
    """First k entries of a seeded permutation of the rows (reference selection.py:145-149)."""
    if seed is not None:
        torch.manual_seed(seed)
    return torch.randperm(prob_logits.var.shape[0])[:k]
