"""bayesvlm_b200 -- B200-native (sm_100a) kernels behind BayesVLM's post-hoc Laplace hot path.

Module layout mirrors the reference package for this path only:
``hessians`` (KFAC / GGN estimation + covariance plumbing), ``vlm`` (probabilistic-logit interface),
``epig`` (EPIG acquisition), ``precompute`` (the ``make_predictions`` driver).  All compute goes through
``libbvlm.so`` (C ABI in ``include/bvlm.h``); there is no CPU or PyTorch fallback for the kernels.
"""
__version__ = "0.1.0"
