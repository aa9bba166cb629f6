"""Error margins of the predictive precision modes vs the fp64 oracle on a row subset of the headline workload."""
import math, sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from oracle import laplace_oracle as O
from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
from bayesvlm_b200.vlm import CLIP, EncoderResult
cfg = dict(bench.PRED); cfg["N"] = 4096
t = bench.predictive_inputs(cfg, 0)
covs = bench.covariances(t, cfg, "cuda")
img = EncoderResult(t["img_e"].cuda(), t["img_a"].cuda()); txt = EncoderResult(t["txt_e"].cuda(), t["txt_a"].cuda())
rm, rv = O.predictive(t["img_e"].numpy(), t["img_a"].numpy(), t["txt_e"].numpy(), t["txt_a"].numpy(),
                      *(c.cpu().numpy() for c in covs), bench.LS, dtype=np.float64)
for prec in ("fp16x3", "fp16+fp8", "fp16"):
    m = CLIP(logit_scale=bench.LS, device="cuda", precision=prec)
    m.set_covariances(KFC(covs[0], covs[1]), KFC(covs[2], covs[3]))
    with torch.no_grad():
        out = m(img, txt)
    mean, var = out.mean.double().cpu().numpy(), out.var.double().cpu().numpy()
    dm = np.abs(mean - rm)
    tol = 1e-3 * np.maximum(np.abs(rm), 1.0)
    print(prec, "mean: max abs err %.3g, max err/tol %.3g, rel fro %.3g | var: max rel %.3g" %
          (dm.max(), (dm / tol).max(), np.linalg.norm(mean - rm) / np.linalg.norm(rm), (np.abs(var - rv) / rv).max()))
