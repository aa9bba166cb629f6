"""A few predictive steps at the headline shape -- profiling target for ncu."""
import math, sys, torch
sys.path.insert(0, ".")
import bench
from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
from bayesvlm_b200.vlm import CLIP, EncoderResult
cfg = bench.PRED
t = bench.predictive_inputs(cfg, 0)
Ai, Bi, At, Bt = bench.covariances(t, cfg, "cuda")
m = CLIP(logit_scale=bench.LS, device="cuda")
m.set_covariances(KFC(Ai, Bi), KFC(At, Bt))
img = EncoderResult(t["img_e"].cuda(), t["img_a"].cuda()); txt = EncoderResult(t["txt_e"].cuda(), t["txt_a"].cuda())
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        out = m(img, txt)
torch.cuda.synchronize(); print("ok", float(out.mean[0, 0]))
