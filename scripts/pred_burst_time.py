"""Burst-regime time of the headline predictive step (what `bench.py --steps 20 --warmup 5` sees: the board idles, then runs
25 steps -- 7.5 ms -- long before the 1 kW power controller reacts, scripts/power_trace.py):
    python scripts/pred_burst_time.py [steps] [reps] [probs]"""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from bayesvlm_b200 import _lib
from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
from bayesvlm_b200.vlm import CLIP, EncoderResult
cfg = bench.PRED
t = bench.predictive_inputs(cfg, 0)
Ai, Bi, At, Bt = bench.covariances(t, cfg, "cuda")
m = CLIP(logit_scale=bench.LS, device="cuda")
m.set_covariances(KFC(Ai, Bi), KFC(At, Bt))
img = EncoderResult(t["img_e"].cuda(), t["img_a"].cuda()); txt = EncoderResult(t["txt_e"].cuda(), t["txt_a"].cuda())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mean = torch.empty((cfg["N"], cfg["C"]), device="cuda"); var = torch.empty_like(mean)
probs = torch.empty_like(mean) if len(sys.argv) > 3 and sys.argv[3] == "probs" else None
res = []
with torch.no_grad():
    for rep in range(reps):
        torch.cuda.synchronize(); time.sleep(1.0)
        for _ in range(5):
            m._smith_into(img.embeds, img.activations, txt, mean, var, probs)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            m._smith_into(img.embeds, img.activations, txt, mean, var, probs)
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / n)
    time.sleep(1.0)
    _lib.timing_enable(True)
    for _ in range(n):
        m._smith_into(img.embeds, img.activations, txt, mean, var, probs)
    torch.cuda.synchronize()
    _lib.timing_enable(False)
    kk = _lib.timing_collect()
print("burst ms/step", " ".join(f"{r:.4f}" for r in res), "| kernels", {k: round(v[1] / v[0], 4) for k, v in kk.items()},
      "checksum", float(mean[0, 0]), float(var[-1, -1]))
