import math, sys, time, torch
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
    for _ in range(5): out = m(img, txt)
    torch.cuda.synchronize()
    for n in (1, 200):
        t0 = time.perf_counter()
        for _ in range(n): out = m(img, txt)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(n, "host enqueue per call %.1f us, total per call %.1f us" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
    # graph capture of one call with static outputs
    mean = torch.empty(cfg["N"], cfg["C"], device="cuda"); var = torch.empty_like(mean)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m._smith_into(img.embeds, img.activations, txt, mean, var)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        m._smith_into(img.embeds, img.activations, txt, mean, var)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(200): g.replay()
    e.record(); torch.cuda.synchronize()
    print("graph replay per call %.1f us" % (b.elapsed_time(e) / 200 * 1e3), float((mean - out.mean).abs().max()))
