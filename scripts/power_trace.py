"""SM clock / board power / throttle reasons sampled through NVML every 5 ms while the headline predictive step (or the GGN
class batch) loops for a few seconds:  python scripts/power_trace.py [pred|ggn] [seconds]"""
import sys, threading, time, torch
sys.path.insert(0, ".")
import bench, pynvml
from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC, compute_hessian_analytic_InfoNCE
from bayesvlm_b200.vlm import CLIP, EncoderResult

what = sys.argv[1] if len(sys.argv) > 1 else "pred"
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], threading.Event()

def sampler():
    t0 = time.perf_counter()
    while not stop.is_set():
        samples.append((time.perf_counter() - t0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.005)

if what == "pred":
    cfg = bench.PRED
    t = bench.predictive_inputs(cfg, 0)
    Ai, Bi, At, Bt = bench.covariances(t, cfg, "cuda")
    m = CLIP(logit_scale=bench.LS, device="cuda")
    m.set_covariances(KFC(Ai, Bi), KFC(At, Bt))
    img = EncoderResult(t["img_e"].cuda(), t["img_a"].cuda()); txt = EncoderResult(t["txt_e"].cuda(), t["txt_a"].cuda())
    mean = torch.empty((cfg["N"], cfg["C"]), device="cuda"); var = torch.empty_like(mean)
    fn = lambda: m._smith_into(img.embeds, img.activations, txt, mean, var)
    group = 100
else:
    e_img, e_txt, _ = bench.kfac_inputs(bench.KFAC, 32768, 1, device="cuda")
    ls = torch.tensor(bench.LS, device="cuda")
    fn = lambda: compute_hessian_analytic_InfoNCE(e_img, e_txt, ls)
    group = 5
with torch.no_grad():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    time.sleep(2.0)  # idle: let the power averaging window drain
    thr = threading.Thread(target=sampler, daemon=True); thr.start()
    time.sleep(0.1)
    t_start = time.perf_counter()
    times = []
    while time.perf_counter() - t_start < secs:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(group):
            fn()
        b.record(); torch.cuda.synchronize()
        times.append((time.perf_counter() - t_start, a.elapsed_time(b) / group))
    stop.set(); thr.join()
print("limit W", pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0, "max sm MHz", pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
print("t_s  ms_per_call")
for t_, ms in times[:: max(1, len(times) // 12)]:
    print(f"{t_:6.3f} {ms:8.4f}")
print("t_s  sm_mhz  power_W  reasons")
for s in samples[:: max(1, len(samples) // 30)]:
    print(f"{s[0]:6.3f} {s[1]:5d} {s[2]:7.1f} {s[3]:#x}")
