"""KFAC leg of the bench in isolation (burst regime): python scripts/kfac_leg_time.py [class_batches=2] [reps=4]"""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from bayesvlm_b200 import _lib
from bayesvlm_b200.hessians import kfac_ggn
from bayesvlm_b200.vlm import CLIP
cb = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kc = bench.KFAC
e_img, e_txt, a_img = bench.kfac_inputs(kc, cb * kc["num_classes"], kc["seed"], device="cuda")
vlm = CLIP(logit_scale=bench.LS, device="cuda")
run = lambda: kfac_ggn(vlm, kc["num_classes"], kc["batch_size"], e_img, a_img, e_txt, "cuda", "info_nce")
run(); res = []
for _ in range(reps):
    torch.cuda.synchronize(); time.sleep(1.0)
    run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        A, B = run()
    b.record(); torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / 3)
_lib.timing_enable(True); run(); torch.cuda.synchronize(); _lib.timing_enable(False)
print("ms per call", " ".join(f"{r:.3f}" for r in res), {k: round(v[1] / v[0], 4) for k, v in _lib.timing_collect().items()}, float(A.abs().max()))
