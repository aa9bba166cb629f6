"""The bench's SigLIP KFAC leg alone (2 class batches of 32768, D=768, d_img=3072+1): python scripts/kfac_siglip_leg.py"""
import sys, torch
sys.path.insert(0, ".")
import bench
from bayesvlm_b200 import _lib
from bayesvlm_b200.hessians import kfac_ggn, syrk_accumulate
from bayesvlm_b200.vlm import SIGLIP
kc = bench.KFAC_SIGLIP
n = 2 * kc["num_classes"]
e_img, e_txt, a_img = bench.kfac_inputs(kc, n, kc["seed"], device="cuda")[:3]
vlm = SIGLIP(logit_scale=kc["logit_scale"], logit_bias=kc["logit_bias"], device="cuda")
kw = dict(num_classes=kc["num_classes"], batch_size=kc["batch_size"], device="cuda", likelihood="siglip")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("kfac_ggn leg ms", t(lambda: kfac_ggn(vlm, source_embeds=e_img, source_activations=a_img, target_embeds=e_txt, **kw)))
print("syrk alone ms", t(lambda: syrk_accumulate(a_img, append_one=True)))
_lib.timing_enable(True); syrk_accumulate(a_img, append_one=True); torch.cuda.synchronize(); _lib.timing_enable(False)
print({k: round(v[1] / v[0], 3) for k, v in _lib.timing_collect().items()})
