"""kfac_ggn end to end from HOST tensors (the reference's calling convention): pageable vs pinned inputs, 8 class batches."""
import math, sys, time, torch
sys.path.insert(0, ".")
import bench
from bayesvlm_b200.hessians import kfac_ggn
from bayesvlm_b200.vlm import CLIP
kc = bench.KFAC
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 8) * kc["num_classes"]
e_img, e_txt, a_img = bench.kfac_inputs(kc, n, kc["seed"])
vlm = CLIP(logit_scale=bench.LS, device="cuda")
kw = dict(num_classes=kc["num_classes"], batch_size=kc["batch_size"], device="cuda", likelihood="info_nce")
for name, place in (("device", lambda t: t.cuda()), ("pinned", lambda t: t.pin_memory()), ("pageable", lambda t: t)):
    s, a, t = place(e_img), place(a_img), place(e_txt)
    kfac_ggn(vlm, source_embeds=s, source_activations=a, target_embeds=t, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        kfac_ggn(vlm, source_embeds=s, source_activations=a, target_embeds=t, **kw)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"{name:9s} {dt * 1e3:8.2f} ms for {n} samples = {n / dt:.3g} samples/s", flush=True)
