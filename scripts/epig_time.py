"""EPIG scoring time at the bench shape (pool 100 000 x target 10 000, K=100, chunk 4096): python scripts/epig_time.py [Cl=10] [pool]"""
import sys, torch
sys.path.insert(0, ".")
import bench
from bayesvlm_b200 import _lib
from bayesvlm_b200.epig import epig_from_logits_using_matmul
from bayesvlm_b200.vlm import ProbabilisticLogits
ec = dict(bench.EPIG)
ec["Cl"] = int(sys.argv[1]) if len(sys.argv) > 1 else 10
if len(sys.argv) > 2: ec["pool"] = int(sys.argv[2])
gen = torch.Generator(device="cuda").manual_seed(1)
mk = lambda n: ProbabilisticLogits(torch.randn(n, ec["Cl"], generator=gen, device="cuda") * 2,
                                   torch.rand(n, ec["Cl"], generator=gen, device="cuda") * 3 + 0.1)
lp, lt = mk(ec["pool"]), mk(ec["target"])
for rep in range(3):
    torch.cuda.synchronize(); _lib.timing_enable(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s = epig_from_logits_using_matmul(lp, lt, seed=0, num_samples=ec["K"], chunk_size=ec["chunk"]); b.record()
    torch.cuda.synchronize(); _lib.timing_enable(False); k = _lib.timing_collect()
    print("total %.3f ms, joint kernel %.3f ms, prepare %.3f ms" % (a.elapsed_time(b), k.get("epig_joint", (0, 0.0))[1],
                                                                    k.get("epig_prepare", (0, 0.0))[1]), float(s.sum()))
