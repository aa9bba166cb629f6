"""One EPIG scoring pass at the bench shape (pool 4096 x target 10000, Cl=10, K=100) -- profiling target for ncu."""
import sys, torch
sys.path.insert(0, ".")
import bench
from bayesvlm_b200.epig import epig_from_logits_using_matmul
from bayesvlm_b200.vlm import ProbabilisticLogits
ec = bench.EPIG
CL = int(sys.argv[2]) if len(sys.argv) > 2 else 10
gen = torch.Generator(device="cuda").manual_seed(1)
mk = lambda n: ProbabilisticLogits(torch.randn(n, CL, generator=gen, device="cuda") * 2,
                                   torch.rand(n, CL, generator=gen, device="cuda") * 3 + 0.1)
lp, lt = mk(4096), mk(ec["target"])
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    s = epig_from_logits_using_matmul(lp, lt, seed=0, num_samples=ec["K"], chunk_size=ec["chunk"])
torch.cuda.synchronize()
print("ok", float(s.max()))
