"""Launch-list probe of the Bayesian PGD loop (BASELINE configs[2] shape): run under
   ncu --metrics gpu__time_duration.sum --csv ...   (scratch tool, not part of the product)"""
import math, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustbnns_b200 import adversarialAttacks as aa
from robustbnns_b200.model_bnn import BNN

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_img, n_s = 1000, 100
g = torch.Generator().manual_seed(1)
bnn = BNN("mnist", 512, "leaky", "fc", "svi", 1, 0.01, None, None, (1, 28, 28), 10)
locs, rhos, fan = [], [], 784
for key, shp in bnn.basenet.layout:
    n = 1
    for v in shp:
        n *= v
    if len(shp) > 1:
        fan = shp[1]
    locs.append(torch.randn(n, generator=g) / math.sqrt(fan))
    rhos.append(torch.randn(n, generator=g) - 5.0)
bnn.set_guide(torch.cat(locs), torch.cat(rhos))
bnn.set_precision(os.environ.get("PREC", "f16x3"))
x = torch.rand((n_img, 1, 28, 28), generator=g).cuda()
y = torch.randint(0, 10, (n_img,), generator=g).cuda()
aa.pgd_attack(bnn, x, y, hyperparams=None, n_samples=n_s, iters=iters)   # same count: the CUDA graph is captured here
torch.cuda.synchronize()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
aa.pgd_attack(bnn, x, y, hyperparams=None, n_samples=n_s, iters=iters)
e1.record()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print("pgd %d iters: device %.3f ms (%.3f ms/iter), host enqueue %.3f ms" % (iters, e0.elapsed_time(e1), e0.elapsed_time(e1) / iters, 1e3 * t_host))
