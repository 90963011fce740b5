"""Bayesian PGD over N ranks (launch with torchrun): input sharding (attack()'s default: every rank draws the same samples,
attacks its block, one all-gather) against sample sharding (two all-reduces per iteration).  Scratch tool."""
import json, math, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from robustbnns_b200 import adversarialAttacks as aa
from robustbnns_b200.model_bnn import BNN

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
os.chdir(tempfile.mkdtemp())
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
x = torch.rand((n_img, 1, 28, 28), generator=g)
y = torch.nn.functional.one_hot(torch.randint(0, 10, (n_img,), generator=g), 10).float()
out = {"world": world, "images": n_img, "posterior_samples": n_s, "iters": 40}
res = {}
for mode in ("inputs", "samples"):
    bnn.attack_sharding = mode
    for rep in range(2):                      # first pass warms up (allocations, scale freeze)
        bnn.reseed(0)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        adv = aa.attack(net=bnn, x_test=x, y_test=y, dataset_name="mnist", device="cuda", method="pgd", filename="a",
                        savedir="a", hyperparams=None, n_samples=n_s)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
    res[mode] = adv
    out[mode] = {"s": dt, "imgs_per_s": n_img / dt}
out["mismatch_fraction"] = float(((res["inputs"] - res["samples"]).abs() > 1e-6).float().mean())
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
