"""Per-(sample, image) error of the conv engines against the fp64 oracle on an INDEPENDENT bank (rows ~ N(0, 1/fan_in)):
locates discrete flips (a few bad rows) vs continuous error (all rows alike).  Scratch tool (GPU box)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as orc
from robustbnns_b200 import _lib
from robustbnns_b200.engine import Net

hidden = int(os.environ.get("HIDDEN", "512"))
B, S = int(os.environ.get("NB", "8")), int(os.environ.get("NS", "6"))
shape, C = (1, 28, 28), 10
net = orc.build_net("conv", shape, hidden, C)
layout = orc.param_layout(net)
seed = int(os.environ.get("SEED", "1"))
g = torch.Generator().manual_seed(seed)
rows = []
for key, shp in layout:
    n = 1
    for v in shp:
        n *= v
    fan = n // shp[0] if len(shp) > 1 else 25
    rows.append(torch.randn((S, n), generator=g) / math.sqrt(fan))
bank = torch.cat(rows, dim=1)
x = torch.rand((B, *shape), generator=g)
labels = torch.randint(0, C, (B,), generator=g)
if os.environ.get("IDX"):                      # keep a subset of the images / samples (as the full-size test does)
    idx = torch.tensor([int(v) for v in os.environ["IDX"].split(",")])
    x, labels, B = x[idx], labels[idx], len(idx)
s_lo, s_hi = int(os.environ.get("S0", "0")), int(os.environ.get("S1", str(S)))
bank = bank[s_lo:s_hi]
S = s_hi - s_lo
ref = torch.stack([orc.expected_loss_gradients(net, layout, bank, x, labels, [s], dtype=torch.float64) for s in range(S)])
ref32 = torch.stack([orc.expected_loss_gradients(net, layout, bank, x, labels, [s]) for s in range(S)]).double()
print("oracle fp32 vs fp64 per-row max rel err: %.3e" % float(((ref32 - ref).abs().flatten(2).max(-1)[0] / ref.abs().flatten(2).max(-1)[0]).max()))
eng = Net("conv", shape, hidden, C)
eng.upload(bank, 0)
for prec in ("fp32", "tf32x3"):
    eng.set_precision(prec)
    got = torch.stack([eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, s, s + 1).cpu().reshape(x.shape) for s in range(S)]).double()
    err = (got - ref).abs().flatten(2).max(-1)[0] / ref.abs().flatten(2).max(-1)[0]      # [S, B]
    print(prec, "per-row rel err: median %.3e  max %.3e  rows > 1e-4: %d of %d" % (float(err.median()), float(err.max()), int((err > 1e-4).sum()), err.numel()))
    mean_err = float((got.mean(0) - ref.mean(0)).abs().max() / ref.mean(0).abs().max())
    print(prec, "mean-of-grads rel err %.3e" % mean_err)
    if prec == "tf32x3":
        worst = torch.nonzero(err > 1e-4)
        print("bad rows (s, b):", worst[:10].tolist())
        print(err)
