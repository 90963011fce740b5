"""CPU check of a hypothesis: in the conv-512 cfg4 workload, how many stride-1 pooling windows of the exact (fp64) A2 have
a top-2 pair that collapses to the same fp32 value (a tie the fp32-stored A2 cannot order)?"""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from oracle import oracle as orc
torch.set_num_threads(16)
B, S, H = 100, 50, 512
net = orc.build_net("conv", (1, 28, 28), H, 10)
layout = orc.param_layout(net)
g = torch.Generator().manual_seed(21)
cols = []
for key, shp in layout:
    n = int(np.prod(shp))
    fan = n // shp[0] if len(shp) > 1 else 25
    cols.append(torch.randn((S, n), generator=g) / math.sqrt(fan))
bank = torch.cat(cols, dim=1)
x = torch.rand((B, 1, 28, 28), generator=g)
tot_units, tot_windows = set(), 0
for s in range(S):
    w = {k: v.double() for k, v in orc.unpack(bank[s], layout).items()}
    keys = list(w.keys())
    a1 = F.leaky_relu(F.conv2d(x.double(), w[keys[0]], w[keys[1]]), 0.01)
    p1 = F.max_pool2d(a1, 2)
    a2 = F.leaky_relu(F.conv2d(p1, w[keys[2]], w[keys[3]]), 0.01)      # [B, H, 8, 8]
    win = a2.unfold(2, 2, 1).unfold(3, 2, 1).reshape(B, H, 7, 7, 4)
    top2 = win.topk(2, dim=-1).values
    tie32 = (top2[..., 0].float() == top2[..., 1].float()) & (top2[..., 0] != top2[..., 1])
    # which of those would be decided wrongly by "first maximal entry wins" on the fp32 values?
    idx64 = win.argmax(-1)
    idx32 = win.float().argmax(-1)
    wrong = (idx64 != idx32)
    tot_windows += int(wrong.sum())
    for b in wrong.flatten(1).any(1).nonzero().flatten().tolist():
        tot_units.add((s, b))
    print(s, int(tie32.sum()), int(wrong.sum()), flush=True)
print("windows decided differently by fp32-rounded exact A2:", tot_windows, "units:", len(tot_units))
