"""Half-moons fc2 2-H-H-2 BNN: expected loss gradients (100 points x 250 stored samples) per width and engine."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustbnns_b200 import lossGradients as lg
from robustbnns_b200.grid_search_halfMoons import MoonsBNN
g = torch.Generator().manual_seed(5)
x = torch.rand((100, 1, 2, 1), generator=g).cuda(); y = torch.randint(0, 2, (100,), generator=g).cuda()
out = {}
for h in (32, 64, 128, 256, 512):
    mb = MoonsBNN(h, "leaky", "fc2", "hmc", None, None, 250, 5, 100, (1, 2, 1), 2)
    mb.set_posterior_samples(torch.randn((250, mb.basenet.n_params), generator=g) / math.sqrt(h))
    row = {}
    for prec in ("fp32", "tf32x3"):
        mb.set_precision(prec)
        lg.expected_loss_gradients(mb, x, y, 250); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): r = lg.expected_loss_gradients(mb, x, y, 250)
        e1.record(); torch.cuda.synchronize()
        row[prec] = round(e0.elapsed_time(e1) / 10, 4)
    out[h] = row
print(json.dumps(out))
