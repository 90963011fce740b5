import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustbnns_b200 import lossGradients as lg
from robustbnns_b200.grid_search_halfMoons import MoonsBNN
h = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator().manual_seed(5)
x = torch.rand((100, 1, 2, 1), generator=g).cuda(); y = torch.randint(0, 2, (100,), generator=g).cuda()
mb = MoonsBNN(h, "leaky", "fc2", "hmc", None, None, 250, 5, 100, (1, 2, 1), 2)
mb.set_posterior_samples(torch.randn((250, mb.basenet.n_params), generator=g) / math.sqrt(h))
for _ in range(3): lg.expected_loss_gradients(mb, x, y, 250)
torch.cuda.synchronize()
