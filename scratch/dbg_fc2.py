import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import oracle as orc
from tests.helpers import rel_err
from tests.test_gpu_parity import _problem
from robustbnns_b200 import _lib
from robustbnns_b200.engine import Net
for (arch, hidden, B, S) in [("fc2", 512, 257, 6), ("fc2", 512, 256, 6), ("fc2", 256, 257, 6), ("fc2", 512, 64, 2)]:
    net, layout, loc, rho, bank, x, labels = _problem(arch, (1, 28, 28), hidden, 10, B, S)
    eng = Net(arch, (1, 28, 28), hidden, 10)
    eng.set_precision("tf32x3")
    eng.upload(bank, 0)
    tot = 0
    for s in range(S):
        g = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, s, s + 1).cpu().reshape(x.shape)
        r = orc.expected_loss_gradients(net, layout, bank, x, labels, [s], dtype=torch.float64)
        tot = tot + g
        print(arch, hidden, B, "single s", s, rel_err(g, r))
    g = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S).cpu().reshape(x.shape)
    r = orc.expected_loss_gradients(net, layout, bank, x, labels, range(S), dtype=torch.float64) * S
    print(arch, hidden, B, "all", rel_err(g, r), "sum of singles", rel_err(tot, r), "all-vs-singles", rel_err(g, tot))
    d = (g.double() - r).abs().flatten(1).max(dim=1)[0]
    print(" worst rows", torch.topk(d, 5))
    eng.set_precision("fp32")
    g32 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S).cpu().reshape(x.shape)
    print(" fp32 engine all", rel_err(g32, r))
    eng.close()
