"""Timing / launch-list probe of the fc2 BNN (saved_BNNs model_1 shape: fc2-512, HMC bank of 100 samples, 1000 inputs):
expected loss gradients and Bayesian PGD per engine.  Scratch tool."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustbnns_b200 import adversarialAttacks as aa
from robustbnns_b200 import lossGradients as lg
from robustbnns_b200.model_bnn import BNN

precs = [sys.argv[1]] if len(sys.argv) > 1 else ["fp32", "tf32x3", "bf16"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
hidden = int(os.environ.get("HIDDEN", "512"))
n_img, n_s = int(os.environ.get("NB", "1000")), int(os.environ.get("NS", "100"))
FLOP = 2 * 2 * (784 * hidden + hidden * hidden + hidden * 10)
g = torch.Generator().manual_seed(1)
bnn = BNN("mnist", hidden, "leaky", "fc2", "hmc", None, None, n_s, 5, (1, 28, 28), 10)
rows, fan = [], 784
for key, shp in bnn.basenet.layout:
    n = 1
    for v in shp:
        n *= v
    if len(shp) > 1:
        fan = shp[1]
    rows.append(torch.randn((n_s, n), generator=g) / math.sqrt(fan))
bnn.set_posterior_samples(torch.cat(rows, dim=1))
x = torch.rand((n_img, 1, 28, 28), generator=g).cuda()
y = torch.randint(0, 10, (n_img,), generator=g).cuda()
out = {"workload": "fc2-%d BNN, %d inputs x %d HMC samples" % (hidden, n_img, n_s)}
ref = None
for prec in precs:
    bnn.set_precision(prec)
    gr = lg.expected_loss_gradients(bnn, x, y, n_s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr = lg.expected_loss_gradients(bnn, x, y, n_s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    r = {"grad_ms": ms, "grads_per_s": n_img * n_s / (ms * 1e-3), "tflops_algorithmic": n_img * n_s * FLOP / (ms * 1e-3) / 1e12}
    if ref is None:
        ref = gr.clone()
    else:
        r["max_rel_deviation_from_" + precs[0]] = float((gr - ref).abs().max() / ref.abs().max())
    if os.environ.get("ORACLE", "1") != "0":          # 64 rows against the fp64 oracle (the north-star tolerance is 1e-4)
        from oracle import oracle as orc
        net = orc.build_net("fc2", (1, 28, 28), hidden, 10)
        layout = orc.param_layout(net)
        idx = torch.arange(0, n_img, max(1, n_img // 64))[:64]
        bank = bnn.engine().download(0, n_s)
        r64 = orc.expected_loss_gradients(net, layout, bank, x.cpu()[idx], y.cpu()[idx], range(n_s), dtype=torch.float64)
        r["max_rel_err_vs_fp64_oracle_64_rows"] = float((gr.cpu()[idx].double() - r64).abs().max() / r64.abs().max())
    aa.pgd_attack(bnn, x, y, hyperparams={"epsilon": 0.2}, n_samples=n_s, iters=4)   # also captures the graph
    torch.cuda.synchronize()
    e0.record()
    aa.pgd_attack(bnn, x, y, hyperparams={"epsilon": 0.2}, n_samples=n_s, iters=4)
    e1.record()
    torch.cuda.synchronize()
    r["pgd_ms_per_iter"] = e0.elapsed_time(e1) / 4
    out[prec] = r
print(json.dumps(out))
