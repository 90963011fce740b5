"""Timing / launch-list probe of the conv BNN (BASELINE configs[3] shape: 100 F-MNIST-shaped inputs, conv hidden 512,
HMC bank of 50 posterior samples): expected loss gradients and Bayesian PGD on the FP32 (CUDA-core) and TF32X3
(tcgen05 implicit-GEMM) engines.  Scratch tool, not part of the product:
    python scratch/conv_probe.py            # prints one JSON line
    ncu --metrics gpu__time_duration.sum --csv ... python scratch/conv_probe.py tf32x3 1     # launch list"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustbnns_b200 import adversarialAttacks as aa
from robustbnns_b200 import lossGradients as lg
from robustbnns_b200.model_bnn import BNN

precs = [sys.argv[1]] if len(sys.argv) > 1 else ["fp32", "tf32x3", "f16x3"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
hidden = int(os.environ.get("HIDDEN", "512"))
n_img, n_s = 100, 50
FLOP_PER_UNIT = {512: 107704320, 1024: 213565440}.get(hidden)
g = torch.Generator().manual_seed(1)
bnn = BNN("fashion_mnist", hidden, "leaky", "conv", "hmc", None, None, n_s, 5, (1, 28, 28), 10)
rows, fan = [], 25
for key, shp in bnn.basenet.layout:
    n = 1
    for v in shp:
        n *= v
    if len(shp) > 1:
        fan = n // shp[0]
    rows.append(torch.randn((n_s, n), generator=g) / math.sqrt(fan))
bnn.set_posterior_samples(torch.cat(rows, dim=1))
x = torch.rand((n_img, 1, 28, 28), generator=g).cuda()
y = torch.randint(0, 10, (n_img,), generator=g).cuda()
out = {"workload": "conv-%d BNN, %d inputs x %d HMC samples" % (hidden, n_img, n_s)}
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
    r = {"grad_ms": ms, "grads_per_s": n_img * n_s / (ms * 1e-3)}
    if FLOP_PER_UNIT:
        r["tflops_algorithmic"] = n_img * n_s * FLOP_PER_UNIT / (ms * 1e-3) / 1e12
    if ref is None:
        ref = gr.clone()
    else:
        r["max_rel_deviation_from_" + precs[0]] = float((gr - ref).abs().max() / ref.abs().max())
    iters = 4
    aa.pgd_attack(bnn, x, y, hyperparams={"epsilon": 0.2}, n_samples=n_s, iters=iters)   # also captures the graph
    torch.cuda.synchronize()
    e0.record()
    aa.pgd_attack(bnn, x, y, hyperparams={"epsilon": 0.2}, n_samples=n_s, iters=iters)
    e1.record()
    torch.cuda.synchronize()
    r["pgd_ms_per_iter"] = e0.elapsed_time(e1) / iters
    r["pgd40_imgs_per_s"] = n_img / (r["pgd_ms_per_iter"] * 40 * 1e-3)
    out[prec] = r
print(json.dumps(out))
