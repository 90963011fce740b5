"""Which side of fp32 / fp64 is the CUDA conv path on at the BASELINE configs[3] size?  The FULL workload (100
F-MNIST-shaped inputs x 50 stored samples, conv-512) on every engine against the oracle in fp32 AND fp64:
per-(sample, image) unit errors and the error of the expected gradient.  Scratch tool (GPU box):
    python scratch/conv_cfg4_oracle.py > gpurun_out/conv_cfg4_oracle.json"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import oracle as orc
from robustbnns_b200 import _lib
from robustbnns_b200.engine import Net

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B, S, H = int(os.environ.get("NB", "100")), int(os.environ.get("NS", "50")), 512
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
labels = torch.randint(0, 10, (B,), generator=g)
t0 = time.time()
ref64 = torch.stack([orc.expected_loss_gradients(net, layout, bank, x, labels, [s], dtype=torch.float64) for s in range(S)])
ref32 = torch.stack([orc.expected_loss_gradients(net, layout, bank, x, labels, [s]) for s in range(S)]).double()
t_oracle = time.time() - t0


def unit_err(a, ref):            # [S, B, ...] -> [S, B] max-norm error of a unit relative to that unit's maximum
    return (a - ref).abs().flatten(2).max(-1)[0] / ref.abs().flatten(2).max(-1)[0]


def mean_err(a, ref):
    return float((a.mean(0) - ref.mean(0)).abs().max() / ref.mean(0).abs().max())


out = {"workload": "conv-%d, %d inputs x %d samples, every (sample, input) unit" % (H, B, S), "oracle_seconds": t_oracle,
       "oracle_fp32_vs_fp64": {"units_above_1e-4": int((unit_err(ref32, ref64) > 1e-4).sum()),
                               "max_unit_err": float(unit_err(ref32, ref64).max()),
                               "expected_gradient_err": mean_err(ref32, ref64)}}
xd, ld = x.cuda(), labels.cuda().to(torch.int32)
for prec in ("fp32", "tf32x3", "f16x3"):
    eng = Net("conv", (1, 28, 28), H, 10)
    eng.set_precision(prec)
    eng.upload(bank, 0)
    got = torch.stack([eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, s, s + 1).cpu().reshape(x.shape) for s in range(S)]).double()
    full = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, 0, S).cpu().reshape(x.shape).double() / S
    e64, e32 = unit_err(got, ref64), unit_err(got, ref32)
    out[prec] = {"vs_fp64": {"units_above_1e-4": int((e64 > 1e-4).sum()), "max_unit_err": float(e64.max()),
                             "expected_gradient_err": float((full - ref64.mean(0)).abs().max() / ref64.mean(0).abs().max())},
                 "vs_fp32": {"units_above_1e-4": int((e32 > 1e-4).sum()), "max_unit_err": float(e32.max()),
                             "expected_gradient_err": float((full - ref32.mean(0)).abs().max() / ref32.mean(0).abs().max())},
                 "units_off_vs_both": int(((e64 > 1e-4) & (e32 > 1e-4)).sum()), "units": int(e64.numel())}
    eng.close()
print(json.dumps(out))
