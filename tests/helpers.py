"""Shared test helpers (CPU side): load golden cases and rebuild the oracle's net for them."""
import os

import numpy as np
import torch

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SVI_CASES = ["svi_fc16_mnist", "svi_fc2_32_moons", "svi_conv16_mnist"]
HMC_CASES = ["hmc_fc16_fmnist", "hmc_fc2_16_mnist", "hmc_conv16_fmnist", "hmc_fc2_32_moons"]


class Case(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.z = z
        self.arch = str(z["arch"])
        self.input_shape = tuple(int(v) for v in z["input_shape"])
        self.hidden = int(z["hidden"])
        self.n_classes = int(z["n_classes"])
        self.dataset = "half_moons" if "moons" in name else "mnist"
        self.net = orc.build_net(self.arch, self.input_shape, self.hidden, self.n_classes,
                                 dataset_name=self.dataset)
        self.layout = orc.param_layout(self.net)
        self.x = torch.from_numpy(z["x"])
        self.y = torch.from_numpy(z["y"])
        self.labels = self.y.argmax(-1)
        self.bank = torch.from_numpy(z["bank"])

    def t(self, key):
        return torch.from_numpy(self.z[key])


def rel_err(a, b):
    """max|a-b| / max|b| -- the tolerance metric of BASELINE.json's north_star."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
