"""Shared test helpers (CPU side): load golden cases and rebuild the oracle's net for them."""
import os

import numpy as np
import torch

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SVI_CASES = ["svi_fc16_mnist", "svi_fc2_32_moons", "svi_conv16_mnist"]
HMC_CASES = ["hmc_fc16_fmnist", "hmc_fc2_16_mnist", "hmc_conv16_fmnist", "hmc_fc2_32_moons"]
ENS_CASES = ["ens_fc2_16_mnist", "ens_conv16_fmnist", "ens_fc16_mnist"]      # tests/golden/make_golden_ensemble.py


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


class OracleEngine(object):
    """CPU stand-in for robustbnns_b200.engine.Net used ONLY to test the host logic of the
    drop-in classes (sample placement, sharding, allreduce plumbing) without a GPU.  It answers
    the Net interface with the oracle; it is test infrastructure and never ships."""

    def __init__(self, arch, input_shape, hidden, n_classes, dataset="mnist", seed_log=None):
        self.device = torch.device("cpu")
        self.net = orc.build_net(arch, input_shape, hidden, n_classes, dataset_name=dataset)
        self.layout = orc.param_layout(self.net)
        self.P = orc.param_count(self.layout)
        self.input_shape, self.n_classes = tuple(input_shape), n_classes
        self.D = int(np.prod(input_shape))
        self.bank = torch.zeros((0, self.P))
        self.sampled = []            # (seed, global index, row) log

    @property
    def capacity(self):
        return self.bank.shape[0]

    def reserve(self, capacity):
        if capacity > self.bank.shape[0]:
            self.bank = torch.cat([self.bank, torch.zeros((capacity - self.bank.shape[0], self.P))])

    def upload(self, weights, s0=0):
        w = torch.as_tensor(weights, dtype=torch.float32).reshape(-1, self.P)
        self.reserve(s0 + w.shape[0])
        self.bank[s0:s0 + w.shape[0]] = w

    def sample_diag(self, loc, rho, seed, sample_index0, s0, count, stride=1):
        self.reserve(s0 + count)
        ids = [sample_index0 + i * stride for i in range(count)]
        self.bank[s0:s0 + count] = orc.philox_bank(loc.cpu(), rho.cpu(), seed, ids)
        self.sampled += [(seed, g, s0 + i) for i, g in enumerate(ids)]

    def _x(self, x):
        return torch.as_tensor(x, dtype=torch.float32).reshape(-1, *self.input_shape)

    keep_valid, keep_serial = False, 0   # no kept-forward route: the drop-in takes the two-pass path

    def forward_probs_sum(self, x, s0, s1, keep=False):
        x = self._x(x)
        if s1 == s0:
            return torch.zeros((x.shape[0], self.n_classes))
        return orc.bnn_forward(self.net, self.layout, self.bank, x, range(s0, s1)) * (s1 - s0)

    def forward_logits(self, x, s):
        return orc.bnn_forward_avg_posterior(self.net, self.layout, self.bank[s], self._x(x))

    def forward_logits_sum(self, x, s0, s1):
        return orc.ensemble_forward(self.net, self.layout, self.bank, self._x(x), range(s0, s1)) * (s1 - s0)

    def set_best_precision(self):
        return "oracle"

    def input_grad_sum(self, head, x, labels, s0, s1, pbar=None, out=None):
        x = self._x(x)
        if s1 == s0:
            g = torch.zeros_like(x)
        else:
            labels = torch.as_tensor(labels).long()
            with torch.enable_grad():
                g = self._grad(head, x, labels, s0, s1, pbar)
        if out is not None:
            out.copy_(g.reshape(out.shape))
            return out
        return g

    def _grad(self, head, x, labels, s0, s1, pbar):
        if head == 0:
            return orc.expected_loss_gradients(self.net, self.layout, self.bank, x, labels, range(s0, s1)) * (s1 - s0)
        xs = x.clone().requires_grad_(True)
        import torch.nn.functional as nnf
        if head == 4:                     # LOGITS_UPSTREAM: g goes to the (sum of the) logits as is
            ls = sum(orc.net_logits(self.net, orc.unpack(self.bank[s], self.layout), xs) for s in range(s0, s1))
            (gx,) = torch.autograd.grad((ls * torch.as_tensor(pbar).detach()).sum(), xs)
            return gx
        ps = torch.stack([nnf.softmax(orc.net_logits(self.net, orc.unpack(self.bank[s], self.layout), xs), -1)
                          for s in range(s0, s1)]).sum(0)
        if head == 1:
            g = nnf.softmax(torch.as_tensor(pbar), -1) - nnf.one_hot(labels, self.n_classes)
        else:
            g = torch.as_tensor(pbar)
        (gx,) = torch.autograd.grad((ps * g.detach()).sum(), xs)
        return gx

    # stateless kernels, torch restatements of adversarialAttacks.py:81-82, :89, :103-105, :179
    @staticmethod
    def fgsm_step(x, grad, eps):
        return torch.clamp(x + eps * grad.sign(), 0, 1)

    @staticmethod
    def pgd_alpha(x):
        return 2 / x.max(dim=1)[0]

    @staticmethod
    def pgd_step(x, x0, grad, alpha, eps):
        eta = torch.clamp(x + alpha[:, None] * grad.sign() - x0, min=-eps, max=eps)
        return torch.clamp(x0 + eta, min=0, max=1)

    @staticmethod
    def count_correct(out, labels, counter):
        counter += (out.argmax(-1) == labels.long()).sum()
