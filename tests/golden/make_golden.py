"""Generate the golden vectors under tests/golden/ by running the reference's OWN code.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference modules (model_nn, model_bnn, lossGradients, adversarialAttacks)
are imported unmodified from /root/reference with `oracle/pyro_shim` ahead of
them on sys.path (Pyro 1.3.0 / keras / matplotlib are not installed here; the
shim restates the few Pyro entry points involved and turns plotting into
no-ops).  Every array stored below is an output of the reference's functions:

  BNN.forward (seeded SVI, avg_posterior, HMC)      model_bnn.py:198-258
  BNN.evaluate                                      model_bnn.py:367-391
  loss_gradient / loss_gradients                    lossGradients.py:20-68
  fgsm_attack / pgd_attack / attack                 adversarialAttacks.py:69-143
  attack_evaluation / softmax_robustness            adversarialAttacks.py:30-62,151-198

plus the inputs needed to replay them (guide parameters, the posterior-sample
bank the reference drew, images, labels).
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyro_shim"))
sys.path.insert(0, ROOT)

import pyro  # noqa: E402  (the shim)
import model_bnn  # noqa: E402  (reference)
import lossGradients  # noqa: E402  (reference)
import adversarialAttacks  # noqa: E402  (reference)
from torch.utils.data import DataLoader  # noqa: E402

from oracle import oracle as orc  # noqa: E402  (only for the synthetic-problem generators)

torch.set_num_threads(4)


def _layout_of(bnn):
    return [(k, tuple(v.shape)) for k, v in bnn.basenet.state_dict().items()]


def _install_guide_params(layout, loc, rho):
    pyro.clear_param_store()
    off = 0
    for key, shp in layout:
        n = int(np.prod(shp))
        pyro.param(f"{key}_loc", loc[off:off + n].reshape(shp).clone())
        pyro.param(f"{key}_scale", rho[off:off + n].reshape(shp).clone())
        off += n


class _DrawLog(object):
    """Records every weight draw random_module makes, in order (one bank row per guide call)."""

    def __init__(self, layout):
        self.layout, self.rows, self._cur = layout, [], []
        self._orig = pyro.sample

    def __enter__(self):
        def logged(name, fn, obs=None, **kw):
            v = self._orig(name, fn, obs=obs, **kw)
            if name.startswith("module$$$"):
                self._cur.append(v.detach().reshape(-1).clone())
                if len(self._cur) == len(self.layout):
                    self.rows.append(torch.cat(self._cur))
                    self._cur = []
            return v
        pyro.sample = logged
        return self

    def __exit__(self, *exc):
        pyro.sample = self._orig
        return False


def _half_predicted_labels(bnn, x, y, n_samp, seeds):
    """Random nets classify random labels at chance; to make the accuracy counts
    informative, images 0,1,3,4,6,7,... get the label the BNN itself predicts."""
    pred = bnn.forward(x, n_samples=n_samp, seeds=seeds).detach().argmax(-1)
    y = y.clone()
    for i in range(len(x)):
        if i % 3 != 2:
            y[i] = torch.nn.functional.one_hot(pred[i], y.shape[1]).float()
    return y


def svi_case(name, arch, input_shape, hidden, n_classes, n_img, n_samp, dataset="mnist"):
    bnn = model_bnn.BNN(dataset, hidden, "leaky", arch, "svi", 1, 0.01, None, None, input_shape, n_classes)
    bnn.device = "cpu"
    bnn.basenet.device = "cpu"
    layout = _layout_of(bnn)
    loc, rho = orc.scaled_guide_params(layout, seed=1, rho_mean=-3.0)
    _install_guide_params(layout, loc, rho)
    x, y = orc.synthetic_inputs(n_img, input_shape, n_classes, seed=0)
    seeds = list(range(n_samp))
    y = _half_predicted_labels(bnn, x, y, n_samp, seeds)

    # the bank the reference draws under seeds 0..S-1  (model_bnn.py:222-226)
    with _DrawLog(layout) as log:
        probs_seeded = bnn.forward(x, n_samples=n_samp, seeds=seeds).detach()
    bank = torch.stack(log.rows)
    logits_avg = bnn.forward(x, n_samples=n_samp, avg_posterior=True).detach()
    # avg_posterior overwrote basenet's weights (model_bnn.py:215); harmless for the guide.

    grads = torch.stack([lossGradients.loss_gradient(net=bnn, image=x[i], label=y[i], n_samples=n_samp)
                         for i in range(n_img)])
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            loader = DataLoader(dataset=list(zip(x, y)), batch_size=2, shuffle=False)
            grads_np = lossGradients.loss_gradients(net=bnn, data_loader=loader, device="cpu",
                                                    filename="g", savedir="g/", n_samples=n_samp)
            loader = DataLoader(dataset=list(zip(x, y)), batch_size=2, shuffle=False)
            acc = bnn.evaluate(loader, "cpu", n_samples=n_samp)
        finally:
            os.chdir(cwd)

    # unseeded SVI attack: every forward call draws fresh weights (model_bnn.py:230-232)
    pyro.set_rng_seed(123)
    hyper = {"epsilon": 0.25}
    with _DrawLog(layout) as log:
        fg = torch.cat([adversarialAttacks.fgsm_attack(bnn, x[i:i + 1].clone(), y[i].argmax(-1).unsqueeze(0),
                                                       hyperparams=hyper, n_samples=n_samp)
                        for i in range(n_img)]).detach()
    fresh_bank = torch.stack(log.rows)

    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        arch=arch, input_shape=np.array(input_shape), hidden=hidden, n_classes=n_classes,
        loc=loc.numpy(), rho=rho.numpy(), x=x.numpy(), y=y.numpy(), bank=bank.numpy(),
        probs_seeded=probs_seeded.numpy(), logits_avg=logits_avg.numpy(),
        loss_gradient=grads.numpy(), loss_gradients_np=grads_np, evaluate_acc=np.float64(acc),
        fresh_bank=fresh_bank.numpy(), fgsm_fresh=fg.numpy(), fgsm_fresh_eps=np.float64(hyper["epsilon"]))
    print(name, "P =", bank.shape[1], "grad max", float(grads.abs().max()))


def hmc_case(name, arch, input_shape, hidden, n_classes, n_img, n_samp, dataset="mnist", eps=0.3):
    bnn = model_bnn.BNN(dataset, hidden, "leaky", arch, "hmc", None, None, n_samp, 5, input_shape, n_classes)
    bnn.device = "cpu"
    bnn.basenet.device = "cpu"
    layout = _layout_of(bnn)
    loc, rho = orc.scaled_guide_params(layout, seed=2, rho_mean=-3.0)
    g = torch.Generator().manual_seed(7)
    bank = loc + orc.softplus(rho) * torch.randn((n_samp, loc.numel()), generator=g)
    # BNN.load's HMC branch (model_bnn.py:184-190) with in-memory state dicts
    import copy
    bnn.posterior_predictive = {}
    for i in range(n_samp):
        net_copy = copy.deepcopy(bnn.basenet)
        net_copy.load_state_dict(orc.unpack(bank[i], layout))
        net_copy.device = "cpu"
        bnn.posterior_predictive.update({i: net_copy})
    x, y = orc.synthetic_inputs(n_img, input_shape, n_classes, seed=3)
    y = _half_predicted_labels(bnn, x, y, n_samp, None)

    probs = bnn.forward(x, n_samples=n_samp).detach()
    grads = torch.stack([lossGradients.loss_gradient(net=bnn, image=x[i], label=y[i], n_samples=n_samp)
                         for i in range(n_img)])
    hyper = {"epsilon": eps}
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for method in ("fgsm", "pgd"):
                for hname, h in (("hyper", hyper), ("default", None)):
                    adv = adversarialAttacks.attack(net=bnn, x_test=x.clone(), y_test=y, dataset_name=dataset,
                                                    device="cpu", method=method, filename="a", savedir="a",
                                                    hyperparams=h, n_samples=n_samp).detach()
                    o_acc, a_acc, rob = adversarialAttacks.attack_evaluation(
                        net=bnn, x_test=x, x_attack=adv, y_test=y, device="cpu", n_samples=n_samp)
                    out[f"{method}_{hname}_adv"] = adv.numpy()
                    out[f"{method}_{hname}_eval"] = np.array([o_acc, a_acc], dtype=np.float64)
                    out[f"{method}_{hname}_rob"] = rob.numpy()
        finally:
            os.chdir(cwd)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        arch=arch, input_shape=np.array(input_shape), hidden=hidden, n_classes=n_classes,
        x=x.numpy(), y=y.numpy(), bank=bank.numpy(), probs=probs.numpy(), loss_gradient=grads.numpy(),
        eps=np.float64(eps), **out)
    print(name, "P =", bank.shape[1], "grad max", float(grads.abs().max()),
          {k: v.tolist() for k, v in out.items() if k.endswith("_eval")})


if __name__ == "__main__":
    svi_case("svi_fc16_mnist", "fc", (1, 28, 28), 16, 10, n_img=6, n_samp=3)
    svi_case("svi_fc2_32_moons", "fc2", (1, 2, 1), 32, 2, n_img=6, n_samp=5, dataset="half_moons")
    svi_case("svi_conv16_mnist", "conv", (1, 28, 28), 16, 10, n_img=2, n_samp=2)
    hmc_case("hmc_fc16_fmnist", "fc", (1, 28, 28), 16, 10, n_img=9, n_samp=3, dataset="fashion_mnist", eps=0.05)
    hmc_case("hmc_fc2_16_mnist", "fc2", (1, 28, 28), 16, 10, n_img=6, n_samp=2, eps=0.02)
    hmc_case("hmc_conv16_fmnist", "conv", (1, 28, 28), 16, 10, n_img=3, n_samp=2, dataset="fashion_mnist", eps=0.1)
    hmc_case("hmc_fc2_32_moons", "fc2", (1, 2, 1), 32, 2, n_img=6, n_samp=4, dataset="half_moons", eps=0.2)
