"""Golden vectors for the deterministic-network rows, produced by the reference's OWN code:

  NN.forward / NN.evaluate                     model_nn.py:126-141, 220-238
  Ensemble_NN.forward / evaluate               model_ensemble.py:57-67, 86-108
  fgsm_attack / pgd_attack / attack /
  attack_evaluation on those nets              adversarialAttacks.py:69-198

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_ensemble.py

The ensemble members are untrained `NN`s initialised under torch.manual_seed(seed) (training is out of scope); their
state dicts are stored as the rows of `bank`, member i = row i.
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

import adversarialAttacks  # noqa: E402  (reference)
import model_ensemble  # noqa: E402  (reference)
import model_nn  # noqa: E402  (reference)
from torch.utils.data import DataLoader  # noqa: E402

from oracle import oracle as orc  # noqa: E402  (synthetic inputs only)

torch.set_num_threads(4)


def ensemble_case(name, arch, input_shape, hidden, n_classes, n_img, size, n_used, dataset="mnist", eps=0.1, gain=4.0):
    ens = model_ensemble.Ensemble_NN(dataset, hidden, "leaky", arch, 1, 0.01, input_shape, n_classes, size)
    ens.device = "cpu"
    rows = []
    for seed in range(size):
        torch.manual_seed(100 + seed)
        net = model_nn.NN(dataset_name=dataset, input_shape=input_shape, output_size=n_classes, hidden_size=hidden,
                          activation="leaky", architecture=arch, epochs=1, lr=0.01)
        net.device = "cpu"
        with torch.no_grad():                      # default init gives near-uniform softmax rows: sharpen the last layer
            last = list(net.state_dict().keys())[-2:]
            for k in last:
                net.state_dict()[k].mul_(gain)
        ens.ensemble_models[str(seed)] = net
        rows.append(torch.cat([v.detach().reshape(-1) for v in net.state_dict().values()]))
    bank = torch.stack(rows)
    x, y = orc.synthetic_inputs(n_img, input_shape, n_classes, seed=11)
    pred = ens.forward(x, n_samples=n_used).detach().argmax(-1)
    y = y.clone()
    for i in range(n_img):                          # two thirds of the labels = the ensemble's own prediction
        if i % 3 != 2:
            y[i] = torch.nn.functional.one_hot(pred[i], n_classes).float()
    out = {"logits_used": ens.forward(x, n_samples=n_used).detach().numpy(),
           "logits_all": ens.forward(x, n_samples=None).detach().numpy(),
           "logits_member0": ens.ensemble_models["0"].forward(x).detach().numpy()}
    hyper = {"epsilon": eps}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            loader = DataLoader(dataset=list(zip(x, y)), batch_size=4, shuffle=False)
            out["evaluate_acc"] = np.float64(float(ens.evaluate(loader, "cpu", n_samples=n_used)))
            loader = DataLoader(dataset=list(zip(x, y)), batch_size=4, shuffle=False)
            out["evaluate_acc_member0"] = np.float64(float(ens.ensemble_models["0"].evaluate(loader, "cpu")))
            for who, net, ns in (("ens", ens, n_used), ("nn", ens.ensemble_models["0"], None)):
                for method in ("fgsm", "pgd"):
                    for hname, h in (("hyper", hyper), ("default", None)):
                        adv = adversarialAttacks.attack(net=net, x_test=x.clone(), y_test=y, dataset_name=dataset,
                                                        device="cpu", method=method, filename="a", savedir="a",
                                                        hyperparams=h, n_samples=ns).detach()
                        o_acc, a_acc, rob = adversarialAttacks.attack_evaluation(
                            net=net, x_test=x, x_attack=adv, y_test=y, device="cpu", n_samples=ns)
                        out[f"{who}_{method}_{hname}_adv"] = adv.numpy()
                        out[f"{who}_{method}_{hname}_eval"] = np.array([o_acc, a_acc], dtype=np.float64)
                        out[f"{who}_{method}_{hname}_rob"] = rob.detach().numpy()
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), arch=arch, input_shape=np.array(input_shape), hidden=hidden,
                        n_classes=n_classes, size=size, n_used=n_used, x=x.numpy(), y=y.numpy(), bank=bank.numpy(),
                        eps=np.float64(eps), **out)
    print(name, "P =", bank.shape[1], {k: v.tolist() for k, v in out.items() if k.endswith("_eval")},
          out["evaluate_acc"], out["evaluate_acc_member0"])


if __name__ == "__main__":
    ensemble_case("ens_fc2_16_mnist", "fc2", (1, 28, 28), 16, 10, n_img=9, size=4, n_used=3)
    ensemble_case("ens_conv16_fmnist", "conv", (1, 28, 28), 16, 10, n_img=4, size=3, n_used=2, dataset="fashion_mnist")
    ensemble_case("ens_fc16_mnist", "fc", (1, 28, 28), 16, 10, n_img=7, size=2, n_used=2)
