"""Data side of the reference's `plot_baseline_attacks.py` (BASELINE configs[2]): attack the deterministic NN, the BNN and
the ensemble of the same architecture and tabulate test / adversarial accuracy and pointwise softmax robustness
(plot_baseline_attacks.py:10-146).  Same columns, row order and CSV; the line plot is out of scope."""
import os

import numpy as np
import pandas
import torch
from torch.utils.data import DataLoader

from .adversarialAttacks import attack, attack_evaluation
from .model_bnn import BNN, saved_BNNs
from .model_ensemble import Ensemble_NN
from .model_nn import NN, saved_NNs
from .savedir import DATA, TESTS
from .utils import load_dataset

COLUMNS = ["attack_method", "epsilon", "test_acc", "adv_acc", "softmax_rob", "attack_samples", "defence_samples",
           "model_type"]


def _block(model_type, method, epsilon, test_acc, adv_acc, softmax_rob, attack_samples, defence_samples):
    rob = np.asarray(torch.as_tensor(softmax_rob).detach().cpu(), dtype=np.float64).reshape(-1)
    return pandas.DataFrame({"attack_method": method, "epsilon": epsilon, "test_acc": test_acc, "adv_acc": adv_acc,
                             "softmax_rob": rob, "attack_samples": attack_samples,
                             "defence_samples": np.full(len(rob), defence_samples, dtype=object),
                             "model_type": model_type}, columns=COLUMNS)


def build_baseline_attacks_df(args, data=None):
    """`args`: model_idx, n_inputs, attack_method, savedir, device, test (the reference's argparse namespace).
    `data` = (x_test, y_test, input_shape, output_size) replaces the reference's `load_dataset` call (MNIST and
    Fashion-MNIST come from keras downloads there, which this package does not do)."""
    rel_path = DATA if args.savedir == "DATA" else TESTS
    epsilon = 0.3
    blocks = []

    def dataset_of(name):
        if data is not None:
            return data
        _, _, x, y, shp, out = load_dataset(dataset_name=name, n_inputs=args.n_inputs)
        return x, y, shp, out

    # --- deterministic NN (plot_baseline_attacks.py:27-57)
    dataset, hid, activ, arch, ep, lr = saved_NNs["model_" + str(args.model_idx)].values()
    x_np, y_np, inp_shape, out_size = dataset_of(dataset)
    x_test = torch.as_tensor(np.asarray(x_np)[:args.n_inputs])
    y_test = torch.as_tensor(np.asarray(y_np)[:args.n_inputs])
    nn = NN(dataset_name=dataset, input_shape=inp_shape, output_size=out_size, hidden_size=hid, activation=activ,
            architecture=arch, epochs=ep, lr=lr)
    nn.load(device=args.device, rel_path=rel_path)
    if args.test:
        nn.evaluate(test_loader=DataLoader(dataset=list(zip(x_test, y_test))), device=args.device)
    nn_attack = attack(net=nn, x_test=x_test, y_test=y_test, dataset_name=dataset, device=args.device,
                       method=args.attack_method, filename=nn.name)
    blocks.append(_block("nn", args.attack_method, epsilon,
                         *attack_evaluation(net=nn, x_test=x_test, x_attack=nn_attack, y_test=y_test, device=args.device),
                         1, None))

    # --- BNN: one attack per attack-sample count, evaluated under several defence-sample counts (:59-89)
    dataset_name, model = saved_BNNs["model_" + str(args.model_idx)]
    bnn = BNN(dataset_name, *list(model.values()), inp_shape, out_size)
    bnn.load(device=args.device, rel_path=rel_path)
    if args.test:
        bnn.evaluate(test_loader=DataLoader(dataset=list(zip(x_test, y_test))), device=args.device, n_samples=10)
    for attack_samples in [1]:
        bnn_attack = attack(net=bnn, x_test=x_test, y_test=y_test, dataset_name=dataset, device=args.device,
                            method=args.attack_method, filename=bnn.name, n_samples=attack_samples)
        for defence_samples in [1, 50, 100]:
            blocks.append(_block("bnn", args.attack_method, epsilon,
                                 *attack_evaluation(net=bnn, x_test=x_test, x_attack=bnn_attack, y_test=y_test,
                                                    device=args.device, n_samples=defence_samples),
                                 attack_samples, defence_samples))

    # --- ensemble of deterministic NNs (:91-126)
    ens = Ensemble_NN(dataset_name=dataset, input_shape=inp_shape, output_size=out_size, hidden_size=hid,
                      activation=activ, architecture=arch, epochs=ep, lr=lr, ensemble_size=100)
    ens.load(device=args.device, rel_path=rel_path)
    for n_samples in [1, 50, 100]:
        if args.test:
            ens.evaluate(test_loader=DataLoader(dataset=list(zip(x_test, y_test))), device=args.device, n_samples=n_samples)
        ens_attack = attack(net=ens, x_test=x_test, y_test=y_test, dataset_name=dataset, device=args.device,
                            method=args.attack_method, filename=ens.name, n_samples=n_samples)
        blocks.append(_block("ensemble", args.attack_method, epsilon,
                             *attack_evaluation(net=ens, x_test=x_test, x_attack=ens_attack, y_test=y_test,
                                                device=args.device, n_samples=n_samples),
                             n_samples, n_samples))

    df = pandas.concat(blocks, ignore_index=True)
    _save_baseline_attacks_df(df=df, dataset_name=dataset_name, attack_method=args.attack_method)
    return df


def _save_baseline_attacks_df(df, dataset_name, attack_method):
    print("\nSaving:", df)
    os.makedirs(os.path.dirname(TESTS + "/"), exist_ok=True)
    df.to_csv(TESTS + "/" + str(dataset_name) + "_baseline_attacks_" + str(attack_method) + ".csv", index=False, header=True)


def load_baseline_attacks_df(dataset_name, attack_method, savedir):
    """Reads TESTS/<savedir>/... while the writer above stores under TESTS/ -- upstream's own mismatch
    (plot_baseline_attacks.py:136-143), kept so that files written by either side are found where upstream looks."""
    df = pandas.read_csv(TESTS + savedir + "/" + str(dataset_name) + "_baseline_attacks_" + str(attack_method) + ".csv")
    print(df.head(300))
    return df
