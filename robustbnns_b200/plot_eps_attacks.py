"""Data side of the reference's `plot_eps_attacks.py` (BASELINE configs[3]): attack a BNN at increasing strength for a list
of sample counts and tabulate accuracy and pointwise softmax robustness (plot_eps_attacks.py:9-42).  One table row per
(epsilon, n_samples, test point), the reference's columns, the reference's CSV."""
import os

import numpy as np
import pandas

from .adversarialAttacks import attack, attack_evaluation
from .savedir import DATA

COLUMNS = ["attack_method", "epsilon", "test_acc", "adv_acc", "softmax_rob", "n_samples"]


def build_eps_attacks_df(bnn, dataset, device, method, x_test, y_test, epsilon_list, n_samples_list, savedir):
    blocks = []
    for epsilon in epsilon_list:
        for n_samples in n_samples_list:
            x_attack = attack(net=bnn, x_test=x_test, y_test=y_test, dataset_name=dataset, device=device, method=method,
                              filename=bnn.name, n_samples=n_samples, hyperparams={"epsilon": epsilon})
            test_acc, adv_acc, softmax_rob = attack_evaluation(net=bnn, x_test=x_test, n_samples=n_samples,
                                                               x_attack=x_attack, y_test=y_test, device=device)
            rob = np.asarray(softmax_rob.detach().cpu(), dtype=np.float64).reshape(-1)
            blocks.append(pandas.DataFrame({"attack_method": method, "epsilon": epsilon, "test_acc": test_acc,
                                            "adv_acc": adv_acc, "softmax_rob": rob, "n_samples": n_samples},
                                           columns=COLUMNS))
    df = pandas.concat(blocks, ignore_index=True) if blocks else pandas.DataFrame(columns=COLUMNS)
    print("\nSaving:", df)
    os.makedirs(os.path.dirname(DATA + savedir + "/"), exist_ok=True)
    df.to_csv(DATA + savedir + "/" + str(dataset) + "_increasing_eps_" + str(method) + ".csv", index=False, header=True)
    return df


def load_eps_attacks_df(dataset, method, savedir):
    return pandas.read_csv(DATA + savedir + "/" + str(dataset) + "_increasing_eps_" + str(method) + ".csv")
