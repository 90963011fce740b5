"""Data side of the reference's `plot_halfMoons_overparam.py` (BASELINE configs[4]): for every model of the half-moons
width / depth grid and every posterior-sample count, the test accuracy and the two loss-gradient components of each test
point (plot_halfMoons_overparam.py:34-78).  Same columns, same CSV; the scatter plots themselves are out of scope."""
import itertools
import os

import numpy as np
import pandas
from torch.utils.data import DataLoader

from .grid_search_halfMoons import MoonsBNN
from .lossGradients import load_loss_gradients
from .savedir import TESTS
from .utils import load_dataset

COLUMNS = ["hidden_size", "activation", "architecture", "inference", "epochs", "lr", "n_samples", "warmup", "n_inputs",
           "posterior_samples", "test_acc", "x", "y", "loss_gradients_x", "loss_gradients_y"]


def build_overparam_scatterplot_dataset(hidden_size, activation, architecture, inference, epochs, lr, n_samples, warmup,
                                        n_inputs, posterior_samples, device, test_points, rel_path):
    _, _, x_test, y_test, inp_shape, out_size = load_dataset(dataset_name="half_moons", n_inputs=test_points,
                                                              channels="first")
    xy = np.asarray(x_test).reshape(len(x_test), -1)[:, :2]
    blocks = []
    for init in itertools.product(hidden_size, activation, architecture, inference, epochs, lr, n_samples, warmup, n_inputs):
        for n_post in posterior_samples:
            bnn = MoonsBNN(*init, inp_shape, out_size)
            bnn.load(device=device, rel_path=rel_path)
            test_loader = DataLoader(dataset=list(zip(x_test, y_test)), batch_size=64)
            test_acc = bnn.evaluate(test_loader=test_loader, device=device, n_samples=n_post)
            grads = load_loss_gradients(n_samples=n_post, filename=bnn.name, savedir=bnn.name + "/", relpath=rel_path)
            grads = np.asarray(grads[:test_points]).reshape(-1, 2)
            block = {c: v for c, v in zip(COLUMNS, init)}
            block.update({"posterior_samples": n_post, "test_acc": test_acc, "x": xy[:len(grads), 0], "y": xy[:len(grads), 1],
                          "loss_gradients_x": grads[:, 0], "loss_gradients_y": grads[:, 1]})
            blocks.append(pandas.DataFrame(block, columns=COLUMNS))
    df = pandas.concat(blocks, ignore_index=True) if blocks else pandas.DataFrame(columns=COLUMNS)
    print("\nSaving:", df.head())
    os.makedirs(os.path.dirname(TESTS), exist_ok=True)
    df.to_csv(TESTS + "halfMoons_lossGrads_final_" + str(test_points) + ".csv", index=False, header=True)
    return df
