"""Data side of the reference's `plot_gradients_components.py`: the loss gradients for a list of sample counts and the
tables its two figures are drawn from (plot_gradients_components.py:17-38, :101-109, :125-142).  Drawing (seaborn /
matplotlib) is out of scope; the functions below hand back what the reference passes to `sns.stripplot` / `sns.heatmap`.
"""
import numpy as np
import pandas as pd

from .lossGradients import compute_vanishing_norms_idxs, load_loss_gradients, loss_gradients_list


def _get_gradients(args, bnn, test_loader, n_samples_list, relpath):
    """plot_gradients_components.py:125-142.  With `args.compute_grads` the whole list comes out of ONE pass over the
    posterior samples (lossGradients.loss_gradients_list) instead of one full evaluation per entry."""
    filename = bnn.name
    if args.compute_grads is True:
        return loss_gradients_list(net=bnn, data_loader=test_loader, device=args.device, filename=filename,
                                   savedir=filename + "/", n_samples_list=n_samples_list)
    return [load_loss_gradients(n_samples=n, filename=filename, relpath=relpath, savedir=filename + "/")
            for n in n_samples_list]


def gradients_components_df(loss_gradients_list, n_samples_list):
    """The long-format frame behind `stripplot_gradients_components` (plot_gradients_components.py:23-36): one row per
    gradient component, columns `loss_gradients` and `n_samples`, blocks in the order of `n_samples_list`."""
    if len(loss_gradients_list) != len(n_samples_list):
        raise ValueError("one gradient array per entry of n_samples_list")
    comps = [np.asarray(g).reshape(-1) for g in loss_gradients_list]
    for n, c in zip(n_samples_list, comps):
        print("\nsamples = ", n, end="\t")
        print(f"min = {c.min():.4f}", end="\t")
        print(f"max = {c.max():.4f}")
    return pd.DataFrame(data={"loss_gradients": np.concatenate(comps),
                              "n_samples": np.concatenate([np.full(len(c), n) for n, c in zip(n_samples_list, comps)])})


def vanishing_gradients_table(loss_gradients_list, n_samples_list, norm="linfty"):
    """What `vanishing_gradients_heatmaps` selects and draws (plot_gradients_components.py:101-117): the gradients as
    [image, sample count, h, w], the indices of the images whose norm does not increase along the list, and per selected
    image the norms written above its heatmaps.  Returns (transposed_gradients, vanishing_idxs, norms[idx] -> list)."""
    transposed = np.transpose(np.array(loss_gradients_list), axes=(1, 0, 2, 3))
    if transposed.shape[1] != len(n_samples_list):
        raise ValueError("Second dimension should contain the number of samples.")
    idxs = compute_vanishing_norms_idxs(loss_gradients=transposed, n_samples_list=n_samples_list, norm=norm)
    norms = {}
    for i in idxs:
        if norm == "linfty":
            norms[i] = [float(np.max(np.abs(g))) for g in transposed[i]]
        else:
            norms[i] = [float(np.linalg.norm(x=g, ord=2)) for g in transposed[i]]      # matrix 2-norm, as upstream (:79)
    return transposed, idxs, norms
