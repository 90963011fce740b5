"""Golden fixtures for the rows either side of the hot path (checkpoint ingestion, result formats), produced
by the reference's OWN code.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_tables.py

As in make_golden.py the reference modules are imported unmodified with `oracle/pyro_shim` ahead of them.  Stored:

  tables_hmc_fc16_fmnist.npz
      loss_gradients_<n>   lossGradients.loss_gradients for n in [1, 2, 3] on the bank of hmc_fc16_fmnist.npz (:52-68)
      vanishing_linfty/l2  lossGradients.compute_vanishing_norms_idxs on those arrays (:78-127)
      eps_csv              the CSV plot_eps_attacks.build_eps_attacks_df writes (:9-39), FGSM, eps [0.05, 0.1], samples [1, 3]
  svi_fc16_mnist_weights.pt
      the file the reference's BNN.save (model_bnn.py:138-155) writes for the guide of svi_fc16_mnist.npz
"""
import copy
import io
import contextlib
import os
import shutil
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
import plot_eps_attacks  # noqa: E402  (reference)
from torch.utils.data import DataLoader  # noqa: E402

from oracle import oracle as orc  # noqa: E402

torch.set_num_threads(4)


def _layout_of(bnn):
    return [(k, tuple(v.shape)) for k, v in bnn.basenet.state_dict().items()]


def hmc_tables(name="hmc_fc16_fmnist"):
    z = np.load(os.path.join(HERE, name + ".npz"))
    arch, shape, hidden, ncls = str(z["arch"]), tuple(int(v) for v in z["input_shape"]), int(z["hidden"]), int(z["n_classes"])
    bank, x, y = torch.from_numpy(z["bank"]), torch.from_numpy(z["x"]), torch.from_numpy(z["y"])
    S = bank.shape[0]
    bnn = model_bnn.BNN("fashion_mnist", hidden, "leaky", arch, "hmc", None, None, S, 5, shape, ncls)
    bnn.device = "cpu"
    bnn.basenet.device = "cpu"
    layout = _layout_of(bnn)
    bnn.posterior_predictive = {}
    for i in range(S):
        net_copy = copy.deepcopy(bnn.basenet)
        net_copy.load_state_dict(orc.unpack(bank[i], layout))
        net_copy.device = "cpu"
        bnn.posterior_predictive.update({i: net_copy})
    out = {}
    n_list = [1, 2, 3]
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            grads = []
            for n in n_list:
                loader = DataLoader(dataset=list(zip(x, y)), batch_size=4, shuffle=False)
                g = lossGradients.loss_gradients(net=bnn, data_loader=loader, device="cpu", filename="g", savedir="g/",
                                                 n_samples=n)
                out["loss_gradients_%d" % n] = g
                grads.append(g)
            tr = np.transpose(np.array(grads), axes=(1, 0, 2, 3))
            for norm in ("linfty", "l2"):
                with contextlib.redirect_stdout(io.StringIO()):
                    idxs = lossGradients.compute_vanishing_norms_idxs(loss_gradients=tr, n_samples_list=n_list, norm=norm)
                out["vanishing_" + norm] = np.array(idxs, dtype=np.int64)
            df = plot_eps_attacks.build_eps_attacks_df(bnn, "fashion_mnist", "cpu", "fgsm", x.clone(), y, [0.05, 0.1], [1, 3],
                                                       "eps_tables")
            csvs = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp) for f in fs if f.endswith("_increasing_eps_fgsm.csv")]
            assert len(csvs) == 1, csvs
            out["eps_csv"] = np.array(open(csvs[0]).read())
            assert len(df) == 2 * 2 * len(x)
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "tables_" + name + ".npz"), n_list=np.array(n_list), **out)
    print("tables_" + name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def svi_checkpoint(name="svi_fc16_mnist"):
    z = np.load(os.path.join(HERE, name + ".npz"))
    arch, shape, hidden, ncls = str(z["arch"]), tuple(int(v) for v in z["input_shape"]), int(z["hidden"]), int(z["n_classes"])
    bnn = model_bnn.BNN("mnist", hidden, "leaky", arch, "svi", 1, 0.01, None, None, shape, ncls)
    layout = _layout_of(bnn)
    loc, rho = torch.from_numpy(z["loc"]), torch.from_numpy(z["rho"])
    pyro.clear_param_store()
    off = 0
    for key, shp in layout:
        n = int(np.prod(shp))
        pyro.param(f"{key}_loc", loc[off:off + n].reshape(shp).clone())
        pyro.param(f"{key}_scale", rho[off:off + n].reshape(shp).clone())
        off += n
    with tempfile.TemporaryDirectory() as tmp:
        bnn.save(rel_path=tmp + "/")                      # the reference's BNN.save, SVI branch
        src = os.path.join(tmp, bnn.name, bnn.name + "_weights.pt")
        shutil.copy(src, os.path.join(HERE, name + "_weights.pt"))
    print(name + "_weights.pt", "written by", model_bnn.__file__, "as", bnn.name)


if __name__ == "__main__":
    hmc_tables()
    svi_checkpoint()
