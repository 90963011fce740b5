"""Rows either side of the hot path: checkpoint ingestion and result formats, against fixtures written by the
reference's own code (tests/golden/make_golden_tables.py).  The same body runs on CPU with the oracle-backed engine double
and -- marked gpu -- with the CUDA engine."""
import io
import os
import pickle
import shutil
import types

import numpy as np
import pandas
import pytest
import torch

from tests.helpers import GOLDEN, Case, OracleEngine, rel_err

REL = 1e-4


def _bnn(case, inference, n_samples, on_gpu):
    from robustbnns_b200.model_bnn import BNN
    eng = None if on_gpu else OracleEngine(case.arch, case.input_shape, case.hidden, case.n_classes, case.dataset)
    return BNN("fashion_mnist" if "fmnist" in case.name else case.dataset, case.hidden, "leaky", case.arch, inference,
               1 if inference == "svi" else None, 0.01 if inference == "svi" else None, n_samples, 5, case.input_shape,
               case.n_classes, engine=eng)


def _tables_body(tmp_path, monkeypatch, on_gpu):
    from robustbnns_b200 import lossGradients as lg
    from robustbnns_b200 import plot_eps_attacks as pea
    from robustbnns_b200 import plot_gradients_components as pgc
    monkeypatch.chdir(tmp_path)
    c = Case("hmc_fc16_fmnist")
    t = np.load(os.path.join(GOLDEN, "tables_hmc_fc16_fmnist.npz"))
    n_list = [int(n) for n in t["n_list"]]
    S = c.bank.shape[0]
    bnn = _bnn(c, "hmc", S, on_gpu)
    bnn.set_posterior_samples(c.bank)
    dev = "cuda" if on_gpu else "cpu"

    # one pass over the posterior samples -> every prefix mean (lossGradients.py:148-151 runs them one by one)
    loader = torch.utils.data.DataLoader(list(zip(c.x, c.y)), batch_size=4)
    grads = lg.loss_gradients_list(bnn, loader, dev, "g", "g/", n_list)
    for n, g in zip(n_list, grads):
        ref = t["loss_gradients_%d" % n]
        assert isinstance(g, np.ndarray) and g.shape == ref.shape
        assert rel_err(g, ref) < (REL if on_gpu else 1e-5), n
        with open(os.path.join("data", "g", "g_samp=%d_lossGrads.pkl" % n), "rb") as f:
            assert np.array_equal(pickle.load(f), g)                       # the reference's pickle per sample count
        loader = torch.utils.data.DataLoader(list(zip(c.x, c.y)), batch_size=4)
        one = lg.loss_gradients(bnn, loader, dev, "h", "h/", n_samples=n)   # the one-count call agrees with its prefix
        assert rel_err(g, one) < 1e-6
    # unsorted list with a repeat: answers come back in the caller's order
    mixed = lg.expected_loss_gradients_prefix(bnn, c.x, c.labels, [3, 1, 3])
    assert rel_err(mixed[1].cpu().squeeze(), grads[0]) < 1e-6 and torch.equal(mixed[0], mixed[2])
    with pytest.raises(ValueError):
        lg.expected_loss_gradients_prefix(bnn, c.x, c.labels, [0, 2])

    # vanishing-gradient selection on the reference's own arrays and on ours
    ref_list = [t["loss_gradients_%d" % n] for n in n_list]
    for norm in ("linfty", "l2"):
        tr, idxs, norms = pgc.vanishing_gradients_table(ref_list, n_list, norm=norm)
        assert idxs == [int(i) for i in t["vanishing_" + norm]]
        assert tr.shape == (len(c.x), len(n_list), 28, 28) and set(norms) == set(idxs)
        assert pgc.vanishing_gradients_table(grads, n_list, norm=norm)[1] == idxs
    with pytest.raises(ValueError):
        pgc.vanishing_gradients_table(ref_list, n_list[:-1])
    args = types.SimpleNamespace(compute_grads=False, device=dev)
    bnn.name = "g"                                                        # _get_gradients reads <name>/<name>_samp=..
    loaded = pgc._get_gradients(args, bnn, None, n_list, "data/")
    assert all(np.array_equal(a, b) for a, b in zip(loaded, grads))
    df = pgc.gradients_components_df(grads, n_list)
    assert list(df.columns) == ["loss_gradients", "n_samples"] and len(df) == sum(g.size for g in grads)
    assert np.array_equal(df["n_samples"].to_numpy()[:grads[0].size], np.full(grads[0].size, n_list[0]))
    assert np.array_equal(df["loss_gradients"].to_numpy()[-grads[-1].size:], grads[-1].reshape(-1))

    # the increasing-epsilon table (plot_eps_attacks.py:9-39) against the CSV the reference wrote
    bnn = _bnn(c, "hmc", S, on_gpu)
    bnn.set_posterior_samples(c.bank)
    got = pea.build_eps_attacks_df(bnn, "fashion_mnist", dev, "fgsm", c.x.clone(), c.y, [0.05, 0.1], [1, 3], "eps_tables")
    ref = pandas.read_csv(io.StringIO(str(t["eps_csv"])))
    assert list(got.columns) == list(ref.columns) and len(got) == len(ref)
    back = pea.load_eps_attacks_df("fashion_mnist", "fgsm", "eps_tables")
    for col in ("attack_method", "epsilon", "n_samples", "test_acc", "adv_acc"):
        assert list(back[col]) == list(ref[col]), col                      # accuracies are integer counts / N: exact
    assert np.abs(back["softmax_rob"].to_numpy() - ref["softmax_rob"].to_numpy()).max() < (1e-4 if on_gpu else 1e-5)


def _checkpoint_body(tmp_path, on_gpu):
    """BNN.load on the param-store file the REFERENCE's BNN.save wrote (model_bnn.py:148-155), then our save."""
    c = Case("svi_fc16_mnist")
    bnn = _bnn(c, "svi", None, on_gpu)
    d = os.path.join(str(tmp_path), bnn.name)
    os.makedirs(d)
    shutil.copy(os.path.join(GOLDEN, "svi_fc16_mnist_weights.pt"), os.path.join(d, bnn.name + "_weights.pt"))
    bnn.load("cuda" if on_gpu else "cpu", rel_path=str(tmp_path) + "/")
    assert torch.equal(bnn._loc.cpu(), c.t("loc")) and torch.equal(bnn._rho.cpu(), c.t("rho"))
    S = c.bank.shape[0]
    logits = bnn.forward(c.x, n_samples=S, avg_posterior=True)              # deterministic given the guide's locs
    assert rel_err(logits.cpu(), c.t("logits_avg")) < (REL if on_gpu else 1e-5)
    # and back: same container as the reference's file (keys, tensors, constraint objects)
    out = os.path.join(str(tmp_path), "out") + "/"
    bnn.save(rel_path=out)
    ours = torch.load(os.path.join(out, bnn.name, bnn.name + "_weights.pt"), weights_only=False)
    theirs = torch.load(os.path.join(GOLDEN, "svi_fc16_mnist_weights.pt"), weights_only=False)
    assert set(ours) == set(theirs) == {"params", "constraints"}
    assert list(ours["params"]) == list(theirs["params"])
    for k in theirs["params"]:
        assert torch.equal(ours["params"][k], theirs["params"][k].detach()), k
        assert type(ours["constraints"][k]) is type(theirs["constraints"][k]) is type(torch.distributions.constraints.real)


def test_prefix_means_and_result_tables(tmp_path, monkeypatch):
    _tables_body(tmp_path, monkeypatch, on_gpu=False)


def test_reference_written_svi_checkpoint(tmp_path):
    _checkpoint_body(tmp_path, on_gpu=False)


@pytest.mark.gpu
def test_prefix_means_and_result_tables_gpu(tmp_path, monkeypatch):
    _tables_body(tmp_path, monkeypatch, on_gpu=True)


@pytest.mark.gpu
def test_reference_written_svi_checkpoint_gpu(tmp_path):
    _checkpoint_body(tmp_path, on_gpu=True)


def test_vanishing_norms_against_reference_outputs():
    """compute_vanishing_norms_idxs on the reference's gradient arrays == what the reference's function returned."""
    from robustbnns_b200.lossGradients import compute_vanishing_norms_idxs
    t = np.load(os.path.join(GOLDEN, "tables_hmc_fc16_fmnist.npz"))
    n_list = [int(n) for n in t["n_list"]]
    tr = np.transpose(np.array([t["loss_gradients_%d" % n] for n in n_list]), axes=(1, 0, 2, 3))
    for norm in ("linfty", "l2"):
        assert compute_vanishing_norms_idxs(tr, n_list, norm) == [int(i) for i in t["vanishing_" + norm]]
