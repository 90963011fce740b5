"""Host-side logic of the drop-in classes, exercised on CPU with the oracle-backed test double
(tests/helpers.py::OracleEngine) in place of the CUDA engine: sample placement, seeds rules,
fresh-draw bookkeeping, 2-rank gloo sharding + allreduce, result formats."""
import os
import pickle

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import oracle as orc
from tests.helpers import Case, OracleEngine, rel_err


def _bnn(case, inference="svi", n_samples=None):
    from robustbnns_b200.model_bnn import BNN
    eng = OracleEngine(case.arch, case.input_shape, case.hidden, case.n_classes, case.dataset)
    return BNN(case.dataset, case.hidden, "leaky", case.arch, inference, 1, 0.01, n_samples, 5,
               case.input_shape, case.n_classes, engine=eng)


def test_forward_seed_rules_and_explicit_bank():
    c = Case("svi_fc16_mnist")
    S = c.bank.shape[0]
    bnn = _bnn(c)
    bnn.set_guide(c.t("loc"), c.t("rho"))
    bnn.set_posterior_samples(c.bank)
    with pytest.raises(ValueError):
        bnn.forward(c.x, n_samples=2, seeds=[0, 1, 2])                # model_bnn.py:200-202
    out = bnn.forward(c.x, n_samples=S, seeds=list(range(S)))
    assert rel_err(out, c.t("probs_seeded")) < 1e-5
    out = bnn.forward(c.x, n_samples=2, seeds=[2, 0])                 # arbitrary seeds -> gathered rows
    ref = orc.bnn_forward(c.net, c.layout, c.bank, c.x, [2, 0])
    assert rel_err(out, ref) < 1e-6
    with pytest.raises(IndexError):
        bnn.forward(c.x, n_samples=S + 1)
    logits = bnn.forward(c.x, n_samples=S, avg_posterior=True)
    assert rel_err(logits, c.t("logits_avg")) < 1e-5


def test_svi_philox_prefix_and_fresh_draws():
    c = Case("svi_fc2_32_moons")
    bnn = _bnn(c)
    loc, rho = c.t("loc"), c.t("rho")
    bnn.set_guide(loc, rho)
    eng = bnn.engine()
    bnn.forward(c.x, n_samples=3, seeds=[0, 1, 2])
    bnn.forward(c.x, n_samples=5, seeds=list(range(5)))
    # seeds 0..4 each generated exactly once, global index == seed, key == rng_seed
    assert sorted(g for (k, g, _) in eng.sampled) == [0, 1, 2, 3, 4]
    assert all(k == bnn.rng_seed for (k, _, _) in eng.sampled)
    assert torch.equal(eng.bank[:5], orc.philox_bank(loc, rho, bnn.rng_seed, range(5)))
    n0 = len(eng.sampled)
    a = bnn.forward(c.x, n_samples=4)          # unseeded: fresh draws, different every call
    b = bnn.forward(c.x, n_samples=4)
    assert not torch.equal(a, b)
    fresh = eng.sampled[n0:]
    assert len(fresh) == 8 and len({g for (_, g, _) in fresh}) == 8
    assert all(g >= (1 << 31) for (_, g, _) in fresh)
    bnn.reseed(0)
    a2 = bnn.forward(c.x, n_samples=4)         # reseed(0) replays the stream (pyro.set_rng_seed(0))
    assert torch.equal(a, a2)


def test_unseeded_svi_attack_fresh_draws_per_image():
    """attack_all on an unseeded SVI BNN: one device pass shares the fresh draws of a gradient evaluation among its images;
    `fresh_draws = "per_image"` gives every image its own draws, as the reference's per-image loop does
    (adversarialAttacks.py:118-131 + model_bnn.py:230-232), and equals attacking the images one call at a time."""
    from robustbnns_b200 import adversarialAttacks as aa
    c = Case("svi_fc2_32_moons")
    S, N = 3, len(c.x)
    bnn = _bnn(c)
    bnn.set_guide(c.t("loc"), c.t("rho"))
    eng = bnn.engine()
    hyper = {"epsilon": 0.2}
    bnn.reseed(0)
    n0 = len(eng.sampled)
    aa.attack_all(bnn, c.x, c.labels, "fgsm", hyperparams=hyper, n_samples=S)
    assert len(eng.sampled) - n0 == S                                      # one evaluation, S draws, shared by the N images
    bnn.fresh_draws = "per_image"
    bnn.reseed(0)
    n0 = len(eng.sampled)
    adv = aa.attack_all(bnn, c.x, c.labels, "fgsm", hyperparams=hyper, n_samples=S)
    drawn = [g for (_, g, _) in eng.sampled[n0:]]
    assert len(drawn) == N * S and len(set(drawn)) == N * S                 # every image its own S draws
    bnn.reseed(0)
    one_by_one = torch.cat([aa.fgsm_attack(bnn, c.x[i:i + 1], c.labels[i:i + 1], hyperparams=hyper, n_samples=S)
                            for i in range(N)])
    assert torch.equal(adv, one_by_one)


def test_loss_gradients_and_attacks_through_the_drop_in(tmp_path, monkeypatch):
    from robustbnns_b200 import adversarialAttacks as aa
    from robustbnns_b200 import lossGradients as lg
    monkeypatch.chdir(tmp_path)
    c = Case("hmc_fc2_32_moons")
    S = c.bank.shape[0]
    bnn = _bnn(c, "hmc", S)
    bnn.set_posterior_samples(c.bank)
    g = torch.stack([lg.loss_gradient(bnn, c.x[i], c.y[i], n_samples=S) for i in range(len(c.x))])
    assert rel_err(g, c.t("loss_gradient")) < 1e-5
    loader = torch.utils.data.DataLoader(list(zip(c.x, c.y)), batch_size=4)
    out = lg.loss_gradients(bnn, loader, "cpu", "f", "f/", n_samples=S)
    assert isinstance(out, np.ndarray) and out.shape == (len(c.x), 2)             # squeezed (lossGradients.py:66)
    with open(os.path.join("data", "f", "f_samp=%d_lossGrads.pkl" % S), "rb") as f:
        assert np.array_equal(pickle.load(f), out)
    assert np.array_equal(lg.load_loss_gradients(S, "f", "f/"), out)
    with pytest.raises(NameError):
        lg.loss_gradient(bnn, c.x[0], c.y[0])                                     # broken upstream branch
    # autograd through forward: what the reference's fgsm does (adversarialAttacks.py:73-79)
    x = c.x.clone().requires_grad_(True)
    loss = torch.nn.CrossEntropyLoss(reduction="sum")(bnn.forward(x, n_samples=S), c.labels)
    loss.backward()
    ref = orc.attack_gradient(c.net, c.layout, c.bank, c.x, c.labels, range(S))
    assert rel_err(x.grad, ref) < 1e-5


def test_vanishing_norms_matches_reference_loop():
    from robustbnns_b200.lossGradients import compute_vanishing_norms_idxs
    rng = np.random.RandomState(0)
    g = rng.randn(40, 4, 5, 5).astype(np.float32) * np.array([1.0, 0.7, 0.5, 0.3], np.float32)[None, :, None, None]
    g[3] = 0.0
    for norm in ("linfty", "l2"):
        expect = []
        for i, ig in enumerate(g):                      # restatement of lossGradients.py:90-117
            nrm = (lambda a: np.max(np.abs(a))) if norm == "linfty" else np.linalg.norm
            cur = nrm(ig[0])
            if cur != 0.0:
                cnt = 0
                for j in range(4):
                    new = nrm(ig[j])
                    if new <= cur:
                        cur = new
                        cnt += 1
                if cnt == 4:
                    expect.append(i)
        assert compute_vanishing_norms_idxs(g, [1, 10, 50, 100], norm) == expect
    with pytest.raises(ValueError):
        compute_vanishing_norms_idxs(g, [1, 10], "l2")


def test_save_load_roundtrip_reference_formats(tmp_path):
    c = Case("svi_fc16_mnist")
    bnn = _bnn(c)
    bnn.set_guide(c.t("loc"), c.t("rho"))
    bnn.save(rel_path=str(tmp_path) + "/")
    state = torch.load(os.path.join(tmp_path, bnn.name, bnn.name + "_weights.pt"), weights_only=False)
    assert set(state.keys()) == {"params", "constraints"}                          # Pyro param-store layout
    assert "model.1.weight_loc" in state["params"] and "model.3.bias_scale" in state["params"]
    b2 = _bnn(c)
    b2.load("cpu", rel_path=str(tmp_path) + "/")
    assert torch.equal(b2._loc, bnn._loc) and torch.equal(b2._rho, bnn._rho)
    h = Case("hmc_fc16_fmnist")
    S = h.bank.shape[0]
    hb = _bnn(h, "hmc", S)
    hb.set_posterior_samples(h.bank)
    hb.save(rel_path=str(tmp_path) + "/")
    sd = torch.load(os.path.join(tmp_path, hb.name, hb.name + "_weights_1.pt"), weights_only=False)
    assert list(sd.keys()) == ["model.1.weight", "model.1.bias", "model.3.weight", "model.3.bias"]
    h2 = _bnn(h, "hmc", S)
    h2.load("cpu", rel_path=str(tmp_path) + "/")
    assert torch.equal(h2._bank_host, h.bank)
    h3 = _bnn(h, "hmc", S + 1)
    with pytest.raises(AttributeError):
        h3.load("cpu", rel_path=str(tmp_path) + "/", filename=hb.name + "_weights")


@pytest.mark.parametrize("name", ["ens_fc2_16_mnist", "ens_fc16_mnist"])
def test_nn_and_ensemble_host_logic(name, tmp_path, monkeypatch):
    """Deterministic NN / Ensemble_NN drop-ins (model_nn.py, model_ensemble.py) with the oracle standing in for the
    engine: weight files in the reference's layout, member selection, errors, attacks and evaluation against the
    reference's own outputs."""
    from robustbnns_b200 import adversarialAttacks as aa
    from robustbnns_b200.model_ensemble import Ensemble_NN
    from robustbnns_b200.model_nn import NN
    from tests.helpers import OracleEngine
    monkeypatch.chdir(tmp_path)
    c = Case(name)
    size, used = int(c.z["size"]), int(c.z["n_used"])
    ens = Ensemble_NN("mnist", c.hidden, "leaky", c.arch, 1, 0.01, c.input_shape, c.n_classes, size)
    assert ens.name == "mnist_ensemble_hid=%d_act=leaky_arch=%s_size=%d" % (c.hidden, c.arch, size)   # model_ensemble.py:26-31
    ens._engine = OracleEngine(c.arch, c.input_shape, c.hidden, c.n_classes)
    wdir = os.path.join("w", ens.name, "weights")
    os.makedirs(wdir)
    for i in range(size):
        torch.save(orc.unpack(c.bank[i], c.layout), os.path.join(wdir, "%s_weights_%d.pt" % (ens.member_name, i)))
    ens.load("cpu", rel_path="w/")
    with pytest.raises(ValueError):
        ens.forward(c.x, n_samples=size + 1)
    with pytest.raises(AttributeError):
        ens.set_members(c.bank[:size - 1])
    assert rel_err(ens.forward(c.x, n_samples=used), c.t("logits_used")) < 1e-5
    assert rel_err(ens.forward(c.x, None), c.t("logits_all")) < 1e-5
    nn0 = NN("mnist", c.input_shape, c.n_classes, c.hidden, "leaky", c.arch, 0.01, 1)
    nn0._engine = OracleEngine(c.arch, c.input_shape, c.hidden, c.n_classes)
    with pytest.raises(RuntimeError):
        nn0.forward(c.x)                              # no weights installed yet
    nn0.load("cpu", savedir=os.path.join(ens.name, "weights"), seed=0, rel_path="w/")
    assert rel_err(nn0.forward(c.x), c.t("logits_member0")) < 1e-5
    nn0.save(savedir="resaved", seed=3)               # the reference's file name under TESTS (model_nn.py:143-151)
    from robustbnns_b200.savedir import TESTS
    again = torch.load(os.path.join(TESTS, "resaved", nn0.name + "_weights_3.pt"), weights_only=False)
    assert list(again.keys()) == [k for k, _ in c.layout] and torch.equal(again[c.layout[0][0]], nn0.state_dict()[c.layout[0][0]])
    hyper = {"epsilon": float(c.z["eps"])}
    for who, net, ns in (("ens", ens, used), ("nn", nn0, None)):
        adv = aa.attack(net=net, x_test=c.x, y_test=c.y, dataset_name="mnist", device="cpu", method="fgsm",
                        filename="a", savedir="a", hyperparams=hyper, n_samples=ns)
        ref = c.t(f"{who}_fgsm_hyper_adv")
        assert float(((adv - ref).abs() > 1e-6).float().mean()) <= 2e-3
        adv = aa.attack(net=net, x_test=c.x, y_test=c.y, dataset_name="mnist", device="cpu", method="pgd",
                        filename="a", savedir="a", hyperparams=hyper, n_samples=ns)
        assert float(((adv - c.t(f"{who}_pgd_hyper_adv")).abs() > 1e-6).float().mean()) <= 2e-3   # per-image alpha (:89)
        o, a, rob = aa.attack_evaluation(net=net, x_test=c.x, x_attack=ref, y_test=c.y, device="cpu", n_samples=ns)
        assert [o, a] == c.z[f"{who}_fgsm_hyper_eval"].tolist()
        assert float((rob - c.t(f"{who}_fgsm_hyper_rob")).abs().max()) <= 1e-6


def test_half_moons_grid_driver(tmp_path, monkeypatch):
    """grid_search_halfMoons drop-in (BASELINE configs[4]) with the oracle standing in for the
    engine: the data set equals the reference's recipe, MoonsBNN names / weight files follow the reference, and the
    sweep's gradients and FGSM examples are the oracle's on the same stored posteriors."""
    import numpy as np
    from robustbnns_b200 import grid_search_halfMoons as gs
    from robustbnns_b200.adversarialAttacks import load_attack
    from robustbnns_b200.lossGradients import load_loss_gradients
    from robustbnns_b200.utils import load_half_moons
    from tests.helpers import OracleEngine
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr("robustbnns_b200.engine.Net",
                        lambda arch, shape, hidden, C: OracleEngine(arch, shape, hidden, C, dataset="half_moons"))
    x_train, y_train, x_test, y_test, shp, ncls = load_half_moons()
    assert x_train.shape == (24000, 1, 2, 1) and x_test.shape == (6000, 1, 2, 1) and tuple(shp) == (1, 2, 1) and ncls == 2
    assert float(min(x_train.min(), x_test.min())) == 0.0 and float(max(x_train.max(), x_test.max())) == 1.0
    assert y_test.shape == (6000, 2) and set(np.unique(y_test)) == {0.0, 1.0}
    S, pts = 4, 12
    banks = {}
    for hidden in (16, 32):
        bnn = gs.MoonsBNN(hidden, "leaky", "fc2", "hmc", None, None, S, 5, 100, (1, 2, 1), 2)
        assert bnn.name == "half_moons_bnn_hmc_hid=%d_act=leaky_arch=fc2_inp=100_samp=%d_warm=5_stepsize=0.001_numsteps=10" % (hidden, S)
        net = orc.build_net("fc2", (1, 2, 1), hidden, 2, dataset_name="half_moons")
        layout = orc.param_layout(net)
        loc, rho = orc.scaled_guide_params(layout, seed=hidden, rho_mean=-2.0)
        bank = loc + orc.softplus(rho) * torch.randn((S, loc.numel()), generator=torch.Generator().manual_seed(hidden))
        bnn.set_posterior_samples(bank)
        bnn.save(rel_path="w/")                                       # the reference's per-sample state-dict files
        banks[hidden] = (net, layout, bank, bnn.name)
    gs.serial_compute_grads([16, 32], ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S],
                            rel_path="w/", test_points=pts)
    gs.grid_attack("fgsm", [16, 32], ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S], pts,
                   device="cpu", rel_path="w/")
    xt, labels = torch.from_numpy(x_test[:pts]), torch.from_numpy(y_test[:pts]).argmax(-1)
    for hidden, (net, layout, bank, name) in banks.items():
        got = load_loss_gradients(n_samples=S, filename=name, savedir=name + "/")
        assert got.shape == (pts, 2)                                  # squeezed, lossGradients.py:66
        ref = orc.expected_loss_gradients(net, layout, bank, xt, labels, range(S)).reshape(pts, 2).numpy()
        order = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]           # the loader shuffles the test points  # noqa: E731
        assert np.abs(order(got) - order(ref)).max() <= 1e-5 * np.abs(ref).max()
        adv = load_attack("fgsm", name, n_samples=S)
        ref_adv = orc.fgsm_attack(net, layout, bank, xt, labels, lambda call: range(S), None)
        assert float((adv.cpu() - ref_adv).abs().max()) <= 1e-6
    # the batched form (upstream: joblib fan-out over CPU processes): same files, same numbers, one synchronisation
    serial = {name: (load_loss_gradients(n_samples=S, filename=name, savedir=name + "/"), load_attack("fgsm", name, n_samples=S))
              for (_, _, _, name) in banks.values()}
    import shutil
    shutil.rmtree("data")
    out = gs.parallel_compute_grads([16, 32], ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S],
                                    rel_path="w/", test_points=pts, device="cpu")
    advs = gs.parallel_grid_attack("fgsm", [16, 32], ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S],
                                   rel_path="w/", test_points=pts, device="cpu")
    assert len(out) == 2 and len(advs) == 2
    for name, (g_ser, a_ser) in serial.items():
        g_par = load_loss_gradients(n_samples=S, filename=name, savedir=name + "/")
        order = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]  # noqa: E731
        assert np.abs(order(g_par) - order(g_ser)).max() <= 1e-6 * np.abs(g_ser).max()
        assert torch.equal(load_attack("fgsm", name, n_samples=S).cpu(), a_ser.cpu())


def _grid_rank_main(rank, world, port, workdir, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import robustbnns_b200.engine as engine_mod
        from robustbnns_b200 import grid_search_halfMoons as gs
        torch.set_num_threads(1)
        os.chdir(workdir)
        engine_mod.Net = lambda arch, shape, hidden, C: OracleEngine(arch, shape, hidden, C, dataset="half_moons")
        S, pts = 3, 8
        out = gs.parallel_compute_grads([16, 32, 64], ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S],
                                        rel_path="w/", test_points=pts, device="cpu")
        advs = gs.parallel_grid_attack("fgsm", [16, 32, 64], ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100],
                                       [S], rel_path="w/", test_points=pts, device="cpu")
        q.put((rank, len(out), len(advs)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_model_sharded_grid(tmp_path, monkeypatch):
    """parallel_compute_grads / parallel_grid_attack under torch.distributed: model m of the grid goes to rank m % 2, no
    collective on the data path, every model's pickles are written by its owner and equal the single-process sweep."""
    import socket
    from robustbnns_b200 import grid_search_halfMoons as gs
    from robustbnns_b200.adversarialAttacks import load_attack
    from robustbnns_b200.lossGradients import load_loss_gradients
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr("robustbnns_b200.engine.Net",
                        lambda arch, shape, hidden, C: OracleEngine(arch, shape, hidden, C, dataset="half_moons"))
    S, pts, names = 3, 8, []
    for hidden in (16, 32, 64):
        bnn = gs.MoonsBNN(hidden, "leaky", "fc2", "hmc", None, None, S, 5, 100, (1, 2, 1), 2)
        net = orc.build_net("fc2", (1, 2, 1), hidden, 2, dataset_name="half_moons")
        loc, rho = orc.scaled_guide_params(orc.param_layout(net), seed=hidden, rho_mean=-2.0)
        bnn.set_posterior_samples(loc + orc.softplus(rho) * torch.randn((S, loc.numel()), generator=torch.Generator().manual_seed(hidden)))
        bnn.save(rel_path="w/")
        names.append(bnn.name)
    gs.parallel_compute_grads([16, 32, 64], ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S],
                              rel_path="w/", test_points=pts, device="cpu")
    gs.parallel_grid_attack("fgsm", [16, 32, 64], ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S],
                            rel_path="w/", test_points=pts, device="cpu")
    single = {n: (load_loss_gradients(n_samples=S, filename=n, savedir=n + "/"), load_attack("fgsm", n, n_samples=S)) for n in names}
    import shutil
    shutil.rmtree("data")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_grid_rank_main, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == [(0, 2, 2), (1, 1, 1)]                      # models 0, 2 -> rank 0; model 1 -> rank 1
    order = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]  # noqa: E731
    for n, (g, a) in single.items():
        g2 = load_loss_gradients(n_samples=S, filename=n, savedir=n + "/")
        assert np.abs(order(g2) - order(g)).max() <= 1e-6 * np.abs(g).max()
        assert torch.equal(load_attack("fgsm", n, n_samples=S).cpu(), a.cpu())


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from robustbnns_b200 import adversarialAttacks as aa
        from robustbnns_b200 import lossGradients as lg
        torch.set_num_threads(1)
        c = Case("hmc_fc16_fmnist")
        S = c.bank.shape[0]
        bnn = _bnn(c, "hmc", S)
        bnn.set_posterior_samples(c.bank)
        assert bnn.engine().bank.shape[0] >= 1 and bnn._pin_rows == len(range(rank, S, world))
        probs = bnn.forward(c.x, n_samples=S)
        grads = lg.expected_loss_gradients(bnn, c.x, c.labels, S)
        # row-sharded I/O: every rank uploads / reads back only its block of rows (9 images -> 5 + 4)
        blk, (lo, hi) = lg.expected_loss_gradients_block(bnn, c.x, c.labels, S)
        assert (lo, hi) == ((0, 5) if rank == 0 else (5, 9)) and blk.shape[0] == hi - lo
        assert rel_err(blk, grads[lo:hi]) < 1e-6
        adv = aa.fgsm_attack(bnn, c.x, c.labels, hyperparams={"epsilon": float(c.z["eps"])}, n_samples=S)
        s = Case("svi_fc2_32_moons")
        sb = _bnn(s)
        sb.set_guide(s.t("loc"), s.t("rho"))
        sp = sb.forward(s.x, n_samples=5, seeds=list(range(5)))
        fr = sb.forward(s.x, n_samples=3)
        pgd = aa.pgd_attack(bnn, c.x, c.labels, hyperparams=None, n_samples=S, iters=3)
        ev = aa.attack_evaluation(bnn, c.x, c.t("fgsm_hyper_adv"), c.y, "cpu", n_samples=S)
        # attack(): INPUT sharding (9 images -> blocks of 5 + 4, every rank holds all samples, no collective until the
        # final all-gather); afterwards the sample sharding is back in place
        import tempfile
        os.chdir(tempfile.mkdtemp())
        hyper = {"epsilon": float(c.z["eps"])}
        att = [aa.attack(net=bnn, x_test=c.x, y_test=c.y, dataset_name="mnist", device="cpu", method=m, filename="a",
                         savedir="a", hyperparams=hyper, n_samples=S) for m in ("fgsm", "pgd")]
        probs_after = bnn.forward(c.x, n_samples=S)
        assert torch.equal(probs_after, probs)
        bnn.attack_sharding = "samples"                                   # the all-reduce-per-iteration variant
        att_s = aa.attack(net=bnn, x_test=c.x, y_test=c.y, dataset_name="mnist", device="cpu", method="fgsm",
                          filename="a", savedir="a", hyperparams=hyper, n_samples=S)
        # numpy, not torch tensors: tensors travel through the queue as file descriptors served by THIS process, and
        # the parent may fetch them only after it has exited (FileNotFoundError in rebuild_storage_fd)
        npy = lambda t: t.detach().cpu().numpy()  # noqa: E731
        q.put((rank, npy(probs), npy(grads), npy(adv), npy(sp), npy(fr), npy(pgd), ev[:2], npy(ev[2]),
               [npy(a) for a in att], npy(att_s)))
        dist.barrier()                      # nobody tears the group down while a peer is still inside a collective
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tt = torch.from_numpy
    res = [(r[0], tt(r[1]), tt(r[2]), tt(r[3]), tt(r[4]), tt(r[5]), tt(r[6]), r[7], tt(r[8]), [tt(a) for a in r[9]], tt(r[10]))
           for r in res]
    c = Case("hmc_fc16_fmnist")
    S = c.bank.shape[0]
    sched = lambda call: range(S)  # noqa: E731
    pgd_ref = orc.pgd_attack(c.net, c.layout, c.bank, c.x, c.labels, sched, None, iters=3)
    for (_, probs, grads, adv, sp, fr, pgd, acc, rob, att, att_s) in res:
        assert att[0].shape == c.x.shape
        assert float((att[0] - c.t("fgsm_hyper_adv")).abs().max()) <= 1e-6
        assert float(((att[1] - c.t("pgd_hyper_adv")).abs() > 1e-6).float().mean()) <= 2e-3
        assert float((att_s - c.t("fgsm_hyper_adv")).abs().max()) <= 1e-6
        assert rel_err(probs, c.t("probs")) < 1e-5
        assert rel_err(grads, c.t("loss_gradient")) < 1e-5
        assert float((adv - c.t("fgsm_hyper_adv")).abs().max()) <= 1e-6
        assert float((pgd - pgd_ref).abs().max()) <= 1e-6
        assert list(acc) == c.z["fgsm_hyper_eval"].tolist()
        assert float((rob - c.t("fgsm_hyper_rob")).abs().max()) <= 1e-6
    # both ranks hold identical reduced results; SVI seeded/fresh draws do not depend on the sharding
    assert torch.equal(res[0][4], res[1][4]) and torch.equal(res[0][5], res[1][5])
    s = Case("svi_fc2_32_moons")
    sb = _bnn(s)
    sb.set_guide(s.t("loc"), s.t("rho"))
    assert rel_err(res[0][4], sb.forward(s.x, n_samples=5, seeds=list(range(5)))) < 1e-6
    assert rel_err(res[0][5], sb.forward(s.x, n_samples=3)) < 1e-6
