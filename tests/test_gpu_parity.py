"""Parity of the CUDA path (through the C ABI / the drop-in classes) against the oracle and the
golden vectors the reference's own code produced.  Needs a B200: run with `-m gpu`.

Tolerances (BASELINE.json north_star): expected gradients / probabilities rel <= 1e-4 in the
max-norm relative to max|ref| (fp32 oracle, TF32 off; fp64 oracle as tie-breaker); adversarial
examples equal within 1e-6 except measure-zero sign ties; accuracy counts bit-exact."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.helpers import ENS_CASES, HMC_CASES, SVI_CASES, Case, rel_err

pytestmark = pytest.mark.gpu
REL = 1e-4

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def _bnn(case, inference="svi", n_samples=None, precision="fp32"):
    """The golden-vector tests pin the FP32 CUDA-core engine (reference-class rounding); the drop-in's own default is
    'auto' (tensor-core engines where the network has one), exercised by test_default_engine_is_the_fastest_parity_grade."""
    from robustbnns_b200.model_bnn import BNN
    bnn = BNN(case.dataset, case.hidden, "leaky", case.arch, inference, 1, 0.01, n_samples, 5,
              case.input_shape, case.n_classes)
    bnn.set_precision(precision)
    return bnn


def _assert_adv_equal_where_determined(adv, ref_adv, g64, tol=REL, what=""):
    """Adversarial examples equal within 1e-6 wherever the sign of the attack gradient is numerically determined, i.e.
    |g| above the gradient tolerance (tol * max|g|, the north-star bound on the gradient error); below it fp32 rounding
    decides the sign in the reference itself.  `g64`: fp64 oracle gradient at the attacked point.  Returns the number of
    undetermined pixels that differ (informational)."""
    adv, ref_adv, g64 = adv.detach().cpu().double(), ref_adv.detach().cpu().double(), g64.detach().cpu().double()
    determined = g64.abs() > tol * g64.abs().max()
    diff = (adv - ref_adv).abs() > 1e-6
    bad = diff & determined.reshape(diff.shape)
    assert int(bad.sum()) == 0, (what, int(bad.sum()), int(diff.sum()), int(determined.sum()))
    return int(diff.sum())


def _teacher_forced_pgd(x0, grad_fn, cuda_step, alpha, eps, ts=(0, 1, 2, 3, 5, 8, 13, 21, 30, 39), what=""):
    """Multi-step PGD is chaotic in floating point (the oracle's own fp32 and fp64 trajectories of 40 steps end up
    differing on most pixels of the conv case), so parity is checked step by step along the ORACLE's fp32 trajectory
    x_0 .. x_40: at x_t (a) the CUDA attack gradient matches the fp64 oracle to the tolerance that holds at that state
    and (b) one CUDA step lands exactly on the oracle's x_{t+1} wherever the sign of the gradient is determined.
    grad_fn(x, dtype) -> oracle gradient; cuda_step(x_t) -> (gradient [like x], next image), both from the CUDA path;
    alpha: float or [B] tensor."""
    al = alpha if not torch.is_tensor(alpha) else alpha.reshape(-1, *([1] * (x0.dim() - 1)))
    traj = [x0.clone()]
    for t in range(max(ts) + 1):
        traj.append(orc.pgd_step(traj[-1], x0, grad_fn(traj[-1], torch.float32), al, eps).detach())
    for t in ts:
        g64 = grad_fn(traj[t], torch.float64)
        # late PGD states saturate the softmax: the reference's own fp32 gradient then carries a cancellation error above
        # 1e-4 (fp32 oracle vs fp64 oracle); parity is held to 3x that error there
        tol_t = max(REL, 3 * rel_err(grad_fn(traj[t], torch.float32), g64))
        g, nxt = cuda_step(traj[t])
        assert rel_err(g.cpu().reshape(g64.shape), g64) < tol_t, (what, t, rel_err(g.cpu().reshape(g64.shape), g64), tol_t)
        ref_nxt = orc.pgd_step(traj[t].double(), x0.double(), g64, al.double() if torch.is_tensor(al) else al, eps)
        _assert_adv_equal_where_determined(nxt, ref_nxt, g64, tol_t, (what, t))


# ------------------------------------------------------------------ golden vectors (reference code) ----
@pytest.mark.parametrize("name", SVI_CASES)
def test_golden_svi(name, tmp_path, monkeypatch):
    from robustbnns_b200 import lossGradients as lg
    monkeypatch.chdir(tmp_path)
    c = Case(name)
    S = c.bank.shape[0]
    bnn = _bnn(c)
    bnn.set_guide(c.t("loc"), c.t("rho"))
    bnn.set_posterior_samples(c.bank)          # seed i == row i: the draws the reference made
    out = bnn.forward(c.x, n_samples=S, seeds=list(range(S)))
    assert out.is_cuda and out.shape == (len(c.x), c.n_classes)
    assert rel_err(out.cpu(), c.t("probs_seeded")) < REL
    assert rel_err(bnn.forward(c.x, n_samples=S, avg_posterior=True).cpu(), c.t("logits_avg")) < REL
    g = torch.stack([lg.loss_gradient(bnn, c.x[i], c.y[i], n_samples=S) for i in range(len(c.x))])
    assert rel_err(g.cpu(), c.t("loss_gradient")) < REL
    loader = torch.utils.data.DataLoader(list(zip(c.x, c.y)), batch_size=2)
    arr = lg.loss_gradients(bnn, loader, "cuda", "f", "f/", n_samples=S)
    assert arr.shape == c.z["loss_gradients_np"].shape and rel_err(arr, c.z["loss_gradients_np"]) < REL
    loader = torch.utils.data.DataLoader(list(zip(c.x, c.y)), batch_size=2)
    assert bnn.evaluate(loader, "cuda", n_samples=S) == float(c.z["evaluate_acc"])
    # unseeded SVI attack replayed on the draws the reference consumed, image by image
    from robustbnns_b200 import adversarialAttacks as aa
    fresh = c.t("fresh_bank")
    hyper = {"epsilon": float(c.z["fgsm_fresh_eps"])}
    for i in range(len(c.x)):
        bnn.set_posterior_samples(fresh[i * S:(i + 1) * S])
        adv = aa.fgsm_attack(bnn, c.x[i:i + 1], c.labels[i:i + 1], hyperparams=hyper, n_samples=S)
        g64 = orc.attack_gradient(c.net, c.layout, fresh[i * S:(i + 1) * S], c.x[i:i + 1], c.labels[i:i + 1], range(S),
                                  dtype=torch.float64)
        _assert_adv_equal_where_determined(adv, c.t("fgsm_fresh")[i:i + 1], g64, what=("fgsm_fresh", i))


@pytest.mark.parametrize("name", HMC_CASES)
def test_golden_hmc_attacks_and_evaluation(name, tmp_path, monkeypatch):
    from robustbnns_b200 import adversarialAttacks as aa
    from robustbnns_b200 import lossGradients as lg
    monkeypatch.chdir(tmp_path)
    c = Case(name)
    S = c.bank.shape[0]
    bnn = _bnn(c, "hmc", S)
    bnn.set_posterior_samples(c.bank)
    assert rel_err(bnn.forward(c.x, n_samples=S).cpu(), c.t("probs")) < REL
    g = lg.expected_loss_gradients(bnn, c.x, c.labels, S)
    assert rel_err(g.cpu(), c.t("loss_gradient")) < REL
    hyper = {"epsilon": float(c.z["eps"])}
    for method in ("fgsm", "pgd"):
        for hname, h in (("hyper", hyper), ("default", None)):
            adv = aa.attack(net=bnn, x_test=c.x, y_test=c.y, dataset_name=c.dataset, device="cuda",
                            method=method, filename="a", savedir="a", hyperparams=h, n_samples=S)
            ref = c.t(f"{method}_{hname}_adv")
            assert adv.is_cuda and adv.shape == ref.shape
            if method == "pgd":
                # multi-step PGD is chaotic in floating point: the end point is only required to stay inside the eps-ball,
                # parity is checked step by step along the reference trajectory (_teacher_forced_pgd below)
                assert float((adv.cpu() - c.x).abs().max()) <= (0.5 if h is None else hyper["epsilon"]) + 1e-6
            else:
                g64 = orc.attack_gradient(c.net, c.layout, c.bank, c.x, c.labels, range(S), dtype=torch.float64)
                _assert_adv_equal_where_determined(adv, ref, g64, what=(method, hname))
            o, a, rob = aa.attack_evaluation(net=bnn, x_test=c.x, x_attack=ref, y_test=c.y, device="cuda",
                                             n_samples=S)
            assert [o, a] == c.z[f"{method}_{hname}_eval"].tolist()          # counts bit-exact
            assert float((rob.cpu() - c.t(f"{method}_{hname}_rob")).abs().max()) <= 1e-6
            loaded = aa.load_attack(method, "a", savedir="a", n_samples=S)
            assert torch.equal(loaded.cpu(), adv.cpu())
    # PGD, teacher-forced, for both step-size rules (adversarialAttacks.py:88-91)
    from robustbnns_b200 import _lib
    x0, y = aa._prep(bnn, c.x, c.labels)
    eng = bnn.engine()

    def grad_fn(x, dtype):
        return orc.attack_gradient(c.net, c.layout, c.bank, x, c.labels, range(S), dtype=dtype)

    for hname, alpha, eps in (("default", torch.full((len(c.x),), 2 / 225), 0.5),
                              ("hyper", 2 / c.x.flatten(1).max(dim=1)[0], hyper["epsilon"])):
        alpha_d = alpha.to(device=x0.device, dtype=torch.float32)

        def cuda_step(xt_cpu):
            xt = xt_cpu.cuda()
            pbar = eng.forward_probs_sum(xt, 0, S) / S
            g = eng.input_grad_sum(_lib.HEAD_GRAD_OF_MEAN, xt, y, 0, S, pbar=pbar) / S
            return g, aa._pgd_loop(bnn, xt, x0, y, alpha_d, eps, S, False, 1)

        _teacher_forced_pgd(c.x, grad_fn, cuda_step, alpha, eps, what=(name, hname))


# ------------------------------------------------------------------ deterministic NN / Ensemble_NN ----
@pytest.mark.parametrize("name", ENS_CASES)
def test_golden_ensemble_and_nn(name, tmp_path, monkeypatch):
    """the reference's Ensemble_NN / NN outputs (mean logits, accuracies, FGSM / PGD examples,
    evaluation counts, robustness) reproduced by the engine-backed drop-ins, weights read from the reference's files."""
    from robustbnns_b200 import adversarialAttacks as aa
    from robustbnns_b200.model_ensemble import Ensemble_NN
    from robustbnns_b200.model_nn import NN
    monkeypatch.chdir(tmp_path)
    c = Case(name)
    size, used = int(c.z["size"]), int(c.z["n_used"])
    ens = Ensemble_NN(c.dataset if "fmnist" not in name else "fashion_mnist", c.hidden, "leaky", c.arch, 1, 0.01,
                      c.input_shape, c.n_classes, size)
    # the members' weight files, as the reference's NN.save(savedir=<ens name>/weights, seed=i) writes them
    import os
    wdir = os.path.join("tests_root", ens.name, "weights")
    os.makedirs(wdir)
    for i in range(size):
        torch.save(orc.unpack(c.bank[i], c.layout), os.path.join(wdir, f"{ens.member_name}_weights_{i}.pt"))
    ens.load("cuda", rel_path="tests_root/")
    if c.arch == "conv":      # conv nets default to the tensor-core engine: same mean logits at the north-star tolerance ...
        assert ens.engine().precision == "f16x3"
        assert rel_err(ens.forward(c.x, n_samples=used).cpu(), c.t("logits_used")) < REL
        ens.engine().set_precision("fp32")          # ... the bit-level checks below pin the FP32 engine
    with pytest.raises(ValueError):
        ens.forward(c.x, n_samples=size + 1)
    assert rel_err(ens.forward(c.x, n_samples=used).cpu(), c.t("logits_used")) < REL
    assert rel_err(ens.forward(c.x, n_samples=None).cpu(), c.t("logits_all")) < REL
    nn0 = NN(ens.dataset_name, c.input_shape, c.n_classes, c.hidden, "leaky", c.arch, 0.01, 1)
    nn0.load("cuda", savedir=os.path.join(ens.name, "weights"), seed=0, rel_path="tests_root/")
    nn0.engine().set_precision("fp32")
    assert rel_err(nn0.forward(c.x).cpu(), c.t("logits_member0")) < REL
    assert torch.equal(nn0.state_dict()[c.layout[0][0]], orc.unpack(c.bank[0], c.layout)[c.layout[0][0]])
    loader = torch.utils.data.DataLoader(list(zip(c.x, c.y)), batch_size=4)
    assert abs(ens.evaluate(loader, "cuda", n_samples=used) - float(c.z["evaluate_acc"])) < 1e-4
    loader = torch.utils.data.DataLoader(list(zip(c.x, c.y)), batch_size=4)
    assert abs(nn0.evaluate(loader, "cuda") - float(c.z["evaluate_acc_member0"])) < 1e-4
    # gradient against the fp64 oracle
    for net, members, ns in ((ens, range(used), used), (nn0, [0], None)):
        g = net.input_grad(c.x, c.labels, ns).cpu().reshape(c.x.shape)
        ref = orc.ensemble_attack_gradient(c.net, c.layout, c.bank, c.x, c.labels, members, dtype=torch.float64)
        assert rel_err(g, ref) < REL
    hyper = {"epsilon": float(c.z["eps"])}
    for who, net, ns in (("ens", ens, used), ("nn", nn0, None)):
        for method in ("fgsm", "pgd"):
            for hname, h in (("hyper", hyper), ("default", None)):
                adv = aa.attack(net=net, x_test=c.x, y_test=c.y, dataset_name=c.dataset, device="cuda", method=method,
                                filename="a", savedir="a", hyperparams=h, n_samples=ns)
                ref = c.t(f"{who}_{method}_{hname}_adv")
                assert adv.is_cuda and adv.shape == ref.shape
                members = range(used) if who == "ens" else [0]
                g64 = orc.ensemble_attack_gradient(c.net, c.layout, c.bank, c.x, c.labels, members, dtype=torch.float64)
                if method == "pgd":
                    # multi-step PGD is chaotic in floating point (see the BNN test): the end point must stay inside the
                    # eps-ball; the first step is held to the oracle wherever the gradient sign is determined
                    assert float((adv.cpu() - c.x).abs().max()) <= (0.5 if h is None else hyper["epsilon"]) + 1e-6
                    one = aa.pgd_attack(net, c.x, c.labels, hyperparams=h, n_samples=ns, iters=1)
                    ref1 = (orc.ensemble_pgd_attack(c.net, c.layout, c.bank, c.x, c.labels, members, hyperparams=h,
                                                    iters=1, dtype=torch.float64))
                    _assert_adv_equal_where_determined(one, ref1, g64, what=(who, method, hname))
                else:
                    _assert_adv_equal_where_determined(adv, ref, g64, what=(who, method, hname))
                o, a, rob = aa.attack_evaluation(net=net, x_test=c.x, x_attack=ref, y_test=c.y, device="cuda",
                                                 n_samples=ns)
                assert [o, a] == c.z[f"{who}_{method}_{hname}_eval"].tolist()
                assert float((rob.cpu() - c.t(f"{who}_{method}_{hname}_rob")).abs().max()) <= 2e-6


# ------------------------------------------------------------------ oracle at larger sizes -------------
def _problem(arch, input_shape, hidden, n_classes, B, S, dataset="mnist", seed=5):
    net = orc.build_net(arch, input_shape, hidden, n_classes, dataset_name=dataset)
    layout = orc.param_layout(net)
    loc, rho = orc.scaled_guide_params(layout, seed=seed, rho_mean=-4.0)
    g = torch.Generator().manual_seed(seed + 1)
    bank = loc + orc.softplus(rho) * torch.randn((S, loc.numel()), generator=g)
    x, y = orc.synthetic_inputs(B, input_shape, n_classes, seed=seed + 2)
    return net, layout, loc, rho, bank, x, y.argmax(-1)


CONFIGS = [
    ("fc", (1, 28, 28), 512, 10, 200, 6, "mnist"),        # headline architecture
    ("fc", (1, 28, 28), 64, 10, 7, 10, "mnist"),          # ragged batch
    ("fc2", (1, 28, 28), 128, 10, 130, 5, "mnist"),
    ("fc2", (1, 2, 1), 128, 2, 100, 100, "half_moons"),   # BASELINE config 1: 100 points x 100 samples
    ("fc2", (1, 2, 1), 512, 2, 100, 20, "half_moons"),
    ("conv", (1, 28, 28), 32, 10, 9, 3, "mnist"),
    ("conv", (1, 28, 28), 512, 10, 4, 2, "mnist"),        # the real model_idx=0 width
]


@pytest.mark.parametrize("arch,shape,hidden,C,B,S,ds", CONFIGS)
def test_engine_vs_oracle(arch, shape, hidden, C, B, S, ds):
    from robustbnns_b200 import _lib
    from robustbnns_b200.engine import Net
    net, layout, loc, rho, bank, x, labels = _problem(arch, shape, hidden, C, B, S, ds)
    eng = Net(arch, shape, hidden, C)
    assert eng.P == bank.shape[1]
    eng.upload(bank, 0)
    assert torch.equal(eng.download(0, S), bank)
    probs = eng.forward_probs_sum(x, 0, S).cpu() / S
    ref_p = orc.bnn_forward(net, layout, bank, x, range(S)).detach()
    assert rel_err(probs, ref_p) < REL
    assert rel_err(eng.forward_logits(x, 1).cpu(), orc.bnn_forward_avg_posterior(net, layout, bank[1], x).detach()) < REL
    g = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S).cpu().reshape(x.shape) / S
    ref64 = orc.expected_loss_gradients(net, layout, bank, x, labels, range(S), dtype=torch.float64)
    ref32 = orc.expected_loss_gradients(net, layout, bank, x, labels, range(S))
    assert min(rel_err(g, ref64), rel_err(g, ref32)) < REL
    pbar = eng.forward_probs_sum(x, 0, S) / S
    ga = eng.input_grad_sum(_lib.HEAD_GRAD_OF_MEAN, x, labels, 0, S, pbar=pbar).cpu().reshape(x.shape) / S
    ra = orc.attack_gradient(net, layout, bank, x, labels, range(S), dtype=torch.float64)
    assert rel_err(ga, ra) < REL
    # the FP32 CUDA-core engine has no kept-forward route: keep=True is a plain forward
    assert rel_err(eng.forward_probs_sum(x, 0, S, keep=True).cpu() / S, ref_p) < REL and not eng.keep_valid
    gl = eng.input_grad_sum(_lib.HEAD_LOGITS_CE, x, labels, S - 1, S).cpu().reshape(x.shape)
    rl = orc.attack_gradient_avg_posterior(net, layout, bank[S - 1], x, labels, dtype=torch.float64)
    assert rel_err(gl, rl) < REL
    # split over row ranges == whole (what sample sharding relies on); empty range == zeros
    ga_ = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S // 2)
    gb_ = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, S // 2, S)
    assert rel_err((ga_ + gb_).cpu().reshape(x.shape) / S, g) < 1e-5
    assert float(eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 1, 1).abs().max()) == 0.0
    # rows of the batch are independent: a sub-batch reproduces the same rows (the split of the
    # sample loop over CTAs depends on the batch size, so only up to fp32 summation order)
    gs = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x[:3], labels[:3], 0, S).cpu().reshape(x[:3].shape) / S
    assert rel_err(gs, g[:3]) < 1e-5
    eng.close()


TC_CONFIGS = [
    ("fc", (1, 28, 28), 512, 10, 300, 9),       # headline architecture; ragged M tile; 9 samples -> several slots
    ("fc", (1, 28, 28), 64, 10, 7, 10),         # tiny batch, narrow hidden layer
    ("fc", (1, 28, 28), 32, 10, 130, 3),
    ("fc2", (1, 28, 28), 128, 10, 130, 5),
    ("fc2", (1, 28, 28), 512, 10, 257, 6),
    ("fc2", (1, 2, 1), 512, 2, 100, 7),         # half moons (D = 2): first layer on the CUDA cores, H x H layer on tcgen05
    ("fc2", (1, 2, 1), 32, 2, 130, 4),
    ("fc2", (1, 5, 1), 64, 3, 50, 4),           # another D % 8 != 0 shape: the small-D kernels are not tied to D = 2
]


@pytest.mark.parametrize("route", ["fused", "unfused"])
@pytest.mark.parametrize("prec,tol", [("tf32x3", REL), ("f16x3", REL), ("bf16", 1.0)])
@pytest.mark.parametrize("arch,shape,hidden,C,B,S", TC_CONFIGS)
def test_tcgen05_engine_vs_oracle(arch, shape, hidden, C, B, S, prec, tol, route, monkeypatch):
    """The tensor-core engine (tcgen05 / TMA / TMEM) against the fp64 oracle: TF32x3 is parity grade
    (rel <= 1e-4, the north-star tolerance); single-pass BF16 is the throughput mode and is only held
    to a loose bound (its measured deviation is reported by bench.py)."""
    from robustbnns_b200 import _lib
    from robustbnns_b200.engine import Net
    if prec == "f16x3" and (arch != "fc" or route == "unfused"):
        pytest.skip("F16X3 rides on the fused forward+head kernel (arch fc)")
    if prec == "bf16" and (shape[0] * shape[1] * shape[2]) % 8:
        pytest.skip("D % 8 != 0 (half moons): TF32X3 / FP32 only")
    if route == "unfused":
        if arch != "fc":
            pytest.skip("fc2 has a single (unfused) route")
        monkeypatch.setenv("RBNN_TC_UNFUSED", "1")      # arch fc: GEMM -> H in HBM -> head kernel instead of the fused kernel
    net, layout, loc, rho, bank, x, labels = _problem(arch, shape, hidden, C, B, S)
    eng = Net(arch, shape, hidden, C)
    eng.set_precision(prec)
    assert eng.precision == prec
    eng.upload(bank, 0)
    probs = eng.forward_probs_sum(x, 0, S).cpu() / S
    ref_p = orc.bnn_forward(net, layout, bank, x, range(S)).detach()
    assert rel_err(probs, ref_p) < tol
    assert rel_err(eng.forward_logits(x, 1).cpu(),
                   orc.bnn_forward_avg_posterior(net, layout, bank[1], x).detach()) < max(tol, 1e-4) * (0.25 if prec == "bf16" else 1)
    g = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S).cpu().reshape(x.shape) / S
    ref64 = orc.expected_loss_gradients(net, layout, bank, x, labels, range(S), dtype=torch.float64)
    e_mean = rel_err(g, ref64)
    pbar = eng.forward_probs_sum(x, 0, S) / S
    ga = eng.input_grad_sum(_lib.HEAD_GRAD_OF_MEAN, x, labels, 0, S, pbar=pbar).cpu().reshape(x.shape) / S
    ra = orc.attack_gradient(net, layout, bank, x, labels, range(S), dtype=torch.float64)
    e_att = rel_err(ga, ra)
    # two-phase form (what the attacks use): forward that keeps logits + masks, then the gradient from the kept data
    pk = eng.forward_probs_sum(x, 0, S, keep=True)
    assert eng.keep_valid          # fused route: logits + masks are kept; unfused / fc2: the hidden activations
    if eng.keep_valid:
        assert rel_err(pk, pbar * S) < 1e-6
        gk = eng.input_grad_sum_kept(_lib.HEAD_GRAD_OF_MEAN, labels, pbar=pk / S).cpu().reshape(x.shape) / S
        assert rel_err(gk, ga) < (1e-6 if prec != "bf16" else 1e-2)
        assert rel_err(gk, ra) < tol
        gm = eng.input_grad_sum_kept(_lib.HEAD_MEAN_OF_GRADS, labels).cpu().reshape(x.shape) / S
        assert rel_err(gm, g) < (1e-6 if prec != "bf16" else 1e-2)
        eng.upload(bank[0:1], 0)                   # touching a kept row invalidates the kept forward
        assert not eng.keep_valid
    gl = eng.input_grad_sum(_lib.HEAD_LOGITS_CE, x, labels, S - 1, S).cpu().reshape(x.shape)
    rl = orc.attack_gradient_avg_posterior(net, layout, bank[S - 1], x, labels, dtype=torch.float64)
    e_log = rel_err(gl, rl)
    # head of the ensembles / deterministic nets: g handed to the logits unchanged (its magnitude sets the F16X3 dH range)
    gu = torch.randn((B, C), generator=torch.Generator().manual_seed(3)) * 1e-3
    got_u = eng.input_grad_sum(_lib.HEAD_LOGITS_UPSTREAM, x, labels, 0, S, pbar=gu).cpu().reshape(x.shape)
    xs = x.double().clone().requires_grad_(True)
    lsum = sum(orc.net_logits(net, {k: v.double() for k, v in orc.unpack(bank[s], layout).items()}, xs) for s in range(S))
    (ref_u,) = torch.autograd.grad((lsum * gu.double()).sum(), xs)
    assert rel_err(got_u, ref_u) < tol
    cos = float(torch.nn.functional.cosine_similarity(g.double().flatten(), ref64.flatten(), dim=0))
    print(f"tcgen05 {prec} {route} {arch}-{hidden} B={B} S={S}: mean-of-grads {e_mean:.2e} grad-of-mean {e_att:.2e} "
          f"logits-CE {e_log:.2e} cosine {cos:.6f}")
    assert cos > 0.98
    # fc2 included: second-layer units whose pre-activation lies within the propagated first-layer tensor-core error of
    # zero are settled from exact first-layer values (refine2_kernel), so two hidden layers meet the same tolerance
    assert max(e_mean, e_att, e_log) < tol
    # the split of the sample range over calls (what sample sharding relies on) and a re-upload of one row
    ga_ = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S // 2)
    gb_ = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, S // 2, S)
    assert rel_err((ga_ + gb_).cpu().reshape(x.shape) / S, g) < max(1e-5, tol * 0.1)
    eng.upload(bank[0:1], 1)                       # row 1 <- row 0: derived tensor-core copies must follow
    g11 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 1, 2).cpu()
    g00 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, 1).cpu()
    assert torch.equal(g11, g00)
    # switching back to the CUDA-core engine on the same handle
    eng.set_precision("fp32")
    g32 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, 1).cpu()
    assert rel_err(g32, g00) < tol
    eng.close()


TC_CONV_CONFIGS = [
    ((1, 28, 28), 32, 10, 9, 3),        # odd batch: the last 128-row tile holds one image + out-of-bounds zeros
    ((1, 28, 28), 64, 10, 17, 4),
    ((1, 28, 28), 512, 10, 4, 2),       # the real model_idx=0 width (two 256-column tiles)
    ((1, 28, 28), 16, 10, 2, 5),        # narrowest legal hidden size: K = 16 < one K-block in the dgrad GEMM
    ((1, 28, 28), 1024, 10, 3, 2),      # saved_BNNs model_2 / model_4 / model_8 width (four 256-column tiles)
]


@pytest.mark.parametrize("prec", ["tf32x3", "f16x3"])
@pytest.mark.parametrize("shape,hidden,C,B,S", TC_CONV_CONFIGS)
def test_tcgen05_conv_engine_vs_oracle(shape, hidden, C, B, S, prec):
    """arch conv on the tensor cores (TF32X3 / F16X3): conv2 as an implicit GEMM over 5-D TMA boxes, its input gradient
    as a tcgen05 GEMM, LeakyReLU signs and pooling arg-maxes settled exactly inside the guard band -- against the fp64
    oracle at the north-star tolerance."""
    from robustbnns_b200 import _lib
    from robustbnns_b200.engine import Net
    net, layout, loc, rho, bank, x, labels = _problem("conv", shape, hidden, C, B, S)
    eng = Net("conv", shape, hidden, C)
    with pytest.raises(RuntimeError):
        eng.set_precision("bf16")
    eng.set_precision(prec)
    assert eng.precision == prec
    eng.upload(bank, 0)
    probs = eng.forward_probs_sum(x, 0, S).cpu() / S
    ref_p = orc.bnn_forward(net, layout, bank, x, range(S)).detach()
    assert rel_err(probs, ref_p) < REL
    assert rel_err(eng.forward_logits(x, 1).cpu(), orc.bnn_forward_avg_posterior(net, layout, bank[1], x).detach()) < REL
    g = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S).cpu().reshape(x.shape) / S
    ref64 = orc.expected_loss_gradients(net, layout, bank, x, labels, range(S), dtype=torch.float64)
    e_mean = rel_err(g, ref64)
    # heads whose g is handed in (autograd through forward, ensembles): the dZ2 range of F16X3 comes from max|g|
    gu = torch.randn((B, C), generator=torch.Generator().manual_seed(3)) * 1e-3
    got_u = eng.input_grad_sum(_lib.HEAD_LOGITS_UPSTREAM, x, labels, 0, S, pbar=gu).cpu().reshape(x.shape)
    xs = x.double().clone().requires_grad_(True)
    lsum = sum(orc.net_logits(net, {k: v.double() for k, v in orc.unpack(bank[s], layout).items()}, xs) for s in range(S))
    (ref_u,) = torch.autograd.grad((lsum * gu.double()).sum(), xs)
    assert rel_err(got_u, ref_u) < REL
    pbar = eng.forward_probs_sum(x, 0, S) / S
    ga = eng.input_grad_sum(_lib.HEAD_GRAD_OF_MEAN, x, labels, 0, S, pbar=pbar).cpu().reshape(x.shape) / S
    ra = orc.attack_gradient(net, layout, bank, x, labels, range(S), dtype=torch.float64)
    e_att = rel_err(ga, ra)
    gl = eng.input_grad_sum(_lib.HEAD_LOGITS_CE, x, labels, S - 1, S).cpu().reshape(x.shape)
    rl = orc.attack_gradient_avg_posterior(net, layout, bank[S - 1], x, labels, dtype=torch.float64)
    e_log = rel_err(gl, rl)
    print(f"tcgen05 {prec} conv-{hidden} B={B} S={S}: mean-of-grads {e_mean:.2e} grad-of-mean {e_att:.2e} logits-CE {e_log:.2e}")
    assert max(e_mean, e_att, e_log) < REL
    # two-phase form (what the attacks use): the forward keeps P1 / arg-max indices / refined A2 / logits, the gradient
    # pass starts from them -- same kernels on the same data, so the results are identical
    pk = eng.forward_probs_sum(x, 0, S, keep=True)
    assert eng.keep_valid and rel_err(pk, pbar * S) < 1e-6
    gk = eng.input_grad_sum_kept(_lib.HEAD_GRAD_OF_MEAN, labels, pbar=pk / S).cpu().reshape(x.shape) / S
    assert torch.equal(gk, ga)
    gm = eng.input_grad_sum_kept(_lib.HEAD_MEAN_OF_GRADS, labels).cpu().reshape(x.shape) / S
    assert torch.equal(gm, g)
    eng.upload(bank[0:1], 0)                       # touching a kept row invalidates the kept forward
    assert not eng.keep_valid
    ga_ = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S // 2)
    gb_ = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, S // 2, S)
    assert rel_err((ga_ + gb_).cpu().reshape(x.shape) / S, g) < 1e-5
    eng.upload(bank[0:1], 1)                       # row 1 <- row 0: the derived filter copies must follow
    g11 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 1, 2).cpu()
    g00 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, 1).cpu()
    assert torch.equal(g11, g00)
    eng.set_precision("fp32")                      # the CUDA-core engine on the same handle
    g32 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, 1).cpu()
    assert rel_err(g32, g00) < REL
    eng.close()


@pytest.mark.parametrize("prec", ["tf32x3", "f16x3"])
def test_golden_hmc_conv_on_tcgen05(prec, tmp_path, monkeypatch):
    """The reference's own outputs for the conv BNN (golden vectors) reproduced by the tensor-core conv engine:
    probabilities, expected loss gradients, FGSM examples, evaluation counts bit-exact."""
    from robustbnns_b200 import adversarialAttacks as aa
    from robustbnns_b200 import lossGradients as lg
    monkeypatch.chdir(tmp_path)
    c = Case("hmc_conv16_fmnist")
    S = c.bank.shape[0]
    bnn = _bnn(c, "hmc", S)
    bnn.set_posterior_samples(c.bank)
    bnn.set_precision(prec)
    assert rel_err(bnn.forward(c.x, n_samples=S).cpu(), c.t("probs")) < REL
    assert rel_err(lg.expected_loss_gradients(bnn, c.x, c.labels, S).cpu(), c.t("loss_gradient")) < REL
    hyper = {"epsilon": float(c.z["eps"])}
    adv = aa.attack(net=bnn, x_test=c.x, y_test=c.y, dataset_name=c.dataset, device="cuda", method="fgsm",
                    filename="a", savedir="a", hyperparams=hyper, n_samples=S)
    ref = c.t("fgsm_hyper_adv")
    g64 = orc.attack_gradient(c.net, c.layout, c.bank, c.x, c.labels, range(S), dtype=torch.float64)
    _assert_adv_equal_where_determined(adv, ref, g64, what=("conv fgsm", prec))
    o, a, rob = aa.attack_evaluation(net=bnn, x_test=c.x, x_attack=ref, y_test=c.y, device="cuda", n_samples=S)
    assert [o, a] == c.z["fgsm_hyper_eval"].tolist()
    assert float((rob.cpu() - c.t("fgsm_hyper_rob")).abs().max()) <= REL      # probabilities from tensor-core logits


@pytest.mark.parametrize("prec", ["tf32x3", "f16x3"])
@pytest.mark.parametrize("gain", [10.0, 30.0])
def test_tcgen05_engine_on_saturated_softmax(prec, gain):
    """Confident networks (output layer scaled up: softmax rows within 1e-6 .. 1e-30 of one-hot) make dlogits and dH
    tiny on most rows.  F16X3 keeps its operands in the 5-bit exponent range of fp16 through power-of-two scaling: the
    expected gradient must still match the fp64 oracle to the north-star tolerance (max-norm over the batch)."""
    from robustbnns_b200 import _lib
    from robustbnns_b200.engine import Net
    arch, shape, hidden, C, B, S = "fc", (1, 28, 28), 128, 10, 160, 6
    net, layout, loc, rho, bank, x, labels = _problem(arch, shape, hidden, C, B, S)
    off = 0
    for key, shp in layout:
        n = int(np.prod(shp))
        if key in ("model.3.weight", "model.3.bias"):
            bank[:, off:off + n] *= gain
        off += n
    eng = Net(arch, shape, hidden, C)
    eng.set_precision(prec)
    eng.upload(bank, 0)
    p0 = orc.bnn_forward(net, layout, bank, x, [0]).detach()
    assert float(p0.max(-1)[0].median()) > 0.99                       # the per-sample softmax rows really are saturated
    g = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S).cpu().reshape(x.shape) / S
    ref64 = orc.expected_loss_gradients(net, layout, bank, x, labels, range(S), dtype=torch.float64)
    ref32 = orc.expected_loss_gradients(net, layout, bank, x, labels, range(S))
    e, e32 = rel_err(g, ref64), rel_err(ref32, ref64)
    print(f"saturated x{gain:g} {prec}: rel err {e:.2e} (fp32 oracle vs fp64: {e32:.2e})")
    assert e < max(REL, 3 * e32)
    pbar = eng.forward_probs_sum(x, 0, S) / S
    ga = eng.input_grad_sum(_lib.HEAD_GRAD_OF_MEAN, x, labels, 0, S, pbar=pbar).cpu().reshape(x.shape) / S
    ra = orc.attack_gradient(net, layout, bank, x, labels, range(S), dtype=torch.float64)
    ra32 = orc.attack_gradient(net, layout, bank, x, labels, range(S))
    assert rel_err(ga, ra) < max(REL, 3 * rel_err(ra32, ra))
    eng.close()


@pytest.mark.parametrize("prec", ["fp32", "f16x3", "tf32x3"])
def test_autograd_through_forward_matches_attack_gradient(prec):
    """torch autograd through BNN.forward (UPSTREAM head).  On the tensor-core engines backward() reuses the logits and
    masks the forward kept; a second forward in between makes it fall back to the recomputing route."""
    from robustbnns_b200.model_bnn import BNN
    net, layout, loc, rho, bank, x, labels = _problem("fc", (1, 28, 28), 64, 10, 12, 4)
    bnn = BNN("mnist", 64, "leaky", "fc", "hmc", None, None, 4, 5, (1, 28, 28), 10)
    bnn.set_posterior_samples(bank)
    bnn.set_precision(prec)
    ref = orc.attack_gradient(net, layout, bank, x, labels, range(4), dtype=torch.float64)
    for interleave in (False, True):
        xg = x.cuda().requires_grad_(True)
        out = bnn.forward(xg, n_samples=4)
        assert bnn.engine().keep_valid == (prec != "fp32")
        if interleave:
            bnn.forward(x.cuda()[:5].requires_grad_(True), n_samples=4)     # overwrites the kept forward
        loss = torch.nn.CrossEntropyLoss(reduction="sum")(out, labels.cuda())
        loss.backward()
        assert rel_err(xg.grad.cpu(), ref) < REL


@pytest.mark.parametrize("arch,hidden,prec", [("fc", 512, "f16x3"), ("fc", 64, "tf32x3"), ("fc2", 128, "tf32x3"),
                                              ("conv", 32, "f16x3"), ("conv", 32, "fp32"), ("fc2", 128, "fp32")])
def test_single_image_and_empty_batch(arch, hidden, prec):
    """The reference's own calling pattern is ONE image at a time (lossGradients.py:29-38, adversarialAttacks.py:118-131):
    B = 1 through every engine, and the degenerate B = 0 / S = 0 calls."""
    from robustbnns_b200 import _lib
    from robustbnns_b200.engine import Net
    S = 3
    net, layout, loc, rho, bank, x, labels = _problem(arch, (1, 28, 28), hidden, 10, 2, S)
    eng = Net(arch, (1, 28, 28), hidden, 10)
    eng.set_precision(prec)
    eng.upload(bank, 0)
    tol = REL
    for i in range(2):
        xi, yi = x[i:i + 1], labels[i:i + 1]
        p = eng.forward_probs_sum(xi, 0, S).cpu() / S
        assert p.shape == (1, 10) and rel_err(p, orc.bnn_forward(net, layout, bank, xi, range(S)).detach()) < tol
        g = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xi, yi, 0, S).cpu().reshape(xi.shape) / S
        ref = orc.expected_loss_gradients(net, layout, bank, xi, yi, range(S), dtype=torch.float64)
        ref32 = orc.expected_loss_gradients(net, layout, bank, xi, yi, range(S))
        assert min(rel_err(g, ref), rel_err(g, ref32)) < tol
        pbar = eng.forward_probs_sum(xi, 0, S, keep=True) / S
        if eng.keep_valid:
            ga = eng.input_grad_sum_kept(_lib.HEAD_GRAD_OF_MEAN, yi, pbar=pbar).cpu().reshape(xi.shape) / S
        else:
            ga = eng.input_grad_sum(_lib.HEAD_GRAD_OF_MEAN, xi, yi, 0, S, pbar=pbar).cpu().reshape(xi.shape) / S
        assert rel_err(ga, orc.attack_gradient(net, layout, bank, xi, yi, range(S), dtype=torch.float64)) < tol
    # empty batch / empty sample range
    e = x[:0]
    assert eng.forward_probs_sum(e, 0, S).shape == (0, 10)
    assert eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, e, labels[:0], 0, S).numel() == 0
    assert float(eng.forward_probs_sum(x, 2, 2).abs().max()) == 0.0
    assert float(eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 1, 1).abs().max()) == 0.0
    eng.close()


def test_half_moons_grid_sweep(tmp_path, monkeypatch):
    """BASELINE configs[0] / [4]: the half-moons over-parametrisation sweep through the grid_search_halfMoons drop-in
    (fc2 2-H-H-2, stored posteriors in the reference's file layout, 100 test points x 100 samples) against the oracle."""
    from robustbnns_b200 import grid_search_halfMoons as gs
    from robustbnns_b200.adversarialAttacks import load_attack
    from robustbnns_b200.lossGradients import load_loss_gradients
    from robustbnns_b200.utils import load_half_moons
    monkeypatch.chdir(tmp_path)
    _, _, x_test, y_test, _, _ = load_half_moons()
    S, pts, widths = 100, 100, (32, 128, 512)
    banks = {}
    for hidden in widths:
        bnn = gs.MoonsBNN(hidden, "leaky", "fc2", "hmc", None, None, S, 5, 100, (1, 2, 1), 2)
        # D = 2: the CUDA-core engine for the narrow nets; from H = 64 the H x H layer runs on tcgen05 (TF32X3)
        assert bnn.engine().precision == ("tf32x3" if hidden >= 64 else "fp32")
        net = orc.build_net("fc2", (1, 2, 1), hidden, 2, dataset_name="half_moons")
        layout = orc.param_layout(net)
        loc, rho = orc.scaled_guide_params(layout, seed=hidden, rho_mean=-2.0)
        bank = loc + orc.softplus(rho) * torch.randn((S, loc.numel()), generator=torch.Generator().manual_seed(hidden))
        bnn.set_posterior_samples(bank)
        bnn.save(rel_path="w/")
        banks[hidden] = (net, layout, bank, bnn.name)
    gs.serial_compute_grads(list(widths), ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S],
                            rel_path="w/", test_points=pts)
    gs.grid_attack("fgsm", list(widths), ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S], pts,
                   device="cuda", rel_path="w/")
    xt, labels = torch.from_numpy(x_test[:pts]), torch.from_numpy(y_test[:pts]).argmax(-1)
    for hidden, (net, layout, bank, name) in banks.items():
        got = load_loss_gradients(n_samples=S, filename=name, savedir=name + "/")
        ref = orc.expected_loss_gradients(net, layout, bank, xt, labels, range(S), dtype=torch.float64).reshape(pts, 2).numpy()
        order = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]           # the loader shuffles the test points  # noqa: E731
        assert got.shape == (pts, 2) and np.abs(order(got) - order(ref)).max() <= REL * np.abs(ref).max()
        adv = load_attack("fgsm", name, n_samples=S)
        ref_adv = orc.fgsm_attack(net, layout, bank, xt, labels, lambda call: range(S), None, dtype=torch.float64)
        g64 = orc.attack_gradient(net, layout, bank, xt, labels, range(S), dtype=torch.float64)
        _assert_adv_equal_where_determined(adv, ref_adv, g64, what=("moons fgsm", hidden))
    # the batched form (all models enqueued back to back, one synchronisation): same files, same numbers
    serial = {name: load_loss_gradients(n_samples=S, filename=name, savedir=name + "/") for (_, _, _, name) in banks.values()}
    import shutil
    shutil.rmtree("data")
    out = gs.parallel_compute_grads(list(widths), ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S],
                                    rel_path="w/", test_points=pts)
    gs.parallel_grid_attack("fgsm", list(widths), ["leaky"], ["fc2"], ["hmc"], [None], [None], [S], [5], [100], [S],
                            rel_path="w/", test_points=pts)
    assert len(out) == len(widths)
    order = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]  # noqa: E731
    for hidden, (net, layout, bank, name) in banks.items():
        got = load_loss_gradients(n_samples=S, filename=name, savedir=name + "/")
        assert np.abs(order(got) - order(serial[name])).max() <= 1e-6 * np.abs(serial[name]).max()
        adv = load_attack("fgsm", name, n_samples=S)
        ref_adv = orc.fgsm_attack(net, layout, bank, xt, labels, lambda call: range(S), None, dtype=torch.float64)
        g64 = orc.attack_gradient(net, layout, bank, xt, labels, range(S), dtype=torch.float64)
        _assert_adv_equal_where_determined(adv, ref_adv, g64, what=("moons fgsm batched", hidden))


def test_default_engine_is_the_fastest_parity_grade():
    """A BNN that creates its own engine picks F16X3 (fc-512, conv), TF32X3 (fc2; half moons from H = 64) or FP32."""
    from robustbnns_b200.model_bnn import BNN
    from robustbnns_b200.model_nn import NN
    for arch, shape, hidden, C, ds, want in (("fc", (1, 28, 28), 512, 10, "mnist", "f16x3"),
                                             ("conv", (1, 28, 28), 64, 10, "mnist", "f16x3"),
                                             ("fc2", (1, 28, 28), 128, 10, "mnist", "tf32x3"),
                                             ("fc2", (1, 2, 1), 32, 2, "half_moons", "fp32"),
                                             ("fc2", (1, 2, 1), 512, 2, "half_moons", "tf32x3"),   # H x H layer on tcgen05
                                             ("fc", (1, 28, 28), 16, 10, "mnist", "fp32")):
        bnn = BNN(ds, hidden, "leaky", arch, "hmc", None, None, 2, 5, shape, C)
        assert bnn.engine().precision == want, (arch, hidden)
        bnn.set_precision("fp32")
        assert bnn.engine().precision == "fp32"
        bnn.set_precision("auto")
        assert bnn.engine().precision == want
    assert NN("mnist", (1, 28, 28), 10, 64, "leaky", "conv", 0.01, 1).engine().precision == "f16x3"
    assert NN("mnist", (1, 28, 28), 10, 64, "leaky", "fc2", 0.01, 1).engine().precision == "tf32x3"


# ------------------------------------------------------------------ sampler -----------------------------
def test_philox_sampler_matches_restatement_and_moments():
    from robustbnns_b200.engine import Net
    net = orc.build_net("fc", (1, 28, 28), 64, 10)
    layout = orc.param_layout(net)
    loc, rho = orc.scaled_guide_params(layout, seed=3, rho_mean=-2.0)
    rho[:50] = torch.linspace(15.0, 30.0, 50)              # crosses the softplus threshold (rho > 20)
    eng = Net("fc", (1, 28, 28), 64, 10)
    seed = 0x1234567812345678
    eng.sample_diag(loc, rho, seed, 5, 0, 3)                # global indices 5,6,7
    eng.sample_diag(loc, rho, seed, 100, 3, 4, stride=8)    # 100,108,116,124 (rank-strided)
    got = eng.download(0, 7)
    ref = orc.philox_bank(loc, rho, seed, [5, 6, 7, 100, 108, 116, 124])
    sd = orc.softplus(rho)
    assert float(((got - ref).abs() / sd).max()) < 2e-5      # same Philox bits; libm-level differences only
    # statistical check against the guide's moments N(loc, softplus(rho)^2), independent across samples
    S = 256
    eng.sample_diag(loc, rho, 7, 0, 0, S)
    z = (eng.download(0, S) - loc) / sd
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    assert float(z.mean(0).abs().max()) < 6.0 / np.sqrt(S)
    corr = float((z[0::2] * z[1::2]).mean())
    assert abs(corr) < 5e-3
    from scipy import stats
    assert stats.kstest(z[:, ::97].reshape(-1).numpy(), "norm").pvalue > 1e-3
    eng.close()


def test_fused_sampler_relayout_equals_two_step_path():
    """F16X3, arch fc: once the weight scale is fixed, rbnn_bank_sample_diag draws the bank rows AND writes the fp16
    operand copies in one kernel.  Same Philox bits as the plain sampler, and the same gradients as an engine that
    received those rows by upload (two-step path: upload -> lazy re-layout)."""
    from robustbnns_b200 import _lib
    from robustbnns_b200.engine import Net
    net, layout, loc, rho, bank, x, labels = _problem("fc", (1, 28, 28), 512, 10, 130, 2)
    S = 5
    eng = Net("fc", (1, 28, 28), 512, 10)
    eng.set_precision("f16x3")
    eng.sample_diag(loc, rho, 11, 0, 0, S)                                   # first draw: plain sampler, scale not fixed yet
    g0 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S)       # fixes the scale (lazy refresh)
    n0 = eng.launch_count
    eng.sample_diag(loc, rho, 11, 100, 0, S, stride=3)                       # fused path: indices 100, 103, ...
    assert eng.launch_count - n0 <= 5                                        # softplus, fused, tail, maxabs, freeze
    got = eng.download(0, S)
    ref = orc.philox_bank(loc, rho, 11, [100 + 3 * i for i in range(S)])
    # same Philox bits: libm-level differences in the normals (2e-5 sigma) plus the fp32 rounding of loc + sigma * eps
    assert bool(((got - ref).abs() <= 2e-5 * orc.softplus(rho) + 2.4e-7 * ref.abs()).all())
    g1 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S)
    other = Net("fc", (1, 28, 28), 512, 10)
    other.set_precision("f16x3")
    other.upload(got, 0)
    g2 = other.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S)
    assert rel_err(g1.cpu(), g2.cpu()) < 1e-6                                # (the two engines may fix different scales)
    ref64 = orc.expected_loss_gradients(net, layout, got, x, labels, range(S), dtype=torch.float64)
    assert rel_err(g1.cpu().reshape(x.shape) / S, ref64) < REL
    assert not torch.equal(g0, g1)
    # the plain sampler on the same handle (fp32 engine) draws the same rows
    eng.set_precision("fp32")
    eng.sample_diag(loc, rho, 11, 100, 0, S, stride=3)
    assert torch.equal(eng.download(0, S), got)
    eng.close(); other.close()


# ------------------------------------------------------------------ stateless kernels -------------------
def test_attack_and_evaluation_kernels_edge_cases():
    from robustbnns_b200 import engine as E
    from robustbnns_b200 import adversarialAttacks as aa
    g = torch.Generator().manual_seed(0)
    for B, D in ((1, 784), (5, 2), (33, 784), (7, 13)):
        x = torch.rand((B, D), generator=g).cuda()
        x0 = torch.rand((B, D), generator=g).cuda()
        gr = torch.randn((B, D), generator=g).cuda()
        gr[0, :D // 2] = 0.0                                 # sign(0) = 0
        ref = torch.clamp(x + 0.3 * gr.sign(), 0, 1)
        assert torch.equal(E.fgsm_step(x.reshape(-1), gr.reshape(-1), 0.3).reshape(B, D), ref)
        alpha = E.pgd_alpha(x)
        assert torch.equal(alpha, 2 / x.max(dim=1)[0])
        eta = torch.clamp(x + alpha[:, None] * gr.sign() - x0, min=-0.1, max=0.1)
        assert torch.equal(E.pgd_step(x, x0, gr, alpha, 0.1), torch.clamp(x0 + eta, 0, 1))
    # all-zero image: alpha = inf, inf*0 = NaN exactly as upstream
    z = torch.zeros((1, 16)).cuda()
    a = E.pgd_alpha(z)
    assert torch.isinf(a).all()
    out = E.pgd_step(z, z, torch.zeros_like(z), a, 0.3)
    assert torch.isnan(out).all()
    for N, C in ((0, 10), (1, 2), (129, 10), (1000, 10), (77, 32)):
        o0 = torch.randn((N, C), generator=g).cuda()
        o1 = torch.randn((N, C), generator=g).cuda()
        lab = torch.randint(0, C, (N,), generator=g).cuda()
        if N:
            o0[0] = 0.5                                      # ties: first maximum wins, like torch.argmax
        cnt = torch.zeros((1,), dtype=torch.int64).cuda()
        E.count_correct(o0, lab.to(torch.int32), cnt)
        assert int(cnt.item()) == int((o0.argmax(-1) == lab).sum().item())
        if N:
            rob = aa.softmax_robustness(o0, o1)
            assert float((rob.cpu() - orc.softmax_robustness(o0.cpu(), o1.cpu())).abs().max()) <= 1e-6
            d = aa.softmax_difference(o0, o1)
            assert float((d.cpu() - orc.softmax_difference(o0.cpu(), o1.cpu())).abs().max()) <= 1e-6
    with pytest.raises(ValueError):
        aa.softmax_difference(torch.rand(4, 10).cuda(), torch.rand(3, 10).cuda())


# ------------------------------------------------------------------ BASELINE-size properties ------------
def test_full_size_properties_fc_headline():
    """cfg2 shape (10 000 MNIST-shaped inputs, fc 784-512-10) with a few samples: size-independent
    properties the oracle cannot check in seconds."""
    from robustbnns_b200 import _lib
    from robustbnns_b200.engine import Net
    B, S = 10000, 6
    net, layout, loc, rho, bank, x, labels = _problem("fc", (1, 28, 28), 512, 10, B, S)
    eng = Net("fc", (1, 28, 28), 512, 10)
    eng.set_precision("f16x3")                     # the headline engine (what bench.py measures)
    eng.upload(bank, 0)
    xd, ld = x.cuda(), labels.cuda().to(torch.int32)
    full = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, 0, S)
    # (1) linearity over the sample range
    parts = sum(eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, s, s + 1) for s in range(S))
    assert rel_err(parts.cpu(), full.cpu()) < 1e-5
    # (2) permuting the batch permutes the rows
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).cuda()
    gp = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd[perm].contiguous(), ld[perm].contiguous(), 0, S)
    assert torch.equal(gp, full[perm])
    # (3) a random subset of rows agrees with the oracle
    idx = torch.randperm(B, generator=torch.Generator().manual_seed(2))[:64]
    ref = orc.expected_loss_gradients(net, layout, bank, x[idx], labels[idx], range(S), dtype=torch.float64)
    assert rel_err(full.cpu()[idx].reshape(ref.shape) / S, ref) < REL
    # (4) probabilities: rows sum to one
    p = eng.forward_probs_sum(xd, 0, S) / S
    assert float((p.sum(-1) - 1).abs().max()) < 1e-5
    # (5) host-buffer entry point == device entry point
    out = eng.loss_gradients_host(x.reshape(B, -1), labels, 0, S, S)
    assert rel_err(out, full.cpu().reshape(B, -1) / S) < 1e-6
    eng.close()


@pytest.mark.parametrize("inputs", ["pixel", "float"])
def test_headline_engine_at_headline_size_against_fp64_oracle(inputs):
    """BASELINE configs[1] at FULL size on the engine bench.py times: 10 000 inputs x 1000 posterior samples drawn on the
    device from bench.py's guide (loc ~ N(0, 1/fan_in), rho ~ N(-5, 1)), fc 784-512-10, F16X3.  32 random rows of the
    expected loss gradient are held to the fp64 oracle evaluated on the very weights the device drew (downloaded bank).
    "pixel": inputs on the 8-bit grid uint8 / 255 as the reference's loaders produce them (utils.py:102-103) and as
    bench.py draws them -- the two-pass forward; "float": arbitrary fp32 inputs -- the three-pass forward."""
    import math
    from robustbnns_b200 import lossGradients as lg
    from robustbnns_b200.model_bnn import BNN
    B, S = 10000, 1000
    net = orc.build_net("fc", (1, 28, 28), 512, 10)
    layout = orc.param_layout(net)
    g = torch.Generator().manual_seed(1)
    locs, rhos, fan = [], [], 784
    for key, shp in layout:                         # exactly bench.py's construction
        n = int(np.prod(shp))
        if len(shp) > 1:
            fan = shp[1]
        locs.append(torch.randn(n, generator=g) / math.sqrt(fan))
        rhos.append(torch.randn(n, generator=g) - 5.0)
    bnn = BNN("mnist", 512, "leaky", "fc", "svi", 1, 0.01, None, None, (1, 28, 28), 10)
    bnn.set_guide(torch.cat(locs), torch.cat(rhos))
    bnn.set_precision("f16x3")
    gx = torch.Generator().manual_seed(0)
    if inputs == "pixel":
        x = torch.randint(0, 256, (B, 1, 28, 28), generator=gx).to(torch.float32) / 255
    else:
        x = torch.rand((B, 1, 28, 28), generator=gx)
    y = torch.randint(0, 10, (B,), generator=gx)
    grads = lg.expected_loss_gradients(bnn, x.cuda(), y.cuda(), S).cpu()
    assert bnn.engine().input_grid == (inputs == "pixel")
    idx = torch.randperm(B, generator=torch.Generator().manual_seed(3))[:32]
    bank = bnn.engine().download(0, S)              # the 1000 weight vectors the Philox kernel drew (1.6 GB)
    ref = orc.expected_loss_gradients(net, layout, bank, x[idx], y[idx], range(S), dtype=torch.float64)
    err = rel_err(grads[idx], ref)
    print(f"headline size, f16x3, {inputs} inputs: 32 rows vs fp64 oracle: {err:.2e}")
    assert err < REL
    # the posterior-mean gradient is a heavily cancelling sum: per-row errors relative to that row's own maximum
    row_err = (grads[idx].double() - ref).abs().flatten(1).max(1)[0] / ref.abs().flatten(1).max(1)[0]
    assert float(row_err.max()) < 10 * REL, row_err


@pytest.mark.parametrize("arch,shape,hidden,ds", [("fc", (1, 28, 28), 64, "mnist"), ("fc2", (1, 28, 28), 32, "mnist"),
                                                  ("fc2", (1, 2, 1), 32, "half_moons")])
@pytest.mark.parametrize("act", ["relu", "sigm", "tanh"])
def test_other_activations_on_the_fp32_engine(act, arch, shape, hidden, ds):
    """model_nn.py:66-75: relu / sigm / tanh hidden layers (no saved model uses them) run for arch fc / fc2 on the FP32
    CUDA-core engine; the tensor-core modes and arch conv, which fuse LeakyReLU, refuse them."""
    from robustbnns_b200 import _lib
    from robustbnns_b200 import adversarialAttacks as aa
    from robustbnns_b200 import lossGradients as lg
    from robustbnns_b200.engine import Net
    from robustbnns_b200.model_bnn import BNN
    C, B, S = (10 if ds == "mnist" else 2), 37, 5
    net = orc.build_net(arch, shape, hidden, C, activation=act, dataset_name=ds)
    layout = orc.param_layout(net)
    loc, rho = orc.scaled_guide_params(layout, seed=9, rho_mean=-4.0)
    g = torch.Generator().manual_seed(10)
    bank = loc + orc.softplus(rho) * torch.randn((S, loc.numel()), generator=g)
    x, y = orc.synthetic_inputs(B, shape, C, seed=12)
    labels = y.argmax(-1)
    eng = Net(arch, shape, hidden, C)
    eng.set_activation(act)
    assert eng.precision == "fp32" and eng.set_best_precision() == "fp32"
    for prec in ("tf32x3", "f16x3", "bf16"):
        with pytest.raises(_lib.RbnnError):
            eng.set_precision(prec)
    eng.upload(bank, 0)
    xd, ld = x.cuda(), labels.cuda().to(torch.int32)
    probs = eng.forward_probs_sum(xd, 0, S) / S
    assert rel_err(probs.cpu(), orc.bnn_forward(net, layout, bank, x, range(S)).detach()) < REL
    gm = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, 0, S).cpu().reshape(x.shape) / S
    assert rel_err(gm, orc.expected_loss_gradients(net, layout, bank, x, labels, range(S), dtype=torch.float64)) < REL
    ga = eng.input_grad_sum(_lib.HEAD_GRAD_OF_MEAN, xd, ld, 0, S, pbar=probs).cpu().reshape(x.shape) / S
    assert rel_err(ga, orc.attack_gradient(net, layout, bank, x, labels, range(S), dtype=torch.float64)) < REL
    eng.close()
    # through the drop-in: the reference's constructor arguments
    bnn = BNN(ds, hidden, act, arch, "hmc", None, None, S, 5, shape, C)
    bnn.set_posterior_samples(bank)
    assert bnn.engine().precision == "fp32"
    gd = lg.expected_loss_gradients(bnn, x, labels, S).cpu()
    assert rel_err(gd, gm) < 1e-6
    adv = aa.fgsm_attack(bnn, x, labels, hyperparams={"epsilon": 0.1}, n_samples=S).cpu()
    ref_adv = orc.fgsm_attack(net, layout, bank, x, labels, lambda call: range(S), {"epsilon": 0.1})
    gref = orc.attack_gradient(net, layout, bank, x, labels, range(S), dtype=torch.float64)
    decided = gref.abs() > REL * gref.abs().max()             # pixels whose gradient sign is above the tolerance
    assert float((adv - ref_adv).abs()[decided].max()) < 1e-6
    with pytest.raises(NotImplementedError):
        BNN("mnist", 32, act, "conv", "hmc", None, None, S, 5, (1, 28, 28), 10)


@pytest.mark.parametrize("S", [1, 10, 1000])
def test_two_pass_forward_for_inputs_on_the_pixel_grid(S):
    """F16X3, arch fc: inputs that are uint8 / 255 (every image set the reference loads, utils.py:102-103, 129-130,
    190-191) have a zero low half once scaled by 255 * 2^j, so the fused forward issues TWO tensor-core passes instead of
    three.  That route must be parity grade on its own: expected loss gradients, the attack gradient and the mean
    prediction against the fp64 oracle at 1, 10 and 1000 posterior samples (tolerance 1e-4, the north-star one), the
    detection must be bit-exact (inputs a hair off the grid, or with more than 11 significant bits, take three passes),
    and a PGD iterate, which leaves the grid, must switch back by itself."""
    from robustbnns_b200 import _lib
    from robustbnns_b200.engine import Net
    hidden, B = (512, 130) if S <= 10 else (256, 96)
    net, layout, loc, rho, bank, x, labels = _problem("fc", (1, 28, 28), hidden, 10, B, S, seed=11)
    g = torch.Generator().manual_seed(4)
    q = torch.randint(0, 256, x.shape, generator=g)
    q[:, :, :6, :] = 0                                      # a blank border, as MNIST has
    x = q.to(torch.float32) / 255                           # exactly what `x_test /= 255` yields
    eng = Net("fc", (1, 28, 28), hidden, 10)
    eng.set_precision("f16x3")
    eng.upload(bank, 0)
    xd, ld = x.cuda(), labels.cuda().to(torch.int32)
    gm = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, 0, S).cpu().reshape(x.shape) / S
    assert eng.input_grid, "inputs on the pixel grid were not recognised"
    ref = orc.expected_loss_gradients(net, layout, bank, x, labels, range(S), dtype=torch.float64)
    e_mean = rel_err(gm, ref)
    probs = eng.forward_probs_sum(xd, 0, S) / S
    e_p = rel_err(probs.cpu(), orc.bnn_forward(net, layout, bank, x, range(S)).detach())
    ga = eng.input_grad_sum(_lib.HEAD_GRAD_OF_MEAN, xd, ld, 0, S, pbar=probs).cpu().reshape(x.shape) / S
    e_att = rel_err(ga, orc.attack_gradient(net, layout, bank, x, labels, range(S), dtype=torch.float64))
    pk = eng.forward_probs_sum(xd, 0, S, keep=True)        # two-phase form of the attacks
    gk = eng.input_grad_sum_kept(_lib.HEAD_GRAD_OF_MEAN, ld, pbar=pk / S).cpu().reshape(x.shape) / S
    assert rel_err(gk, ga) < 1e-6
    print(f"two-pass forward, fc-{hidden} B={B} S={S}: mean-of-grads {e_mean:.2e} grad-of-mean {e_att:.2e} probs {e_p:.2e}")
    assert max(e_mean, e_att, e_p) < REL
    # the three-pass route on the same inputs (RBNN_XGRID cannot be flipped inside a process: perturb ONE pixel instead)
    x3 = x.clone()
    x3[0, 0, 10, 10] = x3[0, 0, 10, 10] * (1 + 1e-6) if float(x3[0, 0, 10, 10]) > 0 else 1e-9
    g3 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x3.cuda(), ld, 0, S).cpu().reshape(x.shape) / S
    assert not eng.input_grid, "one input off the grid by 1e-6 (relative) must send the call down the three-pass route"
    assert rel_err(g3[1:], gm[1:]) < 2e-5                  # rows 1.. did not change: both routes agree far below 1e-4
    # 12 significant bits (q up to 4095) do not fit the fp16 hi part
    x12 = torch.randint(0, 4096, x.shape, generator=g).to(torch.float32) / 255
    eng.forward_probs_sum(x12.cuda(), 0, S)
    assert not eng.input_grid
    # all-zero inputs: no scale to derive, three-pass route, finite result
    z0 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, torch.zeros_like(xd), ld, 0, S)
    assert not eng.input_grid and bool(torch.isfinite(z0).all())
    # a PGD iterate x + alpha * sign(g) leaves the grid: back to three passes without being told
    xi = (xd + (2.0 / 225) * torch.sign(torch.randn(xd.shape, device=xd.device))).clamp(0, 1)
    gi = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xi, ld, 0, S).cpu().reshape(x.shape) / S
    assert not eng.input_grid
    refi = orc.expected_loss_gradients(net, layout, bank, xi.cpu(), labels, range(S), dtype=torch.float64)
    assert rel_err(gi, refi) < REL
    eng.close()


@pytest.mark.parametrize("prec", ["f16x3", "tf32x3"])
def test_full_size_properties_conv_cfg4(prec):
    """BASELINE configs[3] shape (100 F-MNIST-shaped inputs, conv-512, 50 stored posterior samples) on the tensor-core
    conv engine: size-independent properties, and EVERY one of the 5000 (sample, image) units against the fp64 oracle."""
    import math
    from robustbnns_b200 import _lib
    from robustbnns_b200.engine import Net
    B, S, H = 100, 50, 512
    net = orc.build_net("conv", (1, 28, 28), H, 10)
    layout = orc.param_layout(net)
    g = torch.Generator().manual_seed(21)
    cols = []
    for key, shp in layout:                         # independent networks, N(0, 1/fan_in): an HMC-like bank
        n = int(np.prod(shp))
        fan = n // shp[0] if len(shp) > 1 else 25
        cols.append(torch.randn((S, n), generator=g) / math.sqrt(fan))
    bank = torch.cat(cols, dim=1)
    x = torch.rand((B, 1, 28, 28), generator=g)
    labels = torch.randint(0, 10, (B,), generator=g)
    eng = Net("conv", (1, 28, 28), H, 10)
    eng.set_precision(prec)
    eng.upload(bank, 0)
    xd, ld = x.cuda(), labels.cuda().to(torch.int32)
    full = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, 0, S)
    # (1) deterministic: a second evaluation is bit-identical
    assert torch.equal(eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, 0, S), full)
    # (2) linearity over the sample range (what sample sharding relies on)
    parts = sum(eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, s, min(s + 17, S)) for s in range(0, S, 17))
    assert rel_err(parts.cpu(), full.cpu()) < 1e-5
    # (3) permuting the batch permutes the rows (up to the summation order of the sample slices)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).cuda()
    gp = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd[perm].contiguous(), ld[perm].contiguous(), 0, S)
    assert rel_err(gp.cpu(), full[perm].cpu()) < 1e-5
    # (4) the FULL workload against the fp64 oracle, unit by unit.  A unit holds 4608 first-layer and 32 768 second-layer
    # activations, each deciding a LeakyReLU sign and competing in a pooling window, and the gradient jumps by 1e-3..5e-2
    # of its size whenever one of those decisions flips: the reference's own fp32 arithmetic disagrees with fp64 on 70 of
    # these 5000 units (profiles/r2_conv_cfg4_oracle.json).  The tensor-core engine takes every decision as exact
    # arithmetic does (first layer accumulated in fp64, guard-band re-evaluation of the second, fp32 ties of the pooling
    # settled from the exact values), so ALL units must meet the north-star tolerance.
    ref = torch.stack([orc.expected_loss_gradients(net, layout, bank, x, labels, [smp], dtype=torch.float64)
                       for smp in range(S)])                                                  # [S, B, 1, 28, 28]
    got = torch.stack([eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, smp, smp + 1).cpu().reshape(x.shape)
                       for smp in range(S)]).double()
    errs = (got - ref).abs().flatten(2).max(-1)[0] / ref.abs().flatten(2).max(-1)[0]           # [S, B]
    print(f"conv cfg4 {prec}: worst of {errs.numel()} units {float(errs.max()):.2e}, units above 1e-4: {int((errs > REL).sum())}")
    assert int((errs > REL).sum()) == 0, errs.max()
    assert rel_err(full.cpu().reshape(x.shape).double() / S, ref.mean(0)) < REL                # the expected gradient itself
    # (5) probabilities: rows sum to one; mean logits of the bank == sum of single rows
    p = eng.forward_probs_sum(xd, 0, S) / S
    assert float((p.sum(-1) - 1).abs().max()) < 1e-5
    ls = eng.forward_logits_sum(xd, 0, 5)
    assert rel_err(sum(eng.forward_logits(xd, s) for s in range(5)).cpu(), ls.cpu()) < 1e-5
    eng.close()


def test_pgd_cuda_graph_equals_eager_loop():
    """The captured-graph PGD loop (one CUDA graph per iteration, fresh draws indexed by a device-resident counter) must
    give bit-identical adversarial examples to the eager loop, for fresh SVI draws and for a stored bank, also when the
    cached graph is replayed by a second call and after the posterior changed."""
    from robustbnns_b200 import adversarialAttacks as aa
    from robustbnns_b200.model_bnn import BNN
    N, S = 300, 20
    net = orc.build_net("fc", (1, 28, 28), 512, 10)
    layout = orc.param_layout(net)
    loc, rho = orc.scaled_guide_params(layout, seed=4, rho_mean=-5.0)
    x, y = orc.synthetic_inputs(N, (1, 28, 28), 10, seed=9)
    xd, yd = x.cuda(), y.argmax(-1).cuda()
    bnn = BNN("mnist", 512, "leaky", "fc", "svi", 1, 0.01, None, None, (1, 28, 28), 10)
    bnn.set_guide(loc, rho)
    bnn.set_precision("f16x3")

    def run(graph, hyper, iters=9):
        aa.PGD_GRAPH = graph
        try:
            bnn.reseed(0)
            return aa.pgd_attack(bnn, xd, yd, hyperparams=hyper, n_samples=S, iters=iters)
        finally:
            aa.PGD_GRAPH = True

    for hyper in (None, {"epsilon": 0.1}):
        eager = run(False, hyper)
        assert torch.equal(run(True, hyper), eager)          # captures
        assert torch.equal(run(True, hyper), eager)          # replays the cached graph
        assert len(bnn._pgd_graphs) >= 1
    # the fresh-draw counter ends where the eager loop leaves it: the next unseeded forward draws the same samples
    bnn.reseed(0)
    run(False, None)
    a = bnn.forward(xd[:8], n_samples=4)
    bnn.reseed(0)
    run(True, None)
    assert torch.equal(bnn.forward(xd[:8], n_samples=4), a)
    # another posterior: the cached graphs die with the old one
    loc2, rho2 = orc.scaled_guide_params(layout, seed=5, rho_mean=-4.0)
    bnn.set_guide(loc2, rho2)
    assert len(bnn._pgd_graphs) == 0
    assert torch.equal(run(True, None), run(False, None))
    # stored bank (HMC-like): no sampler inside the graph
    bank = loc + orc.softplus(rho) * torch.randn((S, loc.numel()), generator=torch.Generator().manual_seed(2))
    bnn.set_posterior_samples(bank)
    eager = run(False, None)
    assert torch.equal(run(True, None), eager)
    ref1 = orc.pgd_attack(net, layout, bank, x[:16], y.argmax(-1)[:16], lambda call: range(S), None, iters=1)
    aa.PGD_GRAPH = False
    one = aa.pgd_attack(bnn, xd[:16], yd[:16], hyperparams=None, n_samples=S, iters=1)
    aa.PGD_GRAPH = True
    g64 = orc.attack_gradient(net, layout, bank, x[:16], y.argmax(-1)[:16], range(S), dtype=torch.float64)
    _assert_adv_equal_where_determined(one, ref1, g64, what="pgd one step")


def test_f16x3_operand_scale_follows_the_posterior():
    """the F16X3 weight scale is frozen on first use.  A later posterior with much larger weights must not
    return saturated results: (a) through the drop-in (set_posterior_samples / set_guide reset what was derived from the
    old weights) and (b) through the bare engine, where rows uploaded over a frozen scale that does not fit them are
    detected and re-laid under a new scale inside the same call."""
    from robustbnns_b200 import _lib
    from robustbnns_b200 import lossGradients as lg
    from robustbnns_b200.engine import Net
    from robustbnns_b200.model_bnn import BNN
    B, S = 64, 4
    net, layout, loc, rho, bank, x, labels = _problem("fc", (1, 28, 28), 512, 10, B, S)
    big = bank.clone()
    big[:, :512 * 784] *= 300.0                     # first-layer weights 300x larger: far outside the old scale's 2^6 head room
    big[:, 512 * 784 + 512:512 * 784 + 512 + 5120] /= 300.0    # keep the logits in range
    bnn = BNN("mnist", 512, "leaky", "fc", "hmc", None, None, S, 5, (1, 28, 28), 10)
    bnn.set_precision("f16x3")
    for bk in (bank, big, bank):
        bnn.set_posterior_samples(bk)
        g = lg.expected_loss_gradients(bnn, x, labels, S).cpu()
        ref = orc.expected_loss_gradients(net, layout, bk, x, labels, range(S), dtype=torch.float64)
        assert bool(torch.isfinite(g).all()) and rel_err(g, ref) < REL
    eng = Net("fc", (1, 28, 28), 512, 10)
    eng.set_precision("f16x3")
    for bk in (bank, big):                          # no invalidate() in between: the call repairs the scale itself
        eng.upload(bk, 0)
        g = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, x, labels, 0, S).cpu().reshape(x.shape) / S
        ref = orc.expected_loss_gradients(net, layout, bk, x, labels, range(S), dtype=torch.float64)
        assert bool(torch.isfinite(g).all()) and rel_err(g, ref) < REL
    eng.close()


def test_full_size_properties_pgd_cfg3():
    """BASELINE configs[2] shape: Bayesian FGSM / PGD on 1000 MNIST-shaped inputs, fc 784-512-10, fresh SVI samples."""
    from robustbnns_b200 import adversarialAttacks as aa
    from robustbnns_b200.model_bnn import BNN
    N, S = 1000, 100
    net = orc.build_net("fc", (1, 28, 28), 512, 10)
    layout = orc.param_layout(net)
    loc, rho = orc.scaled_guide_params(layout, seed=4, rho_mean=-5.0)
    x, y = orc.synthetic_inputs(N, (1, 28, 28), 10, seed=9)
    labels = y.argmax(-1)
    bnn = BNN("mnist", 512, "leaky", "fc", "svi", 1, 0.01, None, None, (1, 28, 28), 10)
    bnn.set_guide(loc, rho)
    bnn.set_precision("f16x3")
    xd, yd = x.cuda(), labels.cuda()
    for hyper, eps in (({"epsilon": 0.1}, 0.1), (None, 0.5)):
        bnn.reseed(0)
        adv = aa.pgd_attack(bnn, xd, yd, hyperparams=hyper, n_samples=S, iters=20)
        assert adv.shape == xd.shape and bool(torch.isfinite(adv).all())
        assert float(adv.min()) >= 0.0 and float(adv.max()) <= 1.0
        assert float((adv - xd).abs().max()) <= eps + 1e-6                  # inside the L-inf ball around the originals
        bnn.reseed(0)
        again = aa.pgd_attack(bnn, xd, yd, hyperparams=hyper, n_samples=S, iters=20)
        assert torch.equal(adv, again)                                      # same fresh-sample stream => identical
    bnn.reseed(0)
    f0 = aa.fgsm_attack(bnn, xd, yd, hyperparams={"epsilon": 0.0}, n_samples=S)
    assert torch.equal(f0, xd)                                              # eps = 0 leaves the inputs alone
    bnn.reseed(0)
    f1 = aa.fgsm_attack(bnn, xd, yd, hyperparams={"epsilon": 0.2}, n_samples=S)
    d = (f1 - xd).abs()
    assert float(d.max()) <= 0.2 + 1e-6 and float(((d - 0.2).abs() < 1e-6).float().mean()) > 0.5    # full steps unless clipped
    # the attack must lower the expected probability of the true class on average (it ascends the loss)
    bnn.reseed(1)
    p_clean = bnn.forward(xd, n_samples=S).gather(1, yd[:, None]).mean()
    bnn.reseed(1)
    p_adv = bnn.forward(f1, n_samples=S).gather(1, yd[:, None]).mean()
    assert float(p_adv) < float(p_clean)
