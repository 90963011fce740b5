"""The oracle (oracle/oracle.py) held to the vectors the reference's own code produced
(tests/golden/*.npz, generator tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.helpers import ENS_CASES, HMC_CASES, SVI_CASES, Case, rel_err

TOL = 1e-5  # fp32 torch vs fp32 torch, different batching/summation order only


@pytest.mark.parametrize("name", SVI_CASES)
def test_svi_forward_modes(name):
    c = Case(name)
    S = c.bank.shape[0]
    probs = orc.bnn_forward(c.net, c.layout, c.bank, c.x, range(S))
    assert rel_err(probs, c.t("probs_seeded")) < TOL
    logits = orc.bnn_forward_avg_posterior(c.net, c.layout, c.t("loc"), c.x)
    assert rel_err(logits, c.t("logits_avg")) < TOL


@pytest.mark.parametrize("name", SVI_CASES)
def test_guide_sampler_restatement_matches_shim_draws(name):
    """guide_sample_bank consumes the RNG like the reference's guide does (under the shim)."""
    c = Case(name)
    S = c.bank.shape[0]
    bank = orc.guide_sample_bank(c.t("loc"), c.t("rho"), c.layout, range(S))
    assert torch.equal(bank, c.bank)


@pytest.mark.parametrize("name", SVI_CASES + HMC_CASES)
def test_loss_gradient_both_orders(name):
    c = Case(name)
    S = c.bank.shape[0]
    ref = c.t("loss_gradient")
    r1 = torch.stack([orc.loss_gradient_r1(c.net, c.layout, c.bank, c.x[i], c.y[i], S) for i in range(len(c.x))])
    assert rel_err(r1, ref) < TOL
    r2 = orc.expected_loss_gradients(c.net, c.layout, c.bank, c.x, c.labels, range(S))
    assert rel_err(r2, ref) < TOL
    r2d = orc.expected_loss_gradients(c.net, c.layout, c.bank, c.x, c.labels, range(S), dtype=torch.float64)
    assert rel_err(r2d, ref) < 1e-4   # fp64 tie-breaker: the reference's own fp32 rounding is ~2e-5 on conv


@pytest.mark.parametrize("name", SVI_CASES)
def test_loss_gradients_numpy_and_evaluate(name):
    c = Case(name)
    S = c.bank.shape[0]
    out = orc.loss_gradients_r1(c.net, c.layout, c.bank, c.x, c.y, S)
    ref = c.z["loss_gradients_np"]
    assert out.shape == ref.shape          # squeezed: [N,28,28] / [N,2]   (lossGradients.py:66)
    assert rel_err(out, ref) < TOL
    assert orc.evaluate(c.net, c.layout, c.bank, c.x, c.y, S, batch_size=2) == float(c.z["evaluate_acc"])


@pytest.mark.parametrize("name", SVI_CASES)
def test_unseeded_svi_fgsm_fresh_draws(name):
    """Unseeded SVI: image i's forward call consumed draws [i*S, (i+1)*S) (model_bnn.py:230-232)."""
    c = Case(name)
    S = c.bank.shape[0]
    fresh = c.t("fresh_bank")
    assert fresh.shape[0] == S * len(c.x)
    hyper = {"epsilon": float(c.z["fgsm_fresh_eps"])}
    adv = torch.cat([orc.fgsm_attack(c.net, c.layout, fresh, c.x[i:i + 1], c.labels[i:i + 1],
                                     lambda call, i=i: range(i * S, (i + 1) * S), hyper)
                     for i in range(len(c.x))])
    assert float((adv - c.t("fgsm_fresh")).abs().max()) <= 1e-6


@pytest.mark.parametrize("name", HMC_CASES)
def test_hmc_attacks_and_evaluation(name):
    c = Case(name)
    S = c.bank.shape[0]
    sched = lambda call: range(S)  # noqa: E731  HMC: first n stored nets, every call (model_bnn.py:248-249)
    assert rel_err(orc.bnn_forward(c.net, c.layout, c.bank, c.x, range(S)), c.t("probs")) < TOL
    hyper = {"epsilon": float(c.z["eps"])}
    for method, fn in (("fgsm", orc.fgsm_attack), ("pgd", orc.pgd_attack)):
        for hname, h in (("hyper", hyper), ("default", None)):
            adv = fn(c.net, c.layout, c.bank, c.x, c.labels, sched, h)
            ref = c.t(f"{method}_{hname}_adv")
            # sign ties are measure-zero; everything else must agree to fp32 rounding
            assert float((adv - ref).abs().max()) <= 1e-6, (method, hname)
            o, a, rob = orc.attack_evaluation(c.net, c.layout, c.bank, c.x, ref, c.y, sched)
            assert [o, a] == c.z[f"{method}_{hname}_eval"].tolist()
            assert float((rob - c.t(f"{method}_{hname}_rob")).abs().max()) <= 1e-6


@pytest.mark.parametrize("name", ENS_CASES)
def test_ensemble_and_deterministic_nets(name):
    """Ensemble_NN / NN rows against the reference's own outputs."""
    c = Case(name)
    size, used = int(c.z["size"]), int(c.z["n_used"])
    assert rel_err(orc.ensemble_forward(c.net, c.layout, c.bank, c.x, range(used)), c.t("logits_used")) < TOL
    assert rel_err(orc.ensemble_forward(c.net, c.layout, c.bank, c.x, range(size)), c.t("logits_all")) < TOL
    assert rel_err(orc.ensemble_forward(c.net, c.layout, c.bank, c.x, [0]), c.t("logits_member0")) < TOL
    hyper = {"epsilon": float(c.z["eps"])}
    for who, members in (("ens", range(used)), ("nn", [0])):
        for method, fn in (("fgsm", orc.ensemble_fgsm_attack), ("pgd", orc.ensemble_pgd_attack)):
            for hname, h in (("hyper", hyper), ("default", None)):
                adv = fn(c.net, c.layout, c.bank, c.x, c.labels, members, h)
                ref = c.t(f"{who}_{method}_{hname}_adv")
                assert float(((adv - ref).abs() > 1e-6).float().mean()) <= 2e-3, (who, method, hname)
                o, a, rob = orc.ensemble_attack_evaluation(c.net, c.layout, c.bank, c.x, ref, c.y, members)
                assert [o, a] == c.z[f"{who}_{method}_{hname}_eval"].tolist()
                assert float((rob - c.t(f"{who}_{method}_{hname}_rob")).abs().max()) <= 1e-6


def test_softmax_difference_errors():
    a = torch.rand(4, 10)
    with pytest.raises(ValueError):
        orc.softmax_difference(a, torch.rand(3, 10))
    with pytest.raises(ValueError):
        orc.check_seeds([0, 1], 3)
    with pytest.raises(ValueError):
        orc.build_net("fc", (1, 28, 28), 24, 10)
    with pytest.raises(NotImplementedError):
        orc.build_net("conv2", (1, 28, 28), 16, 10)


def test_philox_restatement_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    out = orc.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(v[0]) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = orc.philox4x32_10([0xffffffff], [0xffffffff], [0xffffffff], [0xffffffff], 0xffffffff, 0xffffffff)
    assert [int(v[0]) for v in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = orc.philox4x32_10([0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344], 0xa4093822, 0x299f31d0)
    assert [int(v[0]) for v in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    z = orc.philox_standard_normals(1234, 5, 200000)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
