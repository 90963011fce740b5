"""CPU oracle for the robustBNNs hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A torch-CPU restatement of the reference's arithmetic for the Bayesian expected
loss gradient, the Bayesian FGSM/PGD attacks and the attack evaluation.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this module; the product package `robustbnns_b200`
never does (tests/test_boundary.py::test_product_does_not_import_oracle).

Parity status
-------------
* The reference ships no golden vectors.  The oracle is
  pinned instead against the reference's OWN unmodified Python
  (`/root/reference/{model_nn,model_bnn,lossGradients,adversarialAttacks}.py`)
  executed in the build container under a minimal Pyro-1.3.0 stand-in
  (`oracle/pyro_shim`; Pyro itself is not installable here).  The vectors that
  run produced are committed under `tests/golden/` together with the generator
  `tests/golden/make_golden.py`, and `tests/test_oracle_golden.py` holds the
  oracle to them.
* Boundary with Pyro's sampler (which torch RNG stream a guide draw consumes):
  PARITY UNPINNED -- restated from Pyro 1.3.0's published source
  (`guide_sample_bank`), not verifiable against real Pyro here.  All exact
  parity tests therefore run on explicit posterior-sample banks `[S, P]`.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import copy
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as nnf

ARCHS = ("fc", "fc2", "conv")


# --------------------------------------------------------------------------
# a1  NN.set_model / NN.forward                      model_nn.py:60-141
# --------------------------------------------------------------------------
class _Net(nn.Module):
    """Holder so that state_dict keys read "model.<idx>.<weight|bias>" as in the
    reference (`NN.model` is the nn.Sequential, model_nn.py:78,85,98)."""

    def __init__(self, seq: nn.Sequential):
        super().__init__()
        self.model = seq

    def forward(self, x):
        return self.model(x)


def build_net(arch: str, input_shape: Sequence[int], hidden: int, n_classes: int,
              activation: str = "leaky", dataset_name: str = "mnist") -> _Net:
    """model_nn.py:36-40 (hidden-size check) and :60-124 (set_model)."""
    if math.log(hidden, 2).is_integer() is False or hidden < 16:        # model_nn.py:39-40
        raise ValueError("\nhidden size should be a power of 2 greater than 16.")
    input_size = input_shape[0] * input_shape[1] * input_shape[2]       # :62
    in_channels = input_shape[0]                                        # :63
    if activation == "relu":                                            # :66-75
        activ = nn.ReLU
    elif activation == "leaky":
        activ = nn.LeakyReLU
    elif activation == "sigm":
        activ = nn.Sigmoid
    elif activation == "tanh":
        activ = nn.Tanh
    else:
        raise AssertionError("\nWrong activation name.")
    if arch == "fc":                                                    # :77-82
        seq = nn.Sequential(nn.Flatten(), nn.Linear(input_size, hidden), activ(),
                            nn.Linear(hidden, n_classes))
    elif arch == "fc2":                                                 # :84-91
        seq = nn.Sequential(nn.Flatten(), nn.Linear(input_size, hidden), activ(),
                            nn.Linear(hidden, hidden), activ(), nn.Linear(hidden, n_classes))
    elif arch == "conv":                                                # :93-106
        if dataset_name not in ["mnist", "fashion_mnist"]:
            raise NotImplementedError()
        seq = nn.Sequential(nn.Conv2d(in_channels, 32, kernel_size=5), activ(),
                            nn.MaxPool2d(kernel_size=2),
                            nn.Conv2d(32, hidden, kernel_size=5), activ(),
                            nn.MaxPool2d(kernel_size=2, stride=1), nn.Flatten(),
                            nn.Linear(int(hidden / (4 * 4)) * input_size, n_classes))
    else:                                                               # :123-124 (conv2 is broken upstream)
        raise NotImplementedError()
    return _Net(seq)


def param_layout(net: _Net) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict() key order == the order BNN.guide iterates (model_bnn.py:124)."""
    return [(k, tuple(v.shape)) for k, v in net.state_dict().items()]


def param_count(layout) -> int:
    return int(sum(int(np.prod(s)) for _, s in layout))


def unpack(row: torch.Tensor, layout) -> Dict[str, torch.Tensor]:
    """One bank row [P] -> {state_dict key: tensor} (the HMC state-dict form, model_bnn.py:184-190)."""
    out, off = {}, 0
    for k, shp in layout:
        n = int(np.prod(shp))
        out[k] = row[off:off + n].reshape(shp)
        off += n
    assert off == row.numel()
    return out


def pack(state: Dict[str, torch.Tensor], layout) -> torch.Tensor:
    return torch.cat([state[k].reshape(-1) for k, _ in layout])


def net_logits(net: _Net, weights: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """f_w(x): the deterministic network with the given weights (model_nn.py:126-141)."""
    return torch.func.functional_call(net, weights, (x,))


# --------------------------------------------------------------------------
# a2  BNN.guide posterior parametrisation           model_bnn.py:121-136
# --------------------------------------------------------------------------
def softplus(x: torch.Tensor) -> torch.Tensor:
    """torch.nn.Softplus() defaults beta=1, threshold=20 (model_bnn.py:18)."""
    return nnf.softplus(x, beta=1.0, threshold=20.0)


def guide_sample_bank(loc: torch.Tensor, rho: torch.Tensor, layout, seeds: Sequence[int]) -> torch.Tensor:
    """Seeded guide draws, one row per seed (model_bnn.py:222-226 + :121-130).

    Per seed: set_rng_seed(seed); for every key the guide evaluates two eager
    `torch.randn_like(value)` init arguments (:125-126, discarded once the
    params exist); then random_module draws each parameter as
    Normal(loc, softplus(scale)).rsample() in named_parameters() order (:130).
    PARITY UNPINNED w.r.t. real Pyro/torch-1.4 RNG streams (see header).
    """
    rows = []
    for seed in seeds:
        torch.manual_seed(int(seed))
        for _, shp in layout:
            torch.randn(shp)
            torch.randn(shp)
        parts, off = [], 0
        for _, shp in layout:
            n = int(np.prod(shp))
            mu = loc[off:off + n].reshape(shp)
            sd = softplus(rho[off:off + n].reshape(shp))
            parts.append(torch.distributions.Normal(mu, sd).rsample().reshape(-1))
            off += n
        rows.append(torch.cat(parts))
    return torch.stack(rows)


# ---- restatement of the DEVICE sampler (robustbnns_b200/csrc/sampler.cu) ----
_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011) on numpy uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(a, dtype=np.uint32).copy() for a in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PH_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PH_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_standard_normals(seed: int, sample_index: int, n: int) -> np.ndarray:
    """eps[0:n] of global posterior sample `sample_index` exactly as the device
    sampler defines it: element 4q+j comes from Philox counter
    (q, sample_index, 0x5242'4e4e, 0) under key (seed_lo, seed_hi); uniforms are
    ((x >> 9) + 0.5) * 2^-23 (exact in fp32); Box-Muller pairs (x0,x1) -> j=0,1
    and (x2,x3) -> j=2,3."""
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint32)
    x = philox4x32_10(q, np.full(nq, sample_index, np.uint32), np.full(nq, 0x52424E4E, np.uint32),
                      np.zeros(nq, np.uint32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = [((xi >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23) for xi in x]
    out = np.empty((nq, 4), dtype=np.float32)
    for j, (ua, ub) in enumerate(((u[0], u[1]), (u[2], u[3]))):
        r = np.sqrt(np.float32(-2.0) * np.log(ua, dtype=np.float32), dtype=np.float32)
        th = np.float32(2.0 * math.pi) * ub
        out[:, 2 * j] = r * np.cos(th, dtype=np.float32)
        out[:, 2 * j + 1] = r * np.sin(th, dtype=np.float32)
    return out.reshape(-1)[:n]


def philox_bank(loc: torch.Tensor, rho: torch.Tensor, seed: int, sample_ids: Sequence[int]) -> torch.Tensor:
    """w_s = loc + softplus(rho) * eps_s with the device sampler's eps (fp32)."""
    sd = softplus(rho.float())
    rows = [loc.float() + sd * torch.from_numpy(philox_standard_normals(seed, int(s), loc.numel()))
            for s in sample_ids]
    return torch.stack(rows)


# --------------------------------------------------------------------------
# a3  BNN.forward                                    model_bnn.py:198-258
# --------------------------------------------------------------------------
def bnn_forward(net: _Net, layout, bank: torch.Tensor, x: torch.Tensor,
                sample_ids: Sequence[int]) -> torch.Tensor:
    """mean_s softmax(f_{w_s}(x))  -- model_bnn.py:222-226 / :251-257.
    `sample_ids` indexes rows of `bank` (seeded-SVI seed i == row i; HMC
    posterior_predictive[i] == row i)."""
    preds = [nnf.softmax(net_logits(net, unpack(bank[int(s)], layout), x), dim=-1) for s in sample_ids]
    return torch.stack(preds).mean(0)


def bnn_forward_avg_posterior(net: _Net, layout, loc: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """avg_posterior=True: LOGITS of the guide-mean network (model_bnn.py:206-216)."""
    return net_logits(net, unpack(loc, layout), x)


def check_seeds(seeds, n_samples):
    """model_bnn.py:200-202."""
    if seeds:
        if len(seeds) != n_samples:
            raise ValueError("Number of seeds should match number of samples.")


# --------------------------------------------------------------------------
# a6/a7  loss_gradient(s)                            lossGradients.py:20-68
# --------------------------------------------------------------------------
def loss_gradient_r1(net: _Net, layout, bank: torch.Tensor, image: torch.Tensor,
                     label_onehot: torch.Tensor, n_samples: int) -> torch.Tensor:
    """Reference loop order for ONE image (lossGradients.py:20-40): per sample a
    batch-1 forward through BNN.forward(n_samples=1, seeds=[i]), CE applied to the
    returned PROBABILITIES (:34), full autograd backward, mean of the S gradients."""
    image = image.unsqueeze(0)
    label = label_onehot.argmax(-1).unsqueeze(0)
    grads = []
    for i in range(n_samples):
        x_copy = copy.deepcopy(image)
        x_copy.requires_grad = True
        output = bnn_forward(net, layout, bank, x_copy, [i])
        loss = torch.nn.CrossEntropyLoss()(output, label)
        loss.backward()
        grads.append(copy.deepcopy(x_copy.grad.data[0]))
    return torch.stack(grads, 0).mean(0)


def loss_gradients_r1(net, layout, bank, images, labels_onehot, n_samples) -> np.ndarray:
    """lossGradients.py:52-67 without the pickle side effect: stack, squeeze, numpy."""
    out = [loss_gradient_r1(net, layout, bank, images[i], labels_onehot[i], n_samples)
           for i in range(len(images))]
    return torch.stack(out).cpu().detach().numpy().squeeze()


def loss_gradients_reference_order(net: _Net, layout, loc: torch.Tensor, rho: torch.Tensor,
                                   images: torch.Tensor, labels_onehot: torch.Tensor, n_samples: int):
    """The reference's full CPU cost model for an SVI BNN, used as the timed CPU baseline:
    for every image, for every sample i: set_rng_seed(i), the guide's two wasted randn_like per key,
    a reparameterised draw of ALL weights from (loc, softplus(scale)) that stay attached to the
    autograd graph (the params require grad in Pyro), deepcopy of the base net by random_module, a
    batch-1 forward, CE on probabilities and a FULL backward (weight-side gradients included)
    -- lossGradients.py:29-38,56-60 + model_bnn.py:121-136,222-226."""
    loc = loc.clone().requires_grad_(True)
    rho = rho.clone().requires_grad_(True)
    out = []
    for n in range(len(images)):
        image = images[n].unsqueeze(0)
        label = labels_onehot[n].argmax(-1).unsqueeze(0)
        grads = []
        for i in range(n_samples):
            x_copy = copy.deepcopy(image)
            x_copy.requires_grad = True
            torch.manual_seed(i)
            weights, off = {}, 0
            for key, shp in layout:
                torch.randn(shp)
                torch.randn(shp)
            net_copy = copy.deepcopy(net)
            for key, shp in layout:
                cnt = int(np.prod(shp))
                mu = loc[off:off + cnt].reshape(shp)
                sd = softplus(rho[off:off + cnt].reshape(shp))
                weights[key] = torch.distributions.Normal(mu, sd).rsample()
                off += cnt
            output = nnf.softmax(net_logits(net_copy, weights, x_copy), dim=-1)
            output = torch.stack([output]).mean(0)
            loss = torch.nn.CrossEntropyLoss()(output, label)
            loc.grad = rho.grad = None
            loss.backward()
            grads.append(copy.deepcopy(x_copy.grad.data[0]))
        out.append(torch.stack(grads, 0).mean(0))
    return torch.stack(out)


def expected_loss_gradients(net: _Net, layout, bank: torch.Tensor, x: torch.Tensor,
                            labels: torch.Tensor, sample_ids: Sequence[int],
                            dtype=torch.float32) -> torch.Tensor:
    """R2: the same quantity evaluated sample-major and batched.  Images are
    independent and the reference batch is 1 (no 1/B), so CE uses reduction='sum'
    (SURVEY.md section 3.1).  Returns [B, *input_shape]."""
    x = x.to(dtype)
    acc = torch.zeros_like(x)
    for s in sample_ids:
        w = {k: v.to(dtype) for k, v in unpack(bank[int(s)], layout).items()}
        xs = x.clone().requires_grad_(True)
        probs = nnf.softmax(net_logits(net, w, xs), dim=-1)
        loss = nnf.cross_entropy(probs, labels, reduction="sum")
        (g,) = torch.autograd.grad(loss, xs)
        acc += g
    return acc / float(len(sample_ids))


# --------------------------------------------------------------------------
# a8/a9/a10  attacks                                 adversarialAttacks.py:69-143
# --------------------------------------------------------------------------
def attack_gradient(net: _Net, layout, bank: torch.Tensor, x: torch.Tensor, labels: torch.Tensor,
                    sample_ids: Sequence[int], dtype=torch.float32) -> torch.Tensor:
    """d/dx CE(mean_s softmax f_s(x), y): the gradient-of-the-mean definition used by
    fgsm/pgd (adversarialAttacks.py:74-78, :97-101; model_bnn.py:257)."""
    xs = x.to(dtype).clone().requires_grad_(True)
    preds = []
    for s in sample_ids:
        w = {k: v.to(dtype) for k, v in unpack(bank[int(s)], layout).items()}
        preds.append(nnf.softmax(net_logits(net, w, xs), dim=-1))
    out = torch.stack(preds).mean(0)
    loss = nnf.cross_entropy(out, labels, reduction="sum")
    (g,) = torch.autograd.grad(loss, xs)
    return g


def attack_gradient_avg_posterior(net, layout, loc, x, labels, dtype=torch.float32):
    """avg_posterior=True: CE on the LOGITS of the mean-weight net (model_bnn.py:206-216)."""
    xs = x.to(dtype).clone().requires_grad_(True)
    w = {k: v.to(dtype) for k, v in unpack(loc, layout).items()}
    loss = nnf.cross_entropy(net_logits(net, w, xs), labels, reduction="sum")
    (g,) = torch.autograd.grad(loss, xs)
    return g


def fgsm_step(image, grad, epsilon):
    """adversarialAttacks.py:81-82."""
    return torch.clamp(image + epsilon * grad.sign(), 0, 1)


def pgd_hyper(image: torch.Tensor, hyperparams: Optional[dict]):
    """adversarialAttacks.py:88-91 -- note alpha = 2/image.max() per IMAGE."""
    if hyperparams is not None:
        return hyperparams["epsilon"], 2 / image.max(), 40
    return 0.5, 2 / 225, 40


def pgd_step(image, original, grad, alpha, epsilon):
    """adversarialAttacks.py:103-105."""
    perturbed = image + alpha * grad.sign()
    eta = torch.clamp(perturbed - original, min=-epsilon, max=epsilon)
    return torch.clamp(original + eta, min=0, max=1)


SampleSchedule = Callable[[int], Sequence[int]]


def fgsm_attack(net, layout, bank, images, labels, sample_schedule: SampleSchedule,
                hyperparams=None, dtype=torch.float32):
    """Batched fgsm_attack (adversarialAttacks.py:69-83); `sample_schedule(call)` names
    the bank rows forward call number `call` uses (HMC: always range(n))."""
    epsilon = hyperparams["epsilon"] if hyperparams is not None else 0.3
    g = attack_gradient(net, layout, bank, images, labels, sample_schedule(0), dtype)
    return fgsm_step(images.to(dtype), g, epsilon)


def pgd_attack(net, layout, bank, images, labels, sample_schedule: SampleSchedule,
               hyperparams=None, iters: Optional[int] = None, dtype=torch.float32):
    """Batched pgd_attack (adversarialAttacks.py:86-108).  alpha is per image
    (2/image.max(), :89) so it is carried as a [B,1,1,1] tensor."""
    images = images.to(dtype)
    if hyperparams is not None:
        epsilon = hyperparams["epsilon"]
        alpha = 2 / images.flatten(1).max(dim=1)[0].reshape(-1, *([1] * (images.dim() - 1)))
        n_it = 40
    else:
        epsilon, alpha, n_it = 0.5, 2 / 225, 40
    if iters is not None:
        n_it = iters
    original = images.clone()
    image = images.clone()
    for i in range(n_it):
        g = attack_gradient(net, layout, bank, image, labels, sample_schedule(i), dtype)
        image = pgd_step(image, original, g, alpha, epsilon).detach()
    return image


# --------------------------------------------------------------------------
# deterministic NN / Ensemble_NN     model_nn.py:126-141, model_ensemble.py:57-67
# --------------------------------------------------------------------------
def ensemble_forward(net: _Net, layout, bank: torch.Tensor, x: torch.Tensor, members: Sequence[int],
                     dtype=torch.float32) -> torch.Tensor:
    """Ensemble_NN.forward: mean over the selected members of the LOGITS (model_ensemble.py:62-66);
    a single member is NN.forward (model_nn.py:126-141)."""
    x = x.to(dtype)
    outs = [net_logits(net, {k: v.to(dtype) for k, v in unpack(bank[int(s)], layout).items()}, x) for s in members]
    return torch.stack(outs, 0).mean(0)


def ensemble_attack_gradient(net, layout, bank, x, labels, members, dtype=torch.float32):
    """d/dx CE(mean_s f_s(x), y), the loss fgsm/pgd differentiate for these nets (adversarialAttacks.py:74-78)."""
    xs = x.to(dtype).clone().requires_grad_(True)
    loss = nnf.cross_entropy(ensemble_forward(net, layout, bank, xs, members, dtype), labels, reduction="sum")
    (g,) = torch.autograd.grad(loss, xs)
    return g


def ensemble_fgsm_attack(net, layout, bank, images, labels, members, hyperparams=None, dtype=torch.float32):
    epsilon = hyperparams["epsilon"] if hyperparams is not None else 0.3
    g = ensemble_attack_gradient(net, layout, bank, images, labels, members, dtype)
    return fgsm_step(images.to(dtype), g, epsilon)


def ensemble_pgd_attack(net, layout, bank, images, labels, members, hyperparams=None, iters=None,
                        dtype=torch.float32):
    images = images.to(dtype)
    if hyperparams is not None:
        epsilon = hyperparams["epsilon"]
        alpha = 2 / images.flatten(1).max(dim=1)[0].reshape(-1, *([1] * (images.dim() - 1)))
    else:
        epsilon, alpha = 0.5, 2 / 225
    n_it = 40 if iters is None else iters
    original, image = images.clone(), images.clone()
    for _ in range(n_it):
        g = ensemble_attack_gradient(net, layout, bank, image, labels, members, dtype)
        image = pgd_step(image, original, g, alpha, epsilon).detach()
    return image


def ensemble_attack_evaluation(net, layout, bank, x_test, x_attack, y_onehot, members, batch_size: int = 128):
    """adversarialAttacks.py:151-198 for a net whose forward returns (mean) logits."""
    labels = y_onehot.argmax(-1)
    outs, correct = [[], []], [0.0, 0.0]
    with torch.no_grad():
        for which, data in enumerate((x_test, x_attack)):
            for b0 in range(0, len(data), batch_size):
                out = ensemble_forward(net, layout, bank, data[b0:b0 + batch_size], members)
                correct[which] += (out.argmax(-1) == labels[b0:b0 + batch_size]).sum().item()
                outs[which].append(out)
        rob = softmax_robustness(torch.cat(outs[0]), torch.cat(outs[1]))
    return 100 * correct[0] / len(x_test), 100 * correct[1] / len(x_test), rob


# --------------------------------------------------------------------------
# a11/a12  evaluation                                adversarialAttacks.py:30-62,151-198
# --------------------------------------------------------------------------
def softmax_difference(original_predictions, adversarial_predictions):
    """adversarialAttacks.py:30-51 (softmax applied AGAIN to whatever forward returned)."""
    original_predictions = nnf.softmax(original_predictions, dim=-1)
    adversarial_predictions = nnf.softmax(adversarial_predictions, dim=-1)
    if len(original_predictions) != len(adversarial_predictions):
        raise ValueError("\nInput arrays should have the same length.")
    softmax_diff = original_predictions - adversarial_predictions
    softmax_diff_norms = softmax_diff.abs().max(dim=-1)[0]
    if softmax_diff_norms.min() < 0. or softmax_diff_norms.max() > 1.:
        raise ValueError("Softmax difference should be in [0,1]")
    return softmax_diff_norms


def softmax_robustness(original_outputs, adversarial_outputs):
    """adversarialAttacks.py:53-62 (without the print)."""
    d = softmax_difference(original_outputs, adversarial_outputs)
    return torch.ones_like(d) - d


def attack_evaluation(net, layout, bank, x_test, x_attack, y_onehot,
                      sample_schedule: SampleSchedule, batch_size: int = 128):
    """adversarialAttacks.py:151-198: batches of 128, first all clean batches then all
    adversarial ones, one forward call each (`sample_schedule(call)` as above),
    integer correct counts -> percentages, then softmax_robustness."""
    call = 0
    labels = y_onehot.argmax(-1)
    outs, correct = [[], []], [0.0, 0.0]
    with torch.no_grad():
        for which, data in enumerate((x_test, x_attack)):
            for b0 in range(0, len(data), batch_size):
                out = bnn_forward(net, layout, bank, data[b0:b0 + batch_size], sample_schedule(call))
                call += 1
                correct[which] += (out.argmax(-1) == labels[b0:b0 + batch_size]).sum().item()
                outs[which].append(out)
        original_accuracy = 100 * correct[0] / len(x_test)
        adversarial_accuracy = 100 * correct[1] / len(x_test)
        rob = softmax_robustness(torch.cat(outs[0]), torch.cat(outs[1]))
    return original_accuracy, adversarial_accuracy, rob


def evaluate(net, layout, bank, x, y_onehot, n_samples, batch_size=128):
    """BNN.evaluate (model_bnn.py:367-391): seeds=range(n_samples) for every batch."""
    correct = 0.0
    with torch.no_grad():
        for b0 in range(0, len(x), batch_size):
            out = bnn_forward(net, layout, bank, x[b0:b0 + batch_size], range(n_samples))
            correct += (out.argmax(-1) == y_onehot[b0:b0 + batch_size].argmax(-1)).sum().item()
    return 100 * correct / len(x)


# --------------------------------------------------------------------------
# synthetic problem generators shared by tests / bench (never by the product)
# --------------------------------------------------------------------------
def scaled_guide_params(layout, seed: int = 1, rho_mean: float = -5.0):
    """Well-scaled random guide: loc ~ N(0, 1/fan_in), rho ~ N(rho_mean, 1)
    (SURVEY.md section 7 hard part 1; the reference's literal init is randn/randn,
    model_bnn.py:125-126, which saturates the softmax)."""
    g = torch.Generator().manual_seed(seed)
    locs, rhos = [], []
    for key, shp in layout:
        n = int(np.prod(shp))
        fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else None
        if fan_in is None:                       # bias: use the previous weight's fan_in
            fan_in = last_fan_in
        last_fan_in = fan_in
        locs.append(torch.randn(n, generator=g) / math.sqrt(fan_in))
        rhos.append(torch.randn(n, generator=g) + rho_mean)
    return torch.cat(locs), torch.cat(rhos)


def synthetic_inputs(n: int, input_shape, n_classes: int, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((n, *input_shape), generator=g)
    y = torch.randint(0, n_classes, (n,), generator=g)
    return x, nnf.one_hot(y, n_classes).float()
