"""Drop-in for the reference's `adversarialAttacks.py` hot path (adversarialAttacks.py:30-198).

Same function names, argument names, defaults, return types, prints, exceptions and
pickle side effect.  For a `BNN` the per-image loop (adversarialAttacks.py:118-131) becomes
one batched device pass per gradient evaluation:

  phase 1  sum_s softmax(f_s(x))  over this rank's bank rows  -> allreduce [B,C] -> pbar
  phase 2  sum_s p_s*(g-<p_s,g>) back-propagated to x, g = softmax(pbar)-e_y -> allreduce [B,D]
  update   fused sign / step / project / clip kernel

with no host synchronisation between iterations of PGD.  Other network types
(duck-typed `forward` returning logits) take the reference's generic autograd route.
"""
import os
import random

import torch
import torch.nn.functional as nnf

from . import dist as rdist
from . import engine as E
from ._lib import HEAD_GRAD_OF_MEAN, HEAD_LOGITS_CE
from .model_bnn import BNN
from .savedir import TESTS
from .utils import load_from_pickle, save_to_pickle

DEBUG = False
PGD_ITERS = 40         # hard-coded in the reference (adversarialAttacks.py:89,91)
EVAL_BATCH = 128       # adversarialAttacks.py:170-171


#######################
# robustness measures #
#######################

def softmax_difference(original_predictions, adversarial_predictions):
    """L-inf norm over classes of softmax(orig) - softmax(adv), per point (adversarialAttacks.py:30-51)."""
    if len(original_predictions) != len(adversarial_predictions):
        raise ValueError("\nInput arrays should have the same length.")
    if original_predictions.is_cuda:
        rob, mm = E.softmax_robustness(original_predictions.float().contiguous(),
                                       adversarial_predictions.float().contiguous())
        diff = 1.0 - rob
        lo, hi = mm.tolist()
    else:
        a = nnf.softmax(original_predictions, dim=-1)
        b = nnf.softmax(adversarial_predictions, dim=-1)
        diff = (a - b).abs().max(dim=-1)[0]
        lo, hi = float(diff.min()), float(diff.max())
    if lo < 0. or hi > 1.:
        raise ValueError("Softmax difference should be in [0,1]")
    return diff


def softmax_robustness(original_outputs, adversarial_outputs):
    """1 - softmax_difference, per point (adversarialAttacks.py:53-62)."""
    if len(original_outputs) != len(adversarial_outputs):
        raise ValueError("\nInput arrays should have the same length.")
    if original_outputs.is_cuda:
        robustness, mm = E.softmax_robustness(original_outputs.float().contiguous(),
                                              adversarial_outputs.float().contiguous())
        lo, hi = mm.tolist()
        if lo < 0. or hi > 1.:
            raise ValueError("Softmax difference should be in [0,1]")
    else:
        d = softmax_difference(original_outputs, adversarial_outputs)
        robustness = torch.ones_like(d) - d
    print(f"avg softmax robustness = {robustness.mean().item():.2f}")
    return robustness


#######################
# adversarial attacks #
#######################

def _bnn_input_grad(net, x, labels_i32, n_samples, avg_posterior):
    """d/dx CE(BNN.forward(x), y) summed over the batch, [B, D] on the device (not yet sign()ed)."""
    eng = net.engine()
    if avg_posterior and net.inference == "svi":
        row = net._pin_cap
        net._scratch_generation += 1
        eng.reserve(row + 1)
        eng.upload(net._loc.reshape(1, -1), row)
        return eng.input_grad_sum(HEAD_LOGITS_CE, x, labels_i32, row, row + 1)
    n = 10 if n_samples is None else int(n_samples)      # BNN.forward's default n_samples=10 (model_bnn.py:198)
    rows, _ = net._rows(n, None)
    # the reference evaluates net.forward once and differentiates it (adversarialAttacks.py:74-78): the forward keeps
    # the per-sample logits and LeakyReLU masks, the mean prediction is all-reduced, and the gradient pass reuses them
    pbar = eng.forward_probs_sum(x, rows[0], rows[1], keep=True)
    kept = eng.keep_valid
    rdist.allreduce_sum_(pbar)
    pbar *= 1.0 / n
    if kept:
        g = eng.input_grad_sum_kept(HEAD_GRAD_OF_MEAN, labels_i32, pbar=pbar).reshape(x.shape[0], -1)
    else:                       # engines without a kept route (FP32 CUDA-core engine, fc2, conv): second forward inside
        g = eng.input_grad_sum(HEAD_GRAD_OF_MEAN, x, labels_i32, rows[0], rows[1], pbar=pbar)
    rdist.allreduce_sum_(g)
    return g          # the 1/S factor does not change sign(g)


def _generic_input_grad(net, image, label, n_samples, avg_posterior):
    """Non-BNN networks (adversarialAttacks.py:73-79): the engine-backed `NN` / `Ensemble_NN` drop-ins answer with the
    CUDA input-gradient pass (`input_grad`), anything else takes the reference's autograd route."""
    if hasattr(net, "input_grad"):
        return net.input_grad(image, label, n_samples).reshape(image.shape)
    image = image.detach().clone()
    image.requires_grad = True
    output = net.forward(inputs=image, n_samples=n_samples, avg_posterior=avg_posterior)
    loss = torch.nn.CrossEntropyLoss(reduction="sum")(output, label)
    net.zero_grad()
    loss.backward()
    return image.grad.data


def _to_net_device(net, image, label):
    """Engine-backed deterministic nets compute on their CUDA device: move the attack's tensors there."""
    if hasattr(net, "input_grad"):
        dev = net.engine().device
        return torch.as_tensor(image).to(dev), torch.as_tensor(label).to(dev)
    return image, label


def _prep(net, image, label):
    eng = net.engine()
    x = torch.as_tensor(image).detach().to(device=eng.device, dtype=torch.float32).contiguous()
    y = torch.as_tensor(label).to(device=eng.device).reshape(-1).to(torch.int32).contiguous()
    return x, y


def fgsm_attack(net, image, label, hyperparams=None, n_samples=None, avg_posterior=False):
    """clamp(x + eps*sign(grad), 0, 1), eps = hyperparams["epsilon"] or 0.3 (adversarialAttacks.py:69-83).
    `image` is [B, ch, h, w] (the reference passes B=1), `label` [B] class indices."""
    epsilon = hyperparams["epsilon"] if hyperparams is not None else 0.3
    if not isinstance(net, BNN):
        image, label = _to_net_device(net, image, label)
        grad = _generic_input_grad(net, image, label, n_samples, avg_posterior)
        return torch.clamp(image.detach() + epsilon * grad.sign(), 0, 1)
    x, y = _prep(net, image, label)
    g = _bnn_input_grad(net, x, y, n_samples, avg_posterior)
    return net.engine().fgsm_step(x.reshape(-1), g.reshape(-1), float(epsilon)).reshape(x.shape)


def pgd_attack(net, image, label, hyperparams=None, n_samples=None, avg_posterior=False, iters=PGD_ITERS):
    """Projected gradient ascent in the L-inf ball around the original image
    (adversarialAttacks.py:86-108): (eps, alpha) = (hyperparams eps, 2/image.max()) or (0.5, 2/225),
    `iters`=40 as hard-coded upstream; alpha is per image."""
    if not isinstance(net, BNN):
        image, label = _to_net_device(net, image, label)
        if hyperparams is not None:          # alpha = 2 / image.max() of EACH image (the reference attacks one at a time, :89)
            epsilon = hyperparams["epsilon"]
            alpha = 2 / image.flatten(1).max(dim=1)[0].reshape(-1, *([1] * (image.dim() - 1)))
        else:
            epsilon, alpha = 0.5, 2 / 225
        original_image = image.detach().clone()
        image = original_image.clone()
        for _ in range(iters):
            grad = _generic_input_grad(net, image, label, n_samples, avg_posterior)
            perturbed = image + alpha * grad.sign()
            eta = torch.clamp(perturbed - original_image, min=-epsilon, max=epsilon)
            image = torch.clamp(original_image + eta, min=0, max=1).detach()
        return image
    x0, y = _prep(net, image, label)
    B = x0.shape[0]
    if hyperparams is not None:
        epsilon = float(hyperparams["epsilon"])
        alpha = net.engine().pgd_alpha(x0.reshape(B, -1))
    else:
        epsilon = 0.5
        alpha = torch.full((B,), 2 / 225, dtype=torch.float32, device=x0.device)
    return _pgd_loop(net, x0, x0, y, alpha, epsilon, n_samples, avg_posterior, iters)


PGD_GRAPH = os.environ.get("RBNN_PGD_GRAPH", "1") != "0"    # replay PGD iterations as ONE captured CUDA graph
PGD_GRAPH_MIN_ITERS = 4


def _pgd_iteration(net, x, x0, y, alpha, epsilon, n_samples, avg_posterior):
    B = x0.shape[0]
    g = _bnn_input_grad(net, x, y, n_samples, avg_posterior)
    return net.engine().pgd_step(x.reshape(B, -1), x0.reshape(B, -1), g.reshape(B, -1), alpha, epsilon).reshape(x0.shape)


def _pgd_loop(net, x, x0, y, alpha, epsilon, n_samples, avg_posterior, iters):
    """`iters` PGD updates of `x` inside the eps-ball around `x0` (adversarialAttacks.py:95-105); nothing is read back
    between iterations.  On one device (or inside `dist.replicated()`, where no collective is needed) an iteration --
    K-sample of the fresh posterior draws, kept forward, loss head, input-gradient GEMM, sign / step / project / clip --
    is captured ONCE as a CUDA graph and replayed: one launch per iteration instead of ~20, and the fresh draws of
    every replay come from a device-resident sample counter, so the result is bit-identical to the eager loop."""
    if _graph_eligible(net, x, iters, avg_posterior):
        return _pgd_loop_graph(net, x, x0, y, alpha, epsilon, n_samples, iters)
    for _ in range(iters):
        x = _pgd_iteration(net, x, x0, y, alpha, epsilon, n_samples, avg_posterior)
    return x


def _graph_eligible(net, x, iters, avg_posterior):
    if not PGD_GRAPH or iters < PGD_GRAPH_MIN_ITERS or avg_posterior or not x.is_cuda:
        return False
    eng = net.engine()
    if not hasattr(eng, "alloc_epoch") or rdist.world()[1] != 1:       # sample sharding all-reduces inside an iteration
        return False
    if getattr(net, "_pgd_graph_disabled", False):
        return False
    return not torch.cuda.is_current_stream_capturing()


class _PgdGraph(object):
    __slots__ = ("graph", "x", "x0", "y", "alpha", "offset", "counter0", "epoch", "fresh")


def _pgd_loop_graph(net, x, x0, y, alpha, epsilon, n_samples, iters):
    eng = net.engine()
    n = 10 if n_samples is None else int(n_samples)
    fresh = net._bank_host is None                    # SVI: every iteration draws n new samples (model_bnn.py:230-232)
    key = (tuple(x0.shape), n, float(epsilon), eng.precision, net._posterior_generation, net._fresh_key, net._pin_cap,
           str(x0.device))
    ent = net._pgd_graphs.get(key)
    done = 0
    if ent is None or ent.epoch != eng.alloc_epoch:
        # one eager iteration first: it sizes every workspace of the engine (no allocation may happen during capture)
        x = _pgd_iteration(net, x, x0, y, alpha, epsilon, n_samples, False)
        done = 1
        ent = _PgdGraph()
        ent.x, ent.x0, ent.y, ent.alpha = x.clone(), x0.clone(), y.clone(), alpha.clone()
        ent.offset = torch.zeros((1,), dtype=torch.int64, device=x0.device)
        ent.fresh = fresh
        ent.counter0 = net._fresh_counter
        ent.graph = torch.cuda.CUDAGraph()
        gen0 = net._scratch_generation
        net._graph_offset = ent.offset if fresh else None
        failed = None
        try:
            with torch.cuda.graph(ent.graph):
                xn = _pgd_iteration(net, ent.x, ent.x0, ent.y, ent.alpha, epsilon, n_samples, False)
                ent.x.copy_(xn)
                if fresh:
                    ent.offset.add_(n)
        except Exception as e:                       # capture refused (driver / library state): stay on the eager loop
            failed = e
        finally:
            net._graph_offset = None
            net._fresh_counter = ent.counter0        # the captured iteration has not run
            net._scratch_generation = gen0 + 1
        if failed is not None:
            import warnings
            warnings.warn("robustbnns_b200: CUDA-graph capture of the PGD iteration failed (%s); using the eager loop" % failed)
            net._pgd_graph_disabled = True
            for _ in range(iters - done):
                x = _pgd_iteration(net, x, x0, y, alpha, epsilon, n_samples, False)
            return x
        ent.epoch = eng.alloc_epoch
        if len(net._pgd_graphs) >= 8:
            net._pgd_graphs.clear()
        net._pgd_graphs[key] = ent
    ent.x.copy_(x)
    ent.x0.copy_(x0)
    ent.y.copy_(y)
    ent.alpha.copy_(alpha)
    if ent.fresh:
        ent.offset.fill_(net._fresh_counter - ent.counter0)
    for _ in range(iters - done):
        ent.graph.replay()
    if ent.fresh:
        net._fresh_counter += n * (iters - done)
    net._scratch_generation += 1
    return ent.x.clone()


ATTACK_BATCH = 8192     # images per device pass


def attack_all(net, x_test, labels, method, device=None, hyperparams=None, n_samples=None, avg_posterior=False,
               iters=PGD_ITERS):
    """The device part of `attack`: adversarial examples [N, ch, h, w] for all test points (`labels` are class indices).
    Multi-GPU (torch.distributed initialised, BNN with attack_sharding == "inputs", the default): every rank draws the
    same posterior samples (global Philox indices), attacks its own block of the test points without any collective,
    and the blocks are all-gathered at the end -- identical to the single-GPU result."""
    rank, world = rdist.real_world()
    shard_inputs = isinstance(net, BNN) and world > 1 and getattr(net, "attack_sharding", "inputs") == "inputs"
    n_total = len(x_test)
    lo, hi, per = 0, n_total, n_total
    if shard_inputs:
        per = (n_total + world - 1) // world
        lo, hi = min(n_total, rank * per), min(n_total, (rank + 1) * per)

    # Unseeded SVI: the reference attacks one image per call, so every image sees its OWN fresh weight draws
    # (adversarialAttacks.py:118-131 + model_bnn.py:230-232); one device pass shares the draws of an evaluation among its
    # images (equal per-image marginals, correlated images).  `net.fresh_draws = "per_image"` restores the reference's
    # independence at its price: one device pass -- and n_samples new weight draws per gradient evaluation -- per image.
    per_image = (isinstance(net, BNN) and getattr(net, "fresh_draws", "per_pass") == "per_image"
                 and getattr(net, "_bank_host", None) is None and not avg_posterior)
    batch = 1 if per_image else ATTACK_BATCH

    def run(lo, hi):
        out = []
        for b0 in range(lo, hi, batch):
            b1 = min(hi, b0 + batch)
            image = torch.as_tensor(x_test[b0:b1])
            label = labels[b0:b1]
            if not isinstance(net, BNN):
                image, label = image.to(device), label.to(device)
            if method == "fgsm":
                perturbed_image = fgsm_attack(net=net, image=image, label=label, hyperparams=hyperparams,
                                              n_samples=n_samples, avg_posterior=avg_posterior)
            elif method == "pgd":
                perturbed_image = pgd_attack(net=net, image=image, label=label, hyperparams=hyperparams,
                                             n_samples=n_samples, avg_posterior=avg_posterior, iters=iters)
            out.append(perturbed_image)
        return out

    if not shard_inputs:
        return torch.cat(run(0, n_total))
    # The fresh-draw counter advances once per gradient evaluation of a device pass.  Ranks attack blocks of different
    # sizes (possibly none), so afterwards every rank is put where ONE process attacking all N points would be:
    # later sample-sharded calls then draw from the same global Philox indices on every rank.
    counter0 = getattr(net, "_fresh_counter", 0)
    try:
        with rdist.replicated():
            net._replace_rows()                      # every rank holds all the samples while it attacks its block
            parts = run(lo, hi)
            advance = (net._fresh_counter - counter0) if hi > lo else None
    finally:
        net._replace_rows()                          # back to sample sharding, also when the attack raised
    if hasattr(net, "_fresh_counter"):
        passes_local = max(1, -(-(hi - lo) // batch)) if hi > lo else 0
        passes_all = -(-n_total // batch)
        if advance is None:                          # this rank had no test points: learn the per-pass advance from rank 0
            advance, passes_local = 0, 1
        per_pass = torch.tensor([advance // max(1, passes_local)], dtype=torch.int64, device=net.engine().device)
        if world > 1:
            import torch.distributed as dist
            dist.broadcast(per_pass, src=0)
        net._fresh_counter = counter0 + int(per_pass.item()) * passes_all
    eng = net.engine()
    shape = tuple(torch.as_tensor(x_test).shape[1:])
    local = torch.cat(parts) if parts else torch.zeros((0,) + shape, dtype=torch.float32, device=eng.device)
    return rdist.all_gather_rows(local.reshape((-1,) + shape), per, n_total)


def attack(net, x_test, y_test, dataset_name, device, method, filename, savedir=None,
           hyperparams=None, n_samples=None, avg_posterior=False):
    """All test points at once (adversarialAttacks.py:111-143); returns [N, ch, h, w] on the device (see `attack_all`
    for the multi-GPU form).  The result pickle is written by rank 0 only."""
    print(f"\nProducing {method} attacks on {dataset_name}:")
    labels = torch.as_tensor(y_test).argmax(-1)
    adversarial_attack = attack_all(net, x_test, labels, method, device=device, hyperparams=hyperparams,
                                    n_samples=n_samples, avg_posterior=avg_posterior)
    path = TESTS + filename + "/" if savedir is None else TESTS + savedir + "/"
    name = filename + "_" + str(method)
    # (the reference also writes two PNG grids here through matplotlib, utils.py:276-290: plotting is out of scope)
    name = name + "_attackSamp=" + str(n_samples) + "_attack.pkl" if n_samples else name + "_attack.pkl"
    rank, world = rdist.real_world()
    if rank == 0:
        save_to_pickle(data=adversarial_attack, path=path, filename=name)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()                               # nobody reads the file before rank 0 has written it
    return adversarial_attack


def load_attack(method, filename, savedir=None, n_samples=None, rel_path=TESTS):
    path = TESTS + filename + "/" if savedir is None else TESTS + savedir + "/"
    name = filename + "_" + str(method)
    name = name + "_attackSamp=" + str(n_samples) + "_attack.pkl" if n_samples else name + "_attack.pkl"
    return load_from_pickle(path=path + name)


def attack_evaluation(net, x_test, x_attack, y_test, device, n_samples=None, batch_size=EVAL_BATCH):
    """(original accuracy %, adversarial accuracy %, softmax robustness [N]) (adversarialAttacks.py:151-198)."""
    print(f"\nEvaluating against the attacks", end="")
    if n_samples:
        print(f" with {n_samples} defence samples")
    random.seed(0)
    if not isinstance(net, BNN):
        x_test, x_attack, y_test = x_test.to(device), x_attack.to(device), y_test.to(device)
        outs, correct = [[], []], [0.0, 0.0]
        with torch.no_grad():
            for which, data in enumerate((x_test, x_attack)):
                for b0 in range(0, len(data), batch_size):
                    out = net.forward(data[b0:b0 + batch_size], n_samples)
                    correct[which] += (out.argmax(-1) == y_test[b0:b0 + batch_size].argmax(-1)).sum().item()
                    outs[which].append(out)
        original_accuracy = 100 * correct[0] / len(x_test)
        adversarial_accuracy = 100 * correct[1] / len(x_test)
        print(f"\ntest accuracy = {original_accuracy}\tadversarial accuracy = {adversarial_accuracy}", end="\t")
        return original_accuracy, adversarial_accuracy, softmax_robustness(torch.cat(outs[0]), torch.cat(outs[1]))

    net.reseed(0)                                           # pyro.set_rng_seed(0), adversarialAttacks.py:161
    eng = net.engine()
    x_test = torch.as_tensor(x_test).to(device=eng.device, dtype=torch.float32)
    x_attack = torch.as_tensor(x_attack).to(device=eng.device, dtype=torch.float32)
    labels = torch.as_tensor(y_test).to(eng.device).argmax(-1).to(torch.int32).contiguous()
    n = 10 if n_samples is None else int(n_samples)
    counters = torch.zeros((2,), dtype=torch.int64, device=eng.device)
    outs = [[], []]
    with torch.no_grad():
        for which, data in enumerate((x_test, x_attack)):
            for b0 in range(0, len(data), batch_size):
                out = net.forward(data[b0:b0 + batch_size], n)
                eng.count_correct(out, labels[b0:b0 + batch_size], counters[which:which + 1])
                outs[which].append(out)
    original_correct, adversarial_correct = (float(v) for v in counters.tolist())
    original_accuracy = 100 * original_correct / len(x_test)
    adversarial_accuracy = 100 * adversarial_correct / len(x_test)
    print(f"\ntest accuracy = {original_accuracy}\tadversarial accuracy = {adversarial_accuracy}", end="\t")
    softmax_rob = softmax_robustness(torch.cat(outs[0]), torch.cat(outs[1]))
    return original_accuracy, adversarial_accuracy, softmax_rob
