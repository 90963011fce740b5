"""Multi-GPU parity (`-m gpu`, needs >= 2 visible GPUs; skipped otherwise): posterior samples sharded over
NCCL ranks, [B, D] / [B, C] partial sums all-reduced, checked on every rank against

  * the same quantities evaluated on ONE device through the C ABI (all samples, no collective), and
  * the fp64 oracle on the weights the device drew (bank rows downloaded from rank 0's single-device engine).

The sampler is indexed by the GLOBAL sample number, so the bank does not depend on the sharding.
"""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from oracle import oracle as orc
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

SHAPE, C = (1, 28, 28), 10
# (architecture, hidden, inputs, samples, engine): 13 / 5 samples give uneven shards (7 + 6, 3 + 2)
CASES = [("fc", 128, 200, 13, "fp32"), ("fc", 128, 200, 13, "f16x3"), ("conv", 32, 10, 5, "f16x3")]


def _rank_main(rank, world, port, case, q):
    ARCH, HIDDEN, B, S, prec = case
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from robustbnns_b200 import _lib
        from robustbnns_b200 import adversarialAttacks as aa
        from robustbnns_b200 import lossGradients as lg
        from robustbnns_b200.engine import Net
        from robustbnns_b200.model_bnn import BNN
        net = orc.build_net(ARCH, SHAPE, HIDDEN, C)
        layout = orc.param_layout(net)
        loc, rho = orc.scaled_guide_params(layout, seed=5, rho_mean=-4.0)
        x, y = orc.synthetic_inputs(B, SHAPE, C, seed=2)
        labels = y.argmax(-1)
        bnn = BNN("mnist", HIDDEN, "leaky", ARCH, "svi", 1, 0.01, None, None, SHAPE, C)
        bnn.set_guide(loc, rho)
        bnn.set_precision(prec)
        seeds = list(range(S))
        probs = bnn.forward(x, n_samples=S, seeds=seeds)                       # allreduce of [B, C]
        grads = lg.expected_loss_gradients(bnn, x, labels, S)                  # allreduce of [B, D]
        adv = aa.fgsm_attack(bnn, x, labels, hyperparams={"epsilon": 0.1}, n_samples=S)   # both, fresh draws
        assert bnn._pin_rows == len(range(rank, S, world))
        # attack(): input sharding (every rank draws the same samples and attacks its block, one all-gather at the end)
        # against sample sharding (two all-reduces per gradient) on the same fresh-sample stream
        import tempfile
        os.chdir(tempfile.mkdtemp())
        y1h = torch.nn.functional.one_hot(labels, C).float()
        att = {}
        for mode in ("inputs", "samples"):
            bnn.attack_sharding = mode
            bnn.reseed(3)
            att[mode] = aa.attack(net=bnn, x_test=x, y_test=y1h, dataset_name="mnist", device="cuda", method="fgsm",
                                  filename="a", savedir="a", hyperparams={"epsilon": 0.1}, n_samples=S)
        assert att["inputs"].shape == x.shape
        assert float(((att["inputs"] - att["samples"]).abs() > 1e-6).float().mean()) <= 2e-3      # sign ties only
        probs_again = bnn.forward(x, n_samples=S, seeds=seeds)          # the sample sharding is back in place
        # (not bit-identical on F16X3: the re-drawn rows get their guard-band row norms from the fused sampler kernel,
        # whose summation order differs from the stand-alone norm pass)
        assert rel_err(probs_again.cpu(), probs.cpu()) < 1e-6
        # single-device evaluation of the same global samples on this rank's GPU: no collective
        eng = Net(ARCH, SHAPE, HIDDEN, C)
        eng.set_precision(prec)
        eng.sample_diag(loc, rho, bnn.rng_seed, 0, 0, S)
        xd = x.cuda()
        ld = labels.to(torch.int32).cuda()
        g1 = eng.input_grad_sum(_lib.HEAD_MEAN_OF_GRADS, xd, ld, 0, S).reshape(x.shape) / S
        p1 = eng.forward_probs_sum(xd, 0, S) / S
        bank = eng.download(0, S)
        # numpy, not torch tensors: tensors cross the queue as file descriptors served by this process and the parent
        # may fetch them only after it has exited
        npy = lambda t: t.detach().cpu().numpy()  # noqa: E731
        q.put((rank, rel_err(probs.cpu(), p1.cpu()), rel_err(grads.cpu(), g1.cpu()),
               npy(grads), npy(probs), npy(adv), npy(bank) if rank == 0 else None))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s" % (c[0], c[4]))
def test_two_rank_nccl_sharding_equals_single_device_and_oracle(case):
    ARCH, HIDDEN, B, S, prec = case
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    tt = lambda a: None if a is None else torch.from_numpy(a)  # noqa: E731
    res = [(r[0], r[1], r[2], tt(r[3]), tt(r[4]), tt(r[5]), tt(r[6])) for r in res]
    for (_, e_p, e_g, _, _, _, _) in res:
        assert e_p < 1e-5 and e_g < 1e-5          # only the summation order over samples differs
    # every rank holds the same reduced tensors
    assert torch.equal(res[0][3], res[1][3]) and torch.equal(res[0][4], res[1][4]) and torch.equal(res[0][5], res[1][5])
    # and they match the fp64 oracle on the bank the device drew
    net = orc.build_net(ARCH, SHAPE, HIDDEN, C)
    layout = orc.param_layout(net)
    x, y = orc.synthetic_inputs(B, SHAPE, C, seed=2)
    bank = res[0][6]
    ref = orc.expected_loss_gradients(net, layout, bank, x, y.argmax(-1), range(S), dtype=torch.float64)
    assert rel_err(res[0][3], ref) < 1e-4
    assert rel_err(res[0][4], orc.bnn_forward(net, layout, bank, x, range(S)).detach()) < 1e-4
