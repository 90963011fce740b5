"""Drop-in for the gradient / attack half of the reference's `grid_search_halfMoons.py` (SURVEY.md section 8f rank 4,
BASELINE configs[4]): the half-moons over-parametrisation sweep.  `MoonsBNN` (grid_search_halfMoons.py:18-24),
`serial_compute_grads` (:80-99, :133-153) and `grid_attack` (:118-131, :155-176) keep their names and arguments and run
every model of the grid through the CUDA path; posterior training (`_train`, :30-60) stays with the reference -- the
trained posteriors are read from the reference's weight files (`BNN.load`).  `serial_compute_grads` handles one model
after the other (under torchrun the posterior samples of each model are sharded over the ranks as for any other BNN);
`parallel_compute_grads` / `parallel_grid_attack` -- joblib fan-outs over CPU processes upstream (:80-89, :122-131) --
are the batched form: the MODELS are dealt out to the ranks, everything is enqueued without reading back, one
synchronisation at the end (`sweep_expected_loss_gradients`)."""
import itertools

import torch

from . import dist as rdist
from .adversarialAttacks import attack, attack_all
from .lossGradients import expected_loss_gradients, loss_gradients, save_loss_gradients
from .model_bnn import BNN
from .savedir import TESTS
from .utils import data_loaders, load_dataset, save_to_pickle


class MoonsBNN(BNN):

    def __init__(self, hidden_size, activation, architecture, inference,
                 epochs, lr, n_samples, warmup, n_inputs, input_shape, output_size, engine=None):
        super(MoonsBNN, self).__init__("half_moons", hidden_size, activation, architecture,
                                       inference, epochs, lr, n_samples, warmup, input_shape, output_size,
                                       step_size=0.001, engine=engine)
        self.name = self.get_name(n_inputs)


def _compute_grads(hidden_size, activation, architecture, inference,
                   epochs, lr, n_samples, warmup, n_inputs, posterior_samples, rel_path, test_points, device):
    _, test_loader, inp_shape, out_size = \
        data_loaders(dataset_name="half_moons", batch_size=32, n_inputs=test_points, shuffle=True)
    bnn = MoonsBNN(hidden_size, activation, architecture, inference,
                   epochs, lr, n_samples, warmup, n_inputs, inp_shape, out_size)
    bnn.load(device=device, rel_path=rel_path)
    return loss_gradients(net=bnn, n_samples=posterior_samples, savedir=bnn.name + "/",
                          data_loader=test_loader, device=device, filename=bnn.name)


def serial_compute_grads(hidden_size, activation, architecture, inference,
                         epochs, lr, n_samples, warmup, n_inputs, posterior_samples,
                         rel_path, test_points):
    combinations = list(itertools.product(hidden_size, activation, architecture, inference,
                                          epochs, lr, n_samples, warmup, n_inputs, posterior_samples))
    for init in combinations:
        _compute_grads(*init, rel_path, test_points, "cuda")


def sweep_expected_loss_gradients(nets, images, labels, n_samples):
    """Expected loss gradients of MANY models in one device pass: `nets[m]` on `images[m]` / `labels[m]` (class indices)
    with `n_samples[m]` posterior samples.  The evaluations are enqueued back to back on the current stream, every
    result is copied into pinned host memory asynchronously, and the host waits ONCE at the end -- the sweep's models are
    small (half moons: 2-H-H-2), so what the reference's per-model loop spends between models (process start, loader,
    .cpu() round trips; grid_search_halfMoons.py:66-99) would otherwise dominate.  Returns a list of host tensors."""
    outs = []
    for net, x, y, n in zip(nets, images, labels, n_samples):
        g = expected_loss_gradients(net, x, y, n)
        if g.is_cuda:
            h = torch.empty(g.shape, dtype=g.dtype, pin_memory=True)
            h.copy_(g, non_blocking=True)
        else:
            h = g
        outs.append(h)
    if torch.cuda.is_available():
        torch.cuda.current_stream().synchronize()
    return outs


def parallel_compute_grads(hidden_size, activation, architecture, inference,
                           epochs, lr, n_samples, warmup, n_inputs, posterior_samples,
                           rel_path, test_points, device="cuda"):
    """The whole grid in one device pass (BASELINE configs[4]).  Model m of the itertools product belongs to rank
    m % world_size: the models are independent objects, so there is no data-path collective -- every rank keeps ALL
    posterior samples of its own models (`dist.replicated()`), runs `sweep_expected_loss_gradients` over them and writes
    their pickles (the reference's file names, one writer per file)."""
    combinations = list(itertools.product(hidden_size, activation, architecture, inference,
                                          epochs, lr, n_samples, warmup, n_inputs, posterior_samples))
    rank, world = rdist.real_world()
    mine = combinations[rank::world]
    nets, xs, ys, ns = [], [], [], []
    with rdist.replicated():
        for init in mine:
            _, test_loader, inp_shape, out_size = \
                data_loaders(dataset_name="half_moons", batch_size=32, n_inputs=test_points, shuffle=True)
            bnn = MoonsBNN(*init[:-1], inp_shape, out_size)
            bnn.load(device=device, rel_path=rel_path)
            xb, yb = zip(*[(torch.as_tensor(x), torch.as_tensor(y).argmax(-1)) for x, y in test_loader])
            nets.append(bnn)
            xs.append(torch.cat(xb))
            ys.append(torch.cat(yb))
            ns.append(init[-1])
        grads = sweep_expected_loss_gradients(nets, xs, ys, ns)
    out = []
    for bnn, g, n in zip(nets, grads, ns):
        print(f"\n === Loss gradients on {len(g)} input images:")
        print(f"\nmin = {g.min():.4f} \t max = {g.max():.4f}")
        arr = g.detach().numpy().squeeze()
        save_loss_gradients(arr, n, bnn.name, bnn.name + "/")
        out.append(arr)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return out


def grid_attack(method, hidden_size, activation, architecture, inference, epochs, lr,
                n_samples, warmup, n_inputs, posterior_samples, test_points, device="cuda",
                rel_path=TESTS):
    _, _, x_test, y_test, inp_shape, out_size = \
        load_dataset(dataset_name="half_moons", n_inputs=test_points, channels="first")
    x_test = torch.from_numpy(x_test)
    y_test = torch.from_numpy(y_test)
    combinations = list(itertools.product(hidden_size, activation, architecture, inference,
                                          epochs, lr, n_samples, warmup, n_inputs))
    for init in combinations:
        bnn = MoonsBNN(*init, inp_shape, out_size)
        bnn.load(device=device, rel_path=rel_path)
        for p_samp in posterior_samples:
            attack(net=bnn, x_test=x_test, y_test=y_test, dataset_name="half_moons",
                   device=device, method=method, filename=bnn.name, n_samples=p_samp)


def parallel_grid_attack(method, hidden_size, activation, architecture, inference, epochs, lr,
                         n_samples, warmup, n_inputs, posterior_samples, rel_path, test_points, device="cuda"):
    """grid_search_halfMoons.py:122-131 with the models dealt out to the ranks (model m -> rank m % world_size): every
    rank attacks all test points of its own models with all their posterior samples, no collective."""
    _, _, x_test, y_test, inp_shape, out_size = \
        load_dataset(dataset_name="half_moons", n_inputs=test_points, channels="first")
    x_test = torch.from_numpy(x_test)
    y_test = torch.from_numpy(y_test)
    combinations = list(itertools.product(hidden_size, activation, architecture, inference,
                                          epochs, lr, n_samples, warmup, n_inputs, posterior_samples))
    rank, world = rdist.real_world()
    labels = y_test.argmax(-1)
    pending = []
    with rdist.replicated():
        for init in combinations[rank::world]:
            bnn = MoonsBNN(*init[:-1], inp_shape, out_size)
            bnn.attack_sharding = "none"                # this rank owns the model: all test points, all samples
            bnn.load(device=device, rel_path=rel_path)
            print(f"\nProducing {method} attacks on half_moons:")
            adv = attack_all(bnn, x_test, labels, method, device=device, n_samples=init[-1])
            pending.append((bnn, init[-1], adv))
    for bnn, n, adv in pending:                          # the file attack() writes (adversarialAttacks.py:136-141), by the owner rank
        name = bnn.name + "_" + str(method)
        name = name + "_attackSamp=" + str(n) + "_attack.pkl" if n else name + "_attack.pkl"
        save_to_pickle(data=adv, path=TESTS + bnn.name + "/", filename=name)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return [adv for _, _, adv in pending]
