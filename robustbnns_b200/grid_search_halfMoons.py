"""Drop-in for the gradient / attack half of the reference's `grid_search_halfMoons.py` (SURVEY.md section 8f rank 4,
BASELINE configs[4]): the half-moons over-parametrisation sweep.  `MoonsBNN` (grid_search_halfMoons.py:18-24),
`serial_compute_grads` (:80-99, :133-153) and `grid_attack` (:118-131, :155-176) keep their names and arguments and run
every model of the grid through the CUDA path; posterior training (`_train`, :30-60) stays with the reference -- the
trained posteriors are read from the reference's weight files (`BNN.load`).  Under torchrun the posterior samples of
each model are sharded over the ranks as for any other BNN."""
import itertools

import torch

from .adversarialAttacks import attack
from .lossGradients import loss_gradients
from .model_bnn import BNN
from .savedir import TESTS
from .utils import data_loaders, load_dataset


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


parallel_compute_grads = serial_compute_grads      # the reference fans out over CPU processes (joblib); one GPU pass here


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


parallel_grid_attack = grid_attack
