"""Drop-in for the reference's `model_ensemble.Ensemble_NN` (model_ensemble.py:14-108): the members are the rows of one
weight bank on the CUDA engine; `forward` is the mean of the members' LOGITS (one grouped pass over the bank rows,
`rbnn_forward_logits_sum`), and the attacks differentiate it through head LOGITS_UPSTREAM.  Training is out of scope."""
import torch

from .model_nn import NN
from .savedir import TESTS


class Ensemble_NN(NN):

    def __init__(self, dataset_name, hidden_size, activation, architecture,
                 epochs, lr, input_shape, output_size, ensemble_size):
        super(Ensemble_NN, self).__init__(dataset_name, input_shape, output_size,
                                          hidden_size, activation, architecture, lr, epochs)
        self.ensemble_size = ensemble_size
        self.random_seeds = range(0, ensemble_size)
        self.member_name = self.name                  # file stem of the members' weight files (NN.get_name)
        self.name = self.get_name(ensemble_size)
        self.ensemble_models = {}

    def get_name(self, ensemble_size, *args, **kwargs):
        if not isinstance(ensemble_size, int):        # NN.__init__ calls get_name(dataset, hidden, ...) once
            return NN.get_name(self, ensemble_size, *args, **kwargs)
        return str(self.dataset_name) + "_ensemble_hid=" + str(self.hidden_size) + "_act=" + str(self.activation) + \
            "_arch=" + str(self.architecture) + "_size=" + str(ensemble_size)

    def set_members(self, state_dicts):
        """Install the members from state dicts (or a [size, P] tensor), member i = bank row i."""
        if torch.is_tensor(state_dicts):
            rows = state_dicts.detach().float().cpu()
        else:
            rows = torch.stack([self._pack(sd) for sd in state_dicts])
        if rows.shape[0] != self.ensemble_size:
            raise AttributeError("expected %d ensemble members, got %d" % (self.ensemble_size, rows.shape[0]))
        self._install(rows)
        self.ensemble_models = {str(seed): i for i, seed in enumerate(self.random_seeds)}

    def load(self, device, rel_path=TESTS):
        """Reads the members' `<NN name>_weights_<seed>.pt` files from `<name>/weights` (model_ensemble.py:44-55)."""
        self.device = device
        savedir = self.name + "/weights"
        sds = []
        for seed in self.random_seeds:
            path = rel_path + savedir + "/" + self.member_name + "_weights_" + str(seed) + ".pt"
            print("\nLoading: ", path)
            sds.append(torch.load(path, map_location="cpu"))
        self.set_members(sds)

    def save(self, seed=None, *args, **kwargs):
        raise NotImplementedError("ensemble members are trained and saved by the reference")

    def _members(self, n_samples):
        if n_samples is not None and n_samples > self.ensemble_size:
            raise ValueError("Maximum number of samples allowed is ", self.ensemble_size)
        return self.ensemble_size if n_samples is None else int(n_samples)

    def forward(self, inputs, n_samples=None, *args, **kwargs):
        """Mean of the first `n_samples` members' logits (model_ensemble.py:57-67)."""
        return NN.forward(self, inputs, n_samples=n_samples)

    __call__ = forward

    def evaluate(self, test_loader, device, n_samples, *args, **kwargs):
        if n_samples > self.ensemble_size:
            raise ValueError("Maximum number of samples allowed is ", self.ensemble_size)
        return NN.evaluate(self, test_loader, device, n_samples=n_samples)
