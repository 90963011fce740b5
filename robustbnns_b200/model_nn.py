"""Host-side mirror of the reference's `model_nn.NN` (model_nn.py:34-239).

`NN` carries the constructor's validation, the architecture definition (the ordered list of
state_dict tensors a posterior sample consists of) and the name -- what the Bayesian hot path needs --
and, once weights are installed (`load` / `load_state_dict`), answers `forward` / `evaluate` and the
attacks through the same CUDA engine as a one-row weight bank (`rbnn_forward_logits_sum`, head
LOGITS_UPSTREAM).  Training (`NN.train`, model_nn.py:176-218) is out of scope and raises.
"""
import math
import os

import torch

from .savedir import TESTS

saved_NNs = {"model_0": {"dataset": "mnist", "hidden_size": 512, "activation": "leaky",
                         "architecture": "conv", "epochs": 5, "lr": 0.01},
             "model_5": {"dataset": "mnist", "hidden_size": 512, "activation": "leaky",
                         "architecture": "fc2", "epochs": 10, "lr": 0.01},
             "model_6": {"dataset": "mnist", "hidden_size": 256, "activation": "leaky",
                         "architecture": "conv", "epochs": 10, "lr": 0.05},
             "model_7": {"dataset": "mnist", "hidden_size": 1024, "activation": "leaky",
                         "architecture": "fc2", "epochs": 5, "lr": 0.02},
             "model_8": {"dataset": "mnist", "hidden_size": 1024, "activation": "leaky",
                         "architecture": "fc2", "epochs": 10, "lr": 0.02},
             "model_9": {"dataset": "mnist", "hidden_size": 1024, "activation": "leaky",
                         "architecture": "conv", "epochs": 10, "lr": 0.01},
             }


def param_layout(architecture, input_shape, hidden_size, output_size):
    """[(state_dict key, shape)] in `basenet.state_dict()` order -- the order BNN.guide
    samples in (model_bnn.py:124) and the row layout of the posterior-sample bank."""
    input_size = input_shape[0] * input_shape[1] * input_shape[2]
    in_channels = input_shape[0]
    H, C = hidden_size, output_size
    if architecture == "fc":            # model_nn.py:77-82
        return [("model.1.weight", (H, input_size)), ("model.1.bias", (H,)),
                ("model.3.weight", (C, H)), ("model.3.bias", (C,))]
    if architecture == "fc2":           # model_nn.py:84-91
        return [("model.1.weight", (H, input_size)), ("model.1.bias", (H,)),
                ("model.3.weight", (H, H)), ("model.3.bias", (H,)),
                ("model.5.weight", (C, H)), ("model.5.bias", (C,))]
    if architecture == "conv":          # model_nn.py:98-106
        return [("model.0.weight", (32, in_channels, 5, 5)), ("model.0.bias", (32,)),
                ("model.3.weight", (H, 32, 5, 5)), ("model.3.bias", (H,)),
                ("model.7.weight", (C, int(H / (4 * 4)) * input_size)), ("model.7.bias", (C,))]
    raise NotImplementedError()         # model_nn.py:123-124 ("conv2" draws a fresh Linear per call upstream)


class NN(object):
    """Architecture descriptor with the reference's constructor signature (model_nn.py:36-54)."""

    def __init__(self, dataset_name, input_shape, output_size, hidden_size, activation,
                 architecture, lr, epochs):
        if math.log(hidden_size, 2).is_integer() is False or hidden_size < 16:
            raise ValueError("\nhidden size should be a power of 2 greater than 16.")
        if activation not in ("relu", "leaky", "sigm", "tanh"):
            raise AssertionError("\nWrong activation name.")
        if architecture == "conv" and dataset_name not in ["mnist", "fashion_mnist"]:
            raise NotImplementedError()
        self.dataset_name = dataset_name
        self.architecture = architecture
        self.hidden_size = hidden_size
        self.output_size = output_size
        self.activation = activation
        self.input_shape = tuple(int(v) for v in input_shape)
        self.lr, self.epochs = lr, epochs
        self.layout = param_layout(architecture, self.input_shape, hidden_size, output_size)
        if activation != "leaky" and architecture == "conv":
            # every saved model uses leaky (model_bnn.py:36-66); the conv kernels fuse LeakyReLU with the pooling layers.
            # relu / sigm / tanh (model_nn.py:66-73) run for fc / fc2, on the FP32 CUDA-core engine
            raise NotImplementedError("the B200 path implements arch conv with activation='leaky' only")
        self.name = self.get_name(dataset_name, hidden_size, activation, architecture, lr, epochs)

    def get_name(self, dataset_name, hidden_size, activation, architecture, lr, epochs):
        return str(dataset_name) + "_nn_hid=" + str(hidden_size) + "_act=" + str(activation) + \
            "_arch=" + str(architecture) + "_ep=" + str(epochs) + "_lr=" + str(lr)

    def state_dict_keys(self):
        return [k for k, _ in self.layout]

    @property
    def n_params(self):
        n = 0
        for _, shp in self.layout:
            m = 1
            for v in shp:
                m *= v
            n += m
        return n

    def to(self, device):
        return self

    def zero_grad(self):
        return None

    def train(self, *args, **kwargs):
        raise NotImplementedError("deterministic-network training is outside the accelerated hot path")

    # ---- weights: one row of an engine bank -------------------------------------------------------
    _engine = None
    _rows_host = None          # [n_members, P] CPU tensor (NN: one row)

    def engine(self):
        if self._engine is None:
            from .engine import Net
            self._engine = Net(self.architecture, self.input_shape, self.hidden_size, self.output_size)
            if self.activation != "leaky":
                self._engine.set_activation(self.activation)
            self._engine.set_best_precision()        # tensor-core engines where the network has one
        return self._engine

    def _pack(self, state_dict):
        flat = torch.cat([torch.as_tensor(state_dict[k]).detach().reshape(-1).float().cpu() for k in self.state_dict_keys()])
        if flat.numel() != self.n_params:
            raise ValueError("expected %d parameters, got %d" % (self.n_params, flat.numel()))
        return flat

    def _install(self, rows):
        self._rows_host = rows.contiguous()
        self.engine().upload(self._rows_host, 0)

    def load_state_dict(self, state_dict):
        self._install(self._pack(state_dict).unsqueeze(0))

    def state_dict(self):
        if self._rows_host is None:
            raise RuntimeError("no weights installed: call load() / load_state_dict() first")
        out, off = {}, 0
        for key, shp in self.layout:
            n = 1
            for v in shp:
                n *= v
            out[key] = self._rows_host[0, off:off + n].reshape(shp).clone()
            off += n
        return out

    def _weights_file(self, savedir, seed, rel_path):
        directory = self.name if savedir is None else savedir
        filename = self.name + "_weights.pt" if seed is None else self.name + "_weights_" + str(seed) + ".pt"
        return rel_path + directory + "/" + filename

    def save(self, savedir=None, seed=None):
        """torch.save(state_dict) under the reference's file name (model_nn.py:143-151)."""
        path = self._weights_file(savedir, seed, TESTS)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        print("\nSaving: ", path)
        torch.save(self.state_dict(), path)

    def load(self, device, savedir=None, seed=None, rel_path=TESTS):
        """Reads the reference's `<name>_weights[_seed].pt` state dict (model_nn.py:158-168)."""
        self.device = device
        path = self._weights_file(savedir, seed, rel_path)
        print("\nLoading: ", path)
        sd = torch.load(path, map_location="cpu")
        self.load_state_dict(sd)
        print("\n", list(sd.keys()), "\n")

    # ---- forward / gradient on the engine ---------------------------------------------------------
    def _members(self, n_samples):
        return 1

    def forward(self, inputs, *args, **kwargs):
        """Logits [B, C] on the device (model_nn.py:126-141)."""
        if self._rows_host is None:
            raise RuntimeError("no weights installed: call load() / load_state_dict() first")
        n = self._members(kwargs.get("n_samples", args[0] if args and isinstance(args[0], int) else None))
        return self.engine().forward_logits_sum(inputs, 0, n) / float(n)

    __call__ = forward

    def input_grad(self, image, label, n_samples=None):
        """d/dx CE(forward(x), y), summed over the batch: what the reference's attacks obtain by autograd
        (adversarialAttacks.py:73-79).  [B, D] on the device."""
        from ._lib import HEAD_LOGITS_UPSTREAM
        eng = self.engine()
        n = self._members(n_samples)
        x = torch.as_tensor(image).detach().to(device=eng.device, dtype=torch.float32).contiguous()
        y = torch.as_tensor(label).to(eng.device).reshape(-1).long()
        logits = eng.forward_logits_sum(x, 0, n) / float(n)
        g = torch.softmax(logits, -1)
        g[torch.arange(len(y), device=eng.device), y] -= 1.0
        g /= float(n)                                                  # d(mean logits) / d(member logits)
        return eng.input_grad_sum(HEAD_LOGITS_UPSTREAM, x, y.to(torch.int32), 0, n, pbar=g.contiguous())

    def evaluate(self, test_loader, device, *args, **kwargs):
        """Accuracy in percent (model_nn.py:220-238); integer count inside."""
        eng = self.engine()
        n_samples = kwargs.get("n_samples", args[0] if args else None)
        counter = torch.zeros((1,), dtype=torch.int64, device=eng.device)
        total = 0
        for x_batch, y_batch in test_loader:
            out = self.forward(x_batch, n_samples=n_samples)
            labels = torch.as_tensor(y_batch).to(eng.device).argmax(-1).to(torch.int32).contiguous()
            eng.count_correct(out.contiguous(), labels, counter)
            total += len(labels)
        accuracy = 100 * float(counter.item()) / total
        print("\nAccuracy: %.2f%%" % (accuracy))
        return accuracy
