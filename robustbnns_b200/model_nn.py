"""Host-side mirror of the reference's `model_nn.NN` for the hot path (model_nn.py:34-141).

Only what the Bayesian hot path needs from `NN` lives here: the constructor's
validation, the architecture definition (as the ordered list of state_dict
tensors a posterior sample consists of) and the name.  Deterministic-network
training / saving (`NN.train/save/load`, model_nn.py:143-239) is out of scope
 and raises NotImplementedError.
"""
import math

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
        if activation != "leaky":
            # every saved model uses leaky (model_bnn.py:36-66); the CUDA path implements only it
            raise NotImplementedError("the B200 path implements activation='leaky' only")
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

    def train(self, *args, **kwargs):
        raise NotImplementedError("deterministic-network training is outside the accelerated hot path")

    save = load = evaluate = train
