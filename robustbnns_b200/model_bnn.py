"""Drop-in `BNN` for the hot path of the reference's `model_bnn.BNN` (model_bnn.py:69-391).

Same constructor, `name`, `forward(inputs, n_samples, avg_posterior, seeds)`,
`evaluate`, `load` and `saved_BNNs`; the inside is a posterior-sample bank in HBM
and hand-written sm_100a kernels behind the C ABI (include/rbnn.h).  Training
(`model`, `_train_svi`, `_train_hmc`, `train`) is outside the accelerated path
and raises NotImplementedError.

Posterior samples
-----------------
* SVI, `seeds=[...]`: sample `seed` is the Philox stream of global index `seed`
  (the reference: a guide draw under `pyro.set_rng_seed(seed)`, model_bnn.py:222-226)
  -- the same weights for every input, so `seeds=range(k)` is a prefix of
  `seeds=range(n)`, k<n.  RNG streams differ from torch's, parity is statistical;
  pin an explicit bank with `set_posterior_samples` for exact parity.
* SVI, no seeds: fresh draws on every forward call (model_bnn.py:230-232); the
  draws of one call are shared by the inputs of that call.  `reseed(s)` plays the
  part of `pyro.set_rng_seed(s)`.  `attack()` calls the reference's per-image
  attack once per image, i.e. with fresh weights per image: set
  `bnn.fresh_draws = "per_image"` for that (one device pass per image).
* HMC / explicit bank: row i is `posterior_predictive[i]` (model_bnn.py:184-190,
  :248-255); no seeds means the first `n_samples` rows.
* With torch.distributed initialised, position j of a call's sample list lives
  on rank j % world_size and the [B, C] sums are all-reduced (dist.py).
"""
import os

import torch

from . import dist as rdist
from ._lib import HEAD_UPSTREAM
from .model_nn import NN
from .savedir import TESTS

DEBUG = False

saved_BNNs = {"model_0": ["mnist", {"hidden_size": 512, "activation": "leaky",
                          "architecture": "conv", "inference": "svi", "epochs": 5,
                          "lr": 0.01, "n_samples": None, "warmup": None}],
              "model_1": ["mnist", {"hidden_size": 512, "activation": "leaky",
                          "architecture": "fc2", "inference": "hmc", "epochs": None,
                          "lr": None, "n_samples": 100, "warmup": 50}],
              "model_2": ["fashion_mnist", {"hidden_size": 1024, "activation": "leaky",
                          "architecture": "conv", "inference": "svi", "epochs": 10,
                          "lr": 0.001, "n_samples": None, "warmup": None}],
              "model_3": ["fashion_mnist", {"hidden_size": 1024, "activation": "leaky",
                          "architecture": "fc2", "inference": "hmc", "epochs": None,
                          "lr": None, "n_samples": 100, "warmup": 50}],
              "model_4": ["fashion_mnist", {"hidden_size": 1024, "activation": "leaky",
                          "architecture": "conv", "inference": "svi", "epochs": 5,
                          "lr": 0.01, "n_samples": None, "warmup": None}],
              "model_5": ["mnist", {"hidden_size": 512, "activation": "leaky",
                          "architecture": "fc2", "inference": "svi", "epochs": 10,
                          "lr": 0.01, "n_samples": None, "warmup": None}],
              "model_6": ["mnist", {"hidden_size": 256, "activation": "leaky",
                          "architecture": "conv", "inference": "svi", "epochs": 10,
                          "lr": 0.05, "n_samples": None, "warmup": None}],
              "model_7": ["mnist", {"hidden_size": 1024, "activation": "leaky",
                          "architecture": "fc2", "inference": "svi", "epochs": 5,
                          "lr": 0.02, "n_samples": None, "warmup": None}],
              "model_8": ["mnist", {"hidden_size": 1024, "activation": "leaky",
                          "architecture": "conv", "inference": "svi", "epochs": 10,
                          "lr": 0.02, "n_samples": None, "warmup": None}],
              "model_9": ["fashion_mnist", {"hidden_size": 512, "activation": "leaky",
                          "architecture": "fc", "inference": "hmc", "epochs": None,
                          "lr": None, "n_samples": 100, "warmup": 100}],
              }

FRESH_BASE = 1 << 31          # global sample indices of unseeded ("fresh") draws start here
_GOLDEN = 0x9E3779B97F4A7C15


class _ForwardFn(torch.autograd.Function):
    """Makes BNN.forward differentiable w.r.t. its inputs (what the reference's
    attacks rely on, adversarialAttacks.py:73-79) with the CUDA input-gradient pass."""

    @staticmethod
    def forward(ctx, inputs, bnn, rows, n_total, generation):
        ctx.bnn, ctx.rows, ctx.n_total, ctx.generation = bnn, rows, n_total, generation
        ctx.save_for_backward(inputs)
        eng = bnn.engine()
        out = eng.forward_probs_sum(inputs.detach(), rows[0], rows[1], keep=True)     # logits + masks stay for backward()
        ctx.keep_serial = eng.keep_serial if eng.keep_valid else None
        rdist.allreduce_sum_(out)
        return out / float(n_total)

    @staticmethod
    def backward(ctx, grad_out):
        (inputs,) = ctx.saved_tensors
        bnn = ctx.bnn
        if ctx.generation != bnn._scratch_generation and ctx.rows[0] >= bnn._pin_cap:
            raise RuntimeError("BNN.forward's fresh posterior samples were overwritten by a later forward "
                               "call before backward(); call backward() first or pin a sample bank")
        eng = bnn.engine()
        labels = torch.zeros((inputs.shape[0],), dtype=torch.int32, device=eng.device)
        if ctx.keep_serial is not None and eng.keep_valid and eng.keep_serial == ctx.keep_serial:
            g = eng.input_grad_sum_kept(HEAD_UPSTREAM, labels, pbar=grad_out.contiguous())   # no second forward pass
        else:
            g = eng.input_grad_sum(HEAD_UPSTREAM, inputs.detach(), labels, ctx.rows[0], ctx.rows[1],
                                   pbar=grad_out.contiguous())
        rdist.allreduce_sum_(g)
        g = (g / float(ctx.n_total)).reshape(inputs.shape)
        return g, None, None, None, None


class BNN(object):

    def __init__(self, dataset_name, hidden_size, activation, architecture, inference,
                 epochs, lr, n_samples, warmup, input_shape, output_size,
                 step_size=0.005, num_steps=10, engine=None):
        self.dataset_name = dataset_name
        self.inference = inference
        self.architecture = architecture
        self.epochs = epochs
        self.lr = lr
        self.n_samples = n_samples
        self.warmup = warmup
        self.step_size = step_size
        self.num_steps = num_steps
        self.basenet = NN(dataset_name=dataset_name, input_shape=input_shape, output_size=output_size,
                          hidden_size=hidden_size, activation=activation, architecture=architecture,
                          epochs=epochs, lr=lr)
        self.name = self.get_name()
        self.input_shape = self.basenet.input_shape
        self.output_size = output_size
        self.rng_seed = 0
        self._engine = engine
        self._precision = "auto"              # engine choice for an engine this object creates (set_precision)
        self.attack_sharding = "inputs"       # multi-GPU `attack`: shard the inputs (no collectives) or the "samples"
        self.fresh_draws = "per_pass"         # unseeded SVI attacks: "per_image" = own fresh weights per image, as upstream
        self._loc = self._rho = None          # SVI guide parameters, flattened [P]
        self._bank_host = None                # explicit / HMC bank [S, P] (CPU tensor)
        self._graph_offset = None             # int64[1] device tensor while a CUDA graph of a PGD iteration is captured
        self._posterior_generation = 0        # bumped when another posterior is installed (cached CUDA graphs die with it)
        self._pgd_graphs = {}
        self._reset_rows()
        self.reseed(0)

    # ---- naming (model_bnn.py:90-103) -----------------------------------------------------
    def get_name(self, n_inputs=None):
        name = str(self.dataset_name) + "_bnn_" + str(self.inference) + "_hid=" + \
            str(self.basenet.hidden_size) + "_act=" + str(self.basenet.activation) + \
            "_arch=" + str(self.basenet.architecture)
        if n_inputs:
            name = name + "_inp=" + str(n_inputs)
        if self.inference == "svi":
            return name + "_ep=" + str(self.epochs) + "_lr=" + str(self.lr)
        elif self.inference == "hmc":
            return name + "_samp=" + str(self.n_samples) + "_warm=" + str(self.warmup) + \
                "_stepsize=" + str(self.step_size) + "_numsteps=" + str(self.num_steps)

    # ---- engine / bank bookkeeping -----------------------------------------------------------
    def engine(self):
        if self._engine is None:
            from .engine import Net
            self._engine = Net(self.basenet.architecture, self.input_shape, self.basenet.hidden_size,
                               self.output_size)
            if self.basenet.activation != "leaky":          # relu / sigm / tanh: FP32 engine, arch fc / fc2
                self._engine.set_activation(self.basenet.activation)
            if self._precision == "auto":
                self._engine.set_best_precision()
            else:
                self._engine.set_precision(self._precision)
        return self._engine

    def _reset_rows(self):
        self._pin_rows = 0            # valid local pinned rows
        self._pin_cap = 0             # local rows reserved for the pinned region; scratch starts here
        self._scratch_generation = 0

    def _new_posterior(self):
        """Other weights are about to live in the engine's bank: what the engine derived from the old ones (the frozen
        F16X3 operand scale, kernel-ready copies, the kept forward) and every captured CUDA graph are dropped."""
        self._posterior_generation += 1
        self._pgd_graphs = {}
        if self._engine is not None and hasattr(self._engine, "invalidate"):
            self._engine.invalidate()

    def set_precision(self, name):
        """'auto' (default: the fastest parity-grade engine the network has -- 'f16x3' for arch fc / conv, 'tf32x3'
        for fc2, 'fp32' otherwise), 'fp32' (CUDA-core FFMA), 'f16x3' / 'tf32x3' (tcgen05, fp32-class accuracy) or
        'bf16' (tcgen05 single pass, throughput mode, not parity grade); see DESIGN.md section 4.1."""
        self._precision = name
        self._pgd_graphs = {}
        if name == "auto":
            self.engine().set_best_precision()
        else:
            self.engine().set_precision(name)

    def reseed(self, seed):
        """Stand-in for pyro.set_rng_seed(seed) before unseeded forwards (adversarialAttacks.py:161)."""
        self._fresh_key = (self.rng_seed + _GOLDEN * (int(seed) + 1)) % (1 << 64)
        self._fresh_counter = 0

    def _flatten_params(self, params, suffix):
        if torch.is_tensor(params):
            flat = params.detach().reshape(-1).float()
        else:
            flat = torch.cat([torch.as_tensor(params[k + suffix]).detach().reshape(-1).float()
                              for k in self.basenet.state_dict_keys()])
        if flat.numel() != self.basenet.n_params:
            raise ValueError("expected %d parameters, got %d" % (self.basenet.n_params, flat.numel()))
        return flat

    def set_guide(self, loc, scale):
        """SVI posterior N(loc, softplus(scale)^2) (model_bnn.py:125-127): flat [P] tensors in
        state_dict order, or dicts keyed "<key>_loc" / "<key>_scale" like the Pyro param store."""
        self._loc = self._flatten_params(loc, "_loc")
        self._rho = self._flatten_params(scale, "_scale")
        if self._engine is not None or torch.cuda.is_available():
            dev = self.engine().device
            self._loc, self._rho = self._loc.to(dev), self._rho.to(dev)
        self._new_posterior()
        self._reset_rows()

    def set_posterior_samples(self, bank):
        """Pin an explicit bank: a [S, P] tensor, or a list of S state dicts (the HMC
        `posterior_predictive` networks, model_bnn.py:184-190).  Row i answers seed i."""
        if not torch.is_tensor(bank):
            keys = self.basenet.state_dict_keys()
            bank = torch.stack([torch.cat([torch.as_tensor(sd[k]).detach().reshape(-1).float().cpu() for k in keys])
                                for sd in bank])
        bank = bank.detach().float().cpu().contiguous()
        if bank.dim() != 2 or bank.shape[1] != self.basenet.n_params:
            raise ValueError("bank must be [S, %d]" % self.basenet.n_params)
        same = self._bank_host is not None and self._bank_host.shape == bank.shape and (
            self._bank_host.data_ptr() == bank.data_ptr() or torch.equal(self._bank_host, bank))
        self._bank_host = bank
        if not same:                          # (_replace_rows re-installs the SAME bank under another sharding)
            self._new_posterior()
        self._reset_rows()
        rank, world = rdist.world()
        mine = bank[rank::world]
        eng = self.engine()
        self._pin_cap = max(1, mine.shape[0])
        eng.reserve(self._pin_cap + 1)
        if mine.shape[0]:
            eng.upload(mine, 0)
        self._pin_rows = mine.shape[0]

    # ---- persistence in the reference's formats (model_bnn.py:138-196) ---------------------------
    def save(self, rel_path=TESTS, filename=None):
        if filename is None:
            filename = self.name + "_weights"
        path = rel_path + self.name + "/"
        os.makedirs(os.path.dirname(path), exist_ok=True)
        print(f"\nSaving {path}{filename}")
        if self.inference == "svi":
            params, off = {}, 0
            for key, shp in self.basenet.layout:
                n = int(torch.Size(shp).numel())
                params[key + "_loc"] = self._loc[off:off + n].reshape(shp).cpu().clone()
                params[key + "_scale"] = self._rho[off:off + n].reshape(shp).cpu().clone()
                off += n
            # pyro.get_param_store().get_state(): unconstrained values + the constraint OBJECT of every parameter
            from torch.distributions import constraints
            torch.save({"params": params, "constraints": {k: constraints.real for k in params}}, path + filename + ".pt")
        elif self.inference == "hmc":
            for idx in range(self._bank_host.shape[0]):
                sd, off = {}, 0
                for key, shp in self.basenet.layout:
                    n = int(torch.Size(shp).numel())
                    sd[key] = self._bank_host[idx, off:off + n].reshape(shp).clone()
                    off += n
                torch.save(sd, path + filename + "_" + str(idx) + ".pt")

    def load(self, device, rel_path=TESTS, filename=None):
        if filename is None:
            filename = self.name + "_weights"
        path = rel_path + self.name + "/"
        self.device = device
        if self.inference == "svi":
            state = torch.load(path + filename + ".pt", map_location="cpu", weights_only=False)
            self.set_guide(state["params"], state["params"])
            print("\nLoading ", path + filename + ".pt\n")
        elif self.inference == "hmc":
            nets = []
            for model_idx in range(self.n_samples):
                f = path + filename + "_" + str(model_idx) + ".pt"
                if not os.path.exists(f):
                    break
                nets.append(torch.load(f, map_location="cpu", weights_only=False))
            if len(nets) != self.n_samples:
                raise AttributeError("wrong number of posterior models")
            self.set_posterior_samples(nets)

    def to(self, device):
        return self

    def zero_grad(self):
        """No parameter gradients are ever produced (input gradients only)."""
        return None

    def train(self, *args, **kwargs):
        raise NotImplementedError("posterior inference (SVI/HMC training) is outside the accelerated hot path")

    model = guide = _train_svi = _train_hmc = train

    def _replace_rows(self):
        """Redo the bank placement after the sharding mode changed (rdist.replicated() entered / left)."""
        if self._bank_host is not None:
            self.set_posterior_samples(self._bank_host)
        else:
            self._reset_rows()

    # ---- sample placement --------------------------------------------------------------------
    def _grow_pinned(self, need_local):
        if need_local > self._pin_cap:
            cap = 1
            while cap < need_local:
                cap *= 2
            self._pin_cap = cap
            self._scratch_generation += 1

    def _rows(self, n_samples, seeds):
        """Place this rank's share of the call's samples in consecutive bank rows.
        Returns ((s0, s1), scratch_generation)."""
        rank, world = rdist.world()
        eng = self.engine()
        n = int(n_samples)
        nloc = rdist.local_count(n, rank, world)
        ids = None if seeds is None else [int(s) for s in seeds]
        explicit = self._bank_host is not None
        if explicit:
            S = self._bank_host.shape[0]
            if ids is None:
                ids = list(range(n))
            if any(i >= S or i < -S for i in ids):
                raise IndexError("list index out of range")          # posterior_predictive[seed], model_bnn.py:252
            if ids == list(range(n)):
                return (0, nloc), self._scratch_generation
            rows = self._bank_host[[ids[j] for j in rdist.local_positions(n, rank, world)]]
            self._scratch_generation += 1
            eng.reserve(self._pin_cap + max(nloc, 1))
            if nloc:
                eng.upload(rows, self._pin_cap)
            return (self._pin_cap, self._pin_cap + nloc), self._scratch_generation
        if self._loc is None:
            raise RuntimeError("BNN has no posterior: call load(), set_guide() or set_posterior_samples() first")
        if ids is not None and ids == list(range(n)):
            if nloc > self._pin_rows:
                self._grow_pinned(nloc)
                eng.reserve(self._pin_cap + 1)
                eng.sample_diag(self._loc, self._rho, self.rng_seed, rank + self._pin_rows * world,
                                self._pin_rows, nloc - self._pin_rows, stride=world)
                self._pin_rows = nloc
            return (0, nloc), self._scratch_generation
        self._scratch_generation += 1
        base = self._pin_cap
        eng.reserve(base + max(nloc, 1))
        if ids is not None:
            for i, j in enumerate(rdist.local_positions(n, rank, world)):
                eng.sample_diag(self._loc, self._rho, self.rng_seed, ids[j], base + i, 1)
        else:
            first = FRESH_BASE + self._fresh_counter
            self._fresh_counter += n
            if nloc:
                if self._graph_offset is not None:      # CUDA-graph capture: the draw index advances on the device
                    eng.sample_diag(self._loc, self._rho, self._fresh_key, first + rank, base, nloc, stride=world,
                                    index_offset=self._graph_offset)
                else:
                    eng.sample_diag(self._loc, self._rho, self._fresh_key, first + rank, base, nloc, stride=world)
        return (base, base + nloc), self._scratch_generation

    def _probs_mean(self, inputs, rows, n_total):
        out = self.engine().forward_probs_sum(inputs, rows[0], rows[1])
        rdist.allreduce_sum_(out)
        return out / float(n_total)

    # ---- a3: forward (model_bnn.py:198-258) -----------------------------------------------------
    def forward(self, inputs, n_samples=10, avg_posterior=False, seeds=None):
        if seeds:
            if len(seeds) != n_samples:
                raise ValueError("Number of seeds should match number of samples.")
        eng = self.engine()
        inputs = torch.as_tensor(inputs)
        if self.inference == "svi" and avg_posterior is True:
            if self._loc is None:
                raise RuntimeError("avg_posterior needs the guide parameters (set_guide/load)")
            row = self._pin_cap
            self._scratch_generation += 1
            eng.reserve(row + 1)
            eng.upload(self._loc.reshape(1, -1), row)
            return eng.forward_logits(inputs, row)          # LOGITS (model_bnn.py:206-216)
        x = inputs.to(device=eng.device, dtype=torch.float32)
        rows, gen = self._rows(n_samples, seeds if seeds else None)
        if x.requires_grad:
            return _ForwardFn.apply(x, self, rows, int(n_samples), gen)
        return self._probs_mean(x, rows, int(n_samples))

    __call__ = forward

    # ---- a5: evaluate (model_bnn.py:367-391) ------------------------------------------------------
    def evaluate(self, test_loader, device, n_samples=10, seeds_list=None):
        self.device = device
        self.reseed(0)
        bnn_seeds = list(range(n_samples)) if seeds_list is None else seeds_list
        eng = self.engine()
        counter = torch.zeros((1,), dtype=torch.int64, device=eng.device)
        total = 0
        with torch.no_grad():
            for x_batch, y_batch in test_loader:
                outputs = self.forward(x_batch, n_samples=n_samples, seeds=bnn_seeds)
                labels = torch.as_tensor(y_batch).to(eng.device).argmax(-1).to(torch.int32).contiguous()
                eng.count_correct(outputs.contiguous(), labels, counter)
                total += len(labels)
        n_data = len(test_loader.dataset) if hasattr(test_loader, "dataset") else total
        accuracy = 100 * float(counter.item()) / n_data
        print("Accuracy: %.2f%%" % (accuracy))
        return accuracy
