"""Minimal stand-in for pyro-ppl 1.3.0 -- TEST INFRASTRUCTURE ONLY.

Pyro is not installed in the build container (no network), so the reference
modules under /root/reference cannot be imported as they are.  This shim
re-states the *published* semantics of the handful of Pyro 1.3.0 entry points
the reference's hot path touches (call sites: model_bnn.py:114,125-130,178-181,
208,224-225,231; adversarialAttacks.py:161) so that the reference's own,
unmodified Python can be executed here to produce golden vectors
(tests/golden/make_golden.py).  It is never imported by the product package.

Semantics restated (Pyro 1.3.0, pyro/primitives.py, pyro/poutine/lift_messenger.py,
pyro/util.py):
  * pyro.param(name, init)      -- global param store; first call stores `init`
                                   as an unconstrained leaf requiring grad, later
                                   calls return the stored tensor (the eager
                                   `init` argument is evaluated by the caller
                                   either way and is discarded).
  * pyro.random_module(n, m, d) -- returns a callable; calling it deep-copies the
                                   nn.Module `m` and replaces every named
                                   parameter `p` by a draw from `d[p]` recorded at
                                   site "n$$$p" (rsample for reparameterised
                                   distributions), in named_parameters() order.
  * poutine.trace(fn).get_trace -- runs fn, records param / sample sites and the
                                   return value at node "_RETURN".
  * pyro.set_rng_seed(s)        -- torch.manual_seed(s); random.seed(s);
                                   numpy.random.seed(s).
Training-side entry points (SVI, HMC, optim) exist only as import placeholders.
"""
import copy
import random as _random

import numpy as _np
import torch as _torch

__version__ = "1.3.0+shim"

_PARAM_STORE = {}
_TRACE_STACK = []


class _ParamStore(object):
    def get_all_param_names(self):
        return list(_PARAM_STORE.keys())

    def items(self):
        return list(_PARAM_STORE.items())

    def keys(self):
        return list(_PARAM_STORE.keys())

    def __getitem__(self, name):
        return _PARAM_STORE[name]

    def __contains__(self, name):
        return name in _PARAM_STORE

    def replace_param(self, name, new_param, old_param):
        _PARAM_STORE[name] = new_param

    def save(self, filename):
        _torch.save({"params": {k: v.detach().clone() for k, v in _PARAM_STORE.items()},
                     "constraints": {k: _torch.distributions.constraints.real for k in _PARAM_STORE}}, filename)

    def load(self, filename, map_location=None):
        state = _torch.load(filename, map_location=map_location)
        for k, v in state["params"].items():
            _PARAM_STORE[k] = v.detach().clone().requires_grad_(True)

    def clear(self):
        _PARAM_STORE.clear()


def get_param_store():
    return _ParamStore()


def clear_param_store():
    _PARAM_STORE.clear()


def set_rng_seed(seed):
    _torch.manual_seed(seed)
    _random.seed(seed)
    _np.random.seed(seed)


def _record(name, node):
    if _TRACE_STACK:
        _TRACE_STACK[-1].nodes[name] = node


def param(name, init_tensor=None, constraint=None, event_dim=None):
    if name not in _PARAM_STORE:
        if init_tensor is None:
            raise KeyError(name)
        _PARAM_STORE[name] = init_tensor.detach().clone().requires_grad_(True)
    value = _PARAM_STORE[name]
    _record(name, {"type": "param", "name": name, "value": value})
    return value


def sample(name, fn, obs=None, **kwargs):
    if obs is not None:
        value = obs
    elif getattr(fn, "has_rsample", False):
        value = fn.rsample()
    else:
        value = fn.sample()
    _record(name, {"type": "sample", "name": name, "fn": fn, "value": value,
                   "is_observed": obs is not None})
    return value


class plate(object):
    def __init__(self, name, size=None, subsample_size=None, dim=None, **kwargs):
        self.name, self.size = name, size

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _LiftedModule(object):
    """Deep copy of an nn.Module evaluated with sampled parameter tensors."""

    def __init__(self, module, values):
        self._module = module
        self._values = values

    def __call__(self, *args, **kwargs):
        return _torch.func.functional_call(self._module, self._values, args, kwargs)

    forward = __call__


def random_module(name, nn_module, prior, *args, **kwargs):
    def _fn():
        nn_copy = copy.deepcopy(nn_module)
        values = {}
        for pname, p in nn_copy.named_parameters():
            if isinstance(prior, dict):
                dist = prior[pname]
            else:
                dist = prior
            values[pname] = sample("{}$$${}".format(name, pname), dist)
        return _LiftedModule(nn_copy, values)
    return _fn


from . import poutine  # noqa: E402
from . import distributions  # noqa: E402
from . import infer  # noqa: E402
from . import optim  # noqa: E402
from . import nn  # noqa: E402
