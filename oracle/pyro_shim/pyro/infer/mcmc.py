from . import _Unavailable

MCMC = HMC = NUTS = _Unavailable
