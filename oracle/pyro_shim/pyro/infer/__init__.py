"""Import placeholders (training is out of scope for the hot path)."""


class _Unavailable(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("pyro shim: training/inference of the posterior is out of scope")


SVI = Trace_ELBO = TraceMeanField_ELBO = Predictive = _Unavailable
