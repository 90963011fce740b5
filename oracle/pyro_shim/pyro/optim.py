"""Import placeholder (training is out of scope for the hot path)."""


def Adam(*args, **kwargs):
    raise NotImplementedError("pyro shim: training is out of scope")
