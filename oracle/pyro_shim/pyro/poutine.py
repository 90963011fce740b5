"""poutine.trace restated for the shim (see pyro/__init__.py header)."""
from collections import OrderedDict


class _Trace(object):
    def __init__(self):
        self.nodes = OrderedDict()


class trace(object):
    def __init__(self, fn):
        self.fn = fn

    def get_trace(self, *args, **kwargs):
        import pyro
        tr = _Trace()
        pyro._TRACE_STACK.append(tr)
        try:
            ret = self.fn(*args, **kwargs)
        finally:
            pyro._TRACE_STACK.pop()
        tr.nodes["_RETURN"] = {"type": "return", "name": "_RETURN", "value": ret}
        return tr
