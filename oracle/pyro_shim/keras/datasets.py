class _NoData(object):
    @staticmethod
    def load_data():
        raise RuntimeError("keras shim: no network, no datasets")


mnist = fashion_mnist = _NoData
