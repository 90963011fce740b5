import numpy as np


def to_categorical(y, num_classes=None):
    y = np.asarray(y, dtype="int64").ravel()
    n = int(num_classes if num_classes is not None else y.max() + 1)
    out = np.zeros((len(y), n), dtype="float32")
    out[np.arange(len(y)), y] = 1.0
    return out
