"""Import placeholder for keras (dataset download only in the reference, utils.py:10-11)."""
from . import datasets, utils  # noqa: F401
