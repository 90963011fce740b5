"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/rbnn.h declares, the ctypes table matches the header, argument validation that needs no
GPU behaves, and the product never touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "rbnn.h")).read()
    return re.findall(r"RBNN_API\s+[\w\s\*]+?\b(rbnn_\w+)\s*\(", text)


def test_library_exports_every_header_symbol():
    from robustbnns_b200 import _lib
    names = _header_symbols()
    assert len(names) >= 20
    lib = _lib.lib()
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(names) == sorted(_lib.SIGNATURES.keys())
    assert lib.rbnn_abi_version() == 1


def test_exported_symbols_are_only_the_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "robustbnns_b200", "librbnn.so")],
                         capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(_header_symbols())


def test_argument_validation_without_gpu():
    from robustbnns_b200 import _lib
    lib = _lib.lib()
    h = C.c_void_p()
    assert lib.rbnn_net_create(C.byref(h), 0, 1, 28, 28, 24, 10, 0) != 0          # model_nn.py:39-40
    assert b"power of 2" in lib.rbnn_last_error()
    assert lib.rbnn_net_create(C.byref(h), 3, 1, 28, 28, 32, 10, 0) != 0          # conv2: model_nn.py:123-124
    assert b"not implemented" in lib.rbnn_last_error()
    assert lib.rbnn_net_create(C.byref(h), 2, 1, 2, 1, 32, 2, 0) != 0             # conv on non-image input
    assert lib.rbnn_net_param_count(None) == -1


def test_python_mirror_validation():
    from robustbnns_b200.model_bnn import BNN, saved_BNNs
    from robustbnns_b200.model_nn import NN, param_layout
    with pytest.raises(ValueError):
        NN("mnist", (1, 28, 28), 10, 24, "leaky", "fc", 0.01, 1)
    with pytest.raises(AssertionError):
        NN("mnist", (1, 28, 28), 10, 32, "gelu", "fc", 0.01, 1)
    with pytest.raises(NotImplementedError):
        NN("half_moons", (1, 2, 1), 2, 32, "leaky", "conv", 0.01, 1)
    with pytest.raises(NotImplementedError):
        NN("mnist", (1, 28, 28), 10, 32, "leaky", "conv2", 0.01, 1)
    assert sum(int(__import__("numpy").prod(s)) for _, s in param_layout("fc", (1, 28, 28), 512, 10)) == 407050
    assert sum(int(__import__("numpy").prod(s)) for _, s in param_layout("fc2", (1, 28, 28), 512, 10)) == 669706
    assert sum(int(__import__("numpy").prod(s)) for _, s in param_layout("conv", (1, 28, 28), 512, 10)) == 661834
    dataset, model = saved_BNNs["model_0"]
    b = BNN(dataset, *list(model.values()), (1, 28, 28), 10)
    assert b.name == "mnist_bnn_svi_hid=512_act=leaky_arch=conv_ep=5_lr=0.01"
    dataset, model = saved_BNNs["model_9"]
    b = BNN(dataset, *list(model.values()), (1, 28, 28), 10)
    assert b.name == "fashion_mnist_bnn_hmc_hid=512_act=leaky_arch=fc_samp=100_warm=100_stepsize=0.005_numsteps=10"


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "robustbnns_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "pyro_shim" not in text, f


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from robustbnns_b200.engine import Net
    with pytest.raises(RuntimeError):
        Net("fc", (1, 28, 28), 32, 10)
