"""ctypes binding of librbnn.so (C ABI declared in include/rbnn.h).

There is no CPU fallback: if the shared library is missing, or a call fails, an
exception is raised.  Build it with `python __graft_entry__.py` (or
`make -C robustbnns_b200/csrc`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librbnn.so")

ARCH = {"fc": 0, "fc2": 1, "conv": 2}
PREC = {"fp32": 0, "tf32x3": 1, "bf16": 2, "f16x3": 3}
ACT = {"leaky": 0, "relu": 1, "sigm": 2, "tanh": 3}      # RBNN_ACT_* (model_nn.py:66-75)
HEAD_MEAN_OF_GRADS, HEAD_GRAD_OF_MEAN, HEAD_LOGITS_CE, HEAD_UPSTREAM, HEAD_LOGITS_UPSTREAM = 0, 1, 2, 3, 4

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_i64 = C.c_int64
_u64 = C.c_uint64

# name -> (restype, argtypes); must list every RBNN_API symbol of include/rbnn.h
SIGNATURES = {
    "rbnn_abi_version": (_i, []),
    "rbnn_last_error": (C.c_char_p, []),
    "rbnn_net_create": (_i, [C.POINTER(_p), _i, _i, _i, _i, _i, _i, _i]),
    "rbnn_net_destroy": (_i, [_p]),
    "rbnn_net_param_count": (_i64, [_p]),
    "rbnn_net_set_precision": (_i, [_p, _i]),
    "rbnn_net_get_precision": (_i, [_p]),
    "rbnn_net_launch_count": (_i64, [_p]),
    "rbnn_net_input_grid": (_i, [_p]),
    "rbnn_net_set_activation": (_i, [_p, _i]),
    "rbnn_net_timing_enable": (_i, [_p, _i]),
    "rbnn_net_timing_read": (_i, [_p, _i, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "rbnn_bank_reserve": (_i, [_p, _i]),
    "rbnn_bank_capacity": (_i, [_p]),
    "rbnn_bank_upload": (_i, [_p, _p, _i, _i, _i, _p]),
    "rbnn_bank_sample_diag": (_i, [_p, _p, _p, _u64, _i64, _i64, _i, _i, _p]),
    "rbnn_bank_sample_diag_at": (_i, [_p, _p, _p, _u64, _i64, _i64, _i, _i, _p, _p]),
    "rbnn_bank_invalidate": (_i, [_p]),
    "rbnn_net_alloc_epoch": (_i64, [_p]),
    "rbnn_bank_download": (_i, [_p, _p, _i, _i]),
    "rbnn_forward_probs_sum": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "rbnn_forward_probs_sum_keep": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "rbnn_keep_valid": (_i, [_p]),
    "rbnn_input_grad_sum_kept": (_i, [_p, _i, _p, _p, _p, _p]),
    "rbnn_forward_logits": (_i, [_p, _p, _i, _i, _p, _p]),
    "rbnn_forward_logits_sum": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "rbnn_input_grad_sum": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _p, _p]),
    "rbnn_fgsm_step": (_i, [_p, _p, _f, _p, _i64, _p]),
    "rbnn_pgd_step": (_i, [_p, _p, _p, _p, _f, _p, _i, _i, _p]),
    "rbnn_pgd_alpha": (_i, [_p, _p, _i, _i, _p]),
    "rbnn_softmax_robustness": (_i, [_p, _p, _i, _i, _p, _p, _p]),
    "rbnn_count_correct": (_i, [_p, _p, _i, _i, _p, _p]),
    "rbnn_loss_gradients_host": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
}

_lib = None


class RbnnError(RuntimeError):
    pass


def lib():
    """Load librbnn.so once; raise loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "robustbnns_b200: %s is missing -- the CUDA library is not built. Run "
                "`python __graft_entry__.py` (build()) or `make -C robustbnns_b200/csrc`. "
                "There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RbnnError(lib().rbnn_last_error().decode("utf-8", "replace"))
