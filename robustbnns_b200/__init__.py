"""robustbnns_b200 -- B200-native drop-in for the hot path of ginevracoal/robustBNNs.

Modules mirror the reference's file names for that path:
    model_nn.NN, model_bnn.BNN / saved_BNNs, lossGradients.loss_gradient(s),
    adversarialAttacks.{fgsm_attack, pgd_attack, attack, attack_evaluation,
    softmax_difference, softmax_robustness}
All arithmetic runs in librbnn.so (hand-written sm_100a CUDA behind the C ABI of
include/rbnn.h).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .model_bnn import BNN, saved_BNNs  # noqa: F401
from .model_nn import NN, saved_NNs  # noqa: F401

__all__ = ["BNN", "NN", "saved_BNNs", "saved_NNs"]
