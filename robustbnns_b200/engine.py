"""`Net`: the Python handle on one rbnn_net (architecture + posterior-sample bank on one GPU).

Thin, typed wrapper over the C ABI: it checks tensors (device, dtype, contiguity),
passes raw device pointers plus torch's current CUDA stream, and raises on any
non-zero status.  All arithmetic happens inside librbnn.so.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ARCH, PREC, check, lib


def _stream(device=None):
    """torch's current stream ON `device` (the handle's / the tensors' device, not necessarily the current one)."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Net(object):
    def __init__(self, arch, input_shape, hidden, n_classes, device=None):
        if arch not in ARCH:
            raise NotImplementedError()          # model_nn.py:123-124
        if not torch.cuda.is_available():
            raise RuntimeError("robustbnns_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("robustbnns_b200 runs on CUDA devices only, got %r" % (device,))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.arch, self.input_shape = arch, tuple(int(v) for v in input_shape)
        self.hidden, self.n_classes = int(hidden), int(n_classes)
        self.D = self.input_shape[0] * self.input_shape[1] * self.input_shape[2]
        h = C.c_void_p()
        check(lib().rbnn_net_create(C.byref(h), ARCH[arch], *self.input_shape, self.hidden, self.n_classes,
                                    self.device.index))
        self._h = h
        self.P = int(lib().rbnn_net_param_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().rbnn_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers --------------------------------------------------------------------------
    def _dev(self, t, dtype=torch.float32):
        t = torch.as_tensor(t)
        if t.dtype != dtype or t.device != self.device or not t.is_contiguous():
            t = t.to(device=self.device, dtype=dtype).contiguous()
        return t

    def set_precision(self, name):
        check(lib().rbnn_net_set_precision(self._h, PREC[name]))

    def set_activation(self, name):
        """Hidden-layer activation, model_nn.py:66-75: "leaky" (default; every engine), "relu" / "sigm" / "tanh" (arch fc /
        fc2 on the FP32 CUDA-core engine: the handle is switched to it)."""
        check(lib().rbnn_net_set_activation(self._h, _lib.ACT[name]))
        self.activation = name

    def set_best_precision(self):
        """The fastest parity-grade engine this network has: F16X3 (arch fc with a fused hidden size, arch conv), else
        TF32X3 (fc / fc2 with H >= 32; D % 8 != 0 -- half moons -- only fc2 from H = 64: its H x H layer runs on tcgen05,
        the D-wide first layer on the CUDA cores), else the FP32 CUDA-core engine.  Returns its name."""
        cands = ("f16x3", "tf32x3") if getattr(self, "activation", "leaky") == "leaky" else ()
        if self.D % 8 and self.hidden < 64:
            cands = ()        # half moons with narrow layers: launch-bound, the CUDA-core engine has fewer launches
        for cand in cands:
            try:
                self.set_precision(cand)
                return cand
            except _lib.RbnnError:
                continue
        self.set_precision("fp32")
        return "fp32"

    @property
    def precision(self):
        code = lib().rbnn_net_get_precision(self._h)
        return {v: k for k, v in PREC.items()}[code]

    @property
    def launch_count(self):
        return int(lib().rbnn_net_launch_count(self._h))

    @property
    def input_grid(self):
        """True when the last F16X3 forward found its inputs on the 8-bit pixel grid (uint8 / 255, what the reference's
        image loaders produce) and ran the two-pass forward; synchronises the device."""
        return int(lib().rbnn_net_input_grid(self._h)) == 1

    def timing_enable(self, on=True):
        check(lib().rbnn_net_timing_enable(self._h, int(bool(on))))

    def timing_read(self, cls):
        """(total kernel ms, launches) of class 1 (forward GEMM) / 2 (input-grad GEMM) since the last read."""
        ms, cnt = C.c_double(0.0), C.c_int64(0)
        check(lib().rbnn_net_timing_read(self._h, int(cls), C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    # ---- bank -------------------------------------------------------------------------------
    def reserve(self, capacity):
        check(lib().rbnn_bank_reserve(self._h, int(capacity)))

    @property
    def capacity(self):
        return int(lib().rbnn_bank_capacity(self._h))

    def upload(self, weights, s0=0):
        """rows [s0, s0+len) <- weights [count, P] (CPU or CUDA tensor)."""
        w = torch.as_tensor(weights, dtype=torch.float32)
        if w.dim() == 1:
            w = w.unsqueeze(0)
        if w.shape[1] != self.P:
            raise ValueError("bank rows must have %d parameters, got %d" % (self.P, w.shape[1]))
        w = w.contiguous()
        self.reserve(max(self.capacity, s0 + w.shape[0]))
        on_dev = w.is_cuda
        if on_dev:
            w = self._dev(w)
        check(lib().rbnn_bank_upload(self._h, C.c_void_p(w.data_ptr()), int(s0), int(w.shape[0]), int(on_dev),
                                     _stream(self.device)))
        if not on_dev:
            torch.cuda.current_stream(self.device).synchronize()   # pageable host source must outlive the copy

    def sample_diag(self, loc, rho, seed, sample_index0, s0, count, stride=1, index_offset=None):
        """rows [s0, s0+count) <- loc + softplus(rho)*eps(seed, global index)  (BNN.guide, model_bnn.py:121-130).
        index_offset: optional int64[1] DEVICE tensor added to the sample indices when the kernel runs (CUDA graphs)."""
        loc, rho = self._dev(loc).reshape(-1), self._dev(rho).reshape(-1)
        if loc.numel() != self.P or rho.numel() != self.P:
            raise ValueError("loc/rho must have %d elements" % self.P)
        self.reserve(max(self.capacity, s0 + count))
        off = 0
        if index_offset is not None:
            if index_offset.dtype != torch.int64 or index_offset.device != self.device:
                raise ValueError("index_offset must be an int64 tensor on %s" % self.device)
            off = index_offset.data_ptr()
        check(lib().rbnn_bank_sample_diag_at(self._h, C.c_void_p(loc.data_ptr()), C.c_void_p(rho.data_ptr()),
                                             C.c_uint64(int(seed) & (2 ** 64 - 1)), int(sample_index0), int(stride),
                                             int(s0), int(count), C.c_void_p(off), _stream(self.device)))

    def invalidate(self):
        """A new posterior is being installed: forget the derived operand scale / copies / kept forward."""
        check(lib().rbnn_bank_invalidate(self._h))

    @property
    def alloc_epoch(self):
        return int(lib().rbnn_net_alloc_epoch(self._h))

    def download(self, s0, count):
        out = torch.empty((count, self.P), dtype=torch.float32)
        check(lib().rbnn_bank_download(self._h, C.c_void_p(out.data_ptr()), int(s0), int(count)))
        return out

    # ---- compute ----------------------------------------------------------------------------
    def _x(self, x):
        x = self._dev(x)
        B = x.shape[0] if x.dim() > 0 else 0
        if x.numel() != B * self.D:
            raise ValueError("inputs of shape %s do not flatten to [B, %d]" % (tuple(x.shape), self.D))
        return x, B

    def forward_probs_sum(self, x, s0, s1, keep=False):
        """sum_s softmax(f_s(x)) over bank rows [s0, s1).  keep=True asks the engine to retain the per-sample logits
        and LeakyReLU masks for `input_grad_sum_kept` (two-phase attack gradient); `keep_valid` tells whether it did."""
        x, B = self._x(x)
        out = torch.empty((B, self.n_classes), dtype=torch.float32, device=self.device)
        fn = lib().rbnn_forward_probs_sum_keep if keep else lib().rbnn_forward_probs_sum
        if keep:
            self.keep_serial = getattr(self, "keep_serial", 0) + 1      # identifies WHICH forward is kept
        check(fn(self._h, C.c_void_p(x.data_ptr()), B, int(s0), int(s1), C.c_void_p(out.data_ptr()), _stream(self.device)))
        return out

    @property
    def keep_valid(self):
        return bool(lib().rbnn_keep_valid(self._h))

    def input_grad_sum_kept(self, head, labels, pbar=None):
        """Gradient pass of the last kept forward: [B, D] sum over its bank rows, no second forward GEMM."""
        labels = self._dev(labels, torch.int32).reshape(-1)
        out = torch.empty((labels.numel(), self.D), dtype=torch.float32, device=self.device)
        pb = self._dev(pbar) if pbar is not None else None
        check(lib().rbnn_input_grad_sum_kept(self._h, int(head), C.c_void_p(labels.data_ptr()),
                                             C.c_void_p(pb.data_ptr() if pb is not None else 0),
                                             C.c_void_p(out.data_ptr()), _stream(self.device)))
        return out

    def forward_logits(self, x, s):
        x, B = self._x(x)
        out = torch.empty((B, self.n_classes), dtype=torch.float32, device=self.device)
        check(lib().rbnn_forward_logits(self._h, C.c_void_p(x.data_ptr()), B, int(s), C.c_void_p(out.data_ptr()),
                                        _stream(self.device)))
        return out

    def forward_logits_sum(self, x, s0, s1):
        """sum_s f_s(x) (LOGITS) over bank rows [s0, s1): Ensemble_NN.forward x size (model_ensemble.py:57-67)."""
        x, B = self._x(x)
        out = torch.empty((B, self.n_classes), dtype=torch.float32, device=self.device)
        check(lib().rbnn_forward_logits_sum(self._h, C.c_void_p(x.data_ptr()), B, int(s0), int(s1),
                                            C.c_void_p(out.data_ptr()), _stream(self.device)))
        return out

    def input_grad_sum(self, head, x, labels, s0, s1, pbar=None, out=None):
        x, B = self._x(x)
        labels = self._dev(labels, torch.int32).reshape(-1)
        if labels.numel() != B:
            raise ValueError("need one label per input")
        if out is None:
            out = torch.empty_like(x)
        elif (out.device != self.device or out.dtype != torch.float32 or not out.is_contiguous()
              or out.numel() != x.numel()):
            raise ValueError("out must be a contiguous fp32 tensor of %d elements on %s" % (x.numel(), self.device))
        pb = None
        if pbar is not None:
            pb = self._dev(pbar)
        check(lib().rbnn_input_grad_sum(self._h, int(head), C.c_void_p(x.data_ptr()), C.c_void_p(labels.data_ptr()),
                                        B, int(s0), int(s1), C.c_void_p(pb.data_ptr() if pb is not None else 0),
                                        C.c_void_p(out.data_ptr()), _stream(self.device)))
        return out

    def loss_gradients_host(self, x_host, labels_host, s0, s1, n_samples_global, out_host=None):
        """e2e entry: host buffers in, host buffer out (H2D + compute + D2H inside the call)."""
        x = torch.as_tensor(x_host, dtype=torch.float32).contiguous()
        y = torch.as_tensor(labels_host, dtype=torch.int32).contiguous()
        B = x.shape[0]
        if out_host is None:
            out_host = torch.empty((B, self.D), dtype=torch.float32)
        check(lib().rbnn_loss_gradients_host(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), B, int(s0),
                                             int(s1), int(n_samples_global), C.c_void_p(out_host.data_ptr())))
        return out_host


# ---- stateless kernels ------------------------------------------------------------------------
def fgsm_step(x, grad, eps):
    out = torch.empty_like(x)
    check(lib().rbnn_fgsm_step(C.c_void_p(x.data_ptr()), C.c_void_p(grad.data_ptr()), C.c_float(eps),
                               C.c_void_p(out.data_ptr()), x.numel(), _stream(x.device)))
    return out


def pgd_alpha(x):
    B = x.shape[0]
    alpha = torch.empty((B,), dtype=torch.float32, device=x.device)
    check(lib().rbnn_pgd_alpha(C.c_void_p(x.data_ptr()), C.c_void_p(alpha.data_ptr()), B, x.numel() // max(B, 1),
                               _stream(x.device)))
    return alpha


def pgd_step(x, x0, grad, alpha, eps):
    B = x.shape[0]
    out = torch.empty_like(x)
    check(lib().rbnn_pgd_step(C.c_void_p(x.data_ptr()), C.c_void_p(x0.data_ptr()), C.c_void_p(grad.data_ptr()),
                              C.c_void_p(alpha.data_ptr()), C.c_float(eps), C.c_void_p(out.data_ptr()), B,
                              x.numel() // max(B, 1), _stream(x.device)))
    return out


def softmax_robustness(o0, o1):
    """Returns (rob[N], minmax[2]) on the device."""
    N, Cc = o0.shape
    rob = torch.empty((N,), dtype=torch.float32, device=o0.device)
    mm = torch.empty((2,), dtype=torch.float32, device=o0.device)
    check(lib().rbnn_softmax_robustness(C.c_void_p(o0.data_ptr()), C.c_void_p(o1.data_ptr()), N, Cc,
                                        C.c_void_p(rob.data_ptr()), C.c_void_p(mm.data_ptr()), _stream(o0.device)))
    return rob, mm


def count_correct(out, labels_i32, counter):
    """counter (int64[1] on device) += #{argmax(out) == labels}."""
    N, Cc = out.shape
    check(lib().rbnn_count_correct(C.c_void_p(out.data_ptr()), C.c_void_p(labels_i32.data_ptr()), N, Cc,
                                   C.c_void_p(counter.data_ptr()), _stream(out.device)))


# the stateless kernels are also reachable through a Net (lets tests swap the whole engine)
Net.fgsm_step = staticmethod(fgsm_step)
Net.pgd_alpha = staticmethod(pgd_alpha)
Net.pgd_step = staticmethod(pgd_step)
Net.softmax_robustness = staticmethod(softmax_robustness)
Net.count_correct = staticmethod(count_correct)
