/*
 * rbnn.h -- C ABI of librbnn.so: the B200-native replacement for the one
 * data-parallel hot path of ginevracoal/robustBNNs (Bayesian expected loss
 * gradient, Bayesian FGSM/PGD, attack evaluation).
 *
 * The reference is pure Python and has no FFI; its boundary for this path is a
 * set of Python callables.  Each entry point below names the reference
 * interface it stands behind (file:line under the reference repo).  The host
 * side above this ABI (robustbnns_b200/*.py) mirrors those callables one to one.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message
 *     of the last failure on the calling thread is rbnn_last_error().
 *   - no exceptions cross the ABI; no torch / C++ types in any signature.
 *   - pointers whose name starts with d_ are DEVICE pointers on the net's
 *     device, h_ are HOST pointers; `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream).  All device work is enqueued on
 *     `stream` and is asynchronous unless the comment says otherwise.
 *   - the caller owns every buffer it passes; the library owns only the opaque
 *     rbnn_net handle (posterior-sample bank + workspaces).  One handle per
 *     device; a handle is not thread-safe.
 *   - all real arithmetic is fp32 (the reference's dtype); labels and counts
 *     are integers.
 */
#ifndef RBNN_H_
#define RBNN_H_

#include <stdint.h>

#if defined(__GNUC__)
#define RBNN_API __attribute__((visibility("default")))
#else
#define RBNN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rbnn_net rbnn_net;

/* NN.set_model architectures, model_nn.py:77-106 ("conv2", :108-121, is broken upstream). */
enum { RBNN_ARCH_FC = 0, RBNN_ARCH_FC2 = 1, RBNN_ARCH_CONV = 2 };

/* GEMM engines.  FP32 = CUDA-core FFMA (reference-class rounding, any shape);
 * TF32X3 = tcgen05 kind::tf32 with a 3-term split (fp32-class accuracy; arch fc / fc2, and arch conv, where
 *          conv2 runs as an implicit GEMM over 5-D TMA boxes and its input gradient as a tcgen05 GEMM);
 * BF16 = tcgen05 kind::f16 single pass (throughput mode, NOT parity-grade);
 * F16X3 = tcgen05 kind::f16 with a 3-term split of power-of-two-scaled fp16 hi/lo
 *         operands (fp32-class accuracy at twice the TF32X3 rate; arch fc and arch conv). */
enum { RBNN_PREC_FP32 = 0, RBNN_PREC_TF32X3 = 1, RBNN_PREC_BF16 = 2, RBNN_PREC_F16X3 = 3 };
/* Hidden-layer activation, NN.set_model model_nn.py:66-75 ("leaky" = nn.LeakyReLU(), the only one the saved models use,
 * model_bnn.py:36-66, and the default). */
enum { RBNN_ACT_LEAKY = 0, RBNN_ACT_RELU = 1, RBNN_ACT_SIGM = 2, RBNN_ACT_TANH = 3 };

/* Which scalar loss the input gradient is taken of. */
enum {
  RBNN_HEAD_MEAN_OF_GRADS = 0, /* lossGradients.py:29-40: per-sample CE(softmax(softmax(z_s))) */
  RBNN_HEAD_GRAD_OF_MEAN = 1,  /* adversarialAttacks.py:74-78: CE(softmax(mean_s softmax(z_s))) */
  RBNN_HEAD_LOGITS_CE = 2,     /* avg_posterior=True, model_bnn.py:206-216: CE(z) on the mean-weight net */
  RBNN_HEAD_UPSTREAM = 3,      /* autograd through BNN.forward: d_pbar holds dL/d(mean probs) [B,C] itself */
  RBNN_HEAD_LOGITS_UPSTREAM = 4 /* Ensemble_NN.forward (mean of LOGITS, model_ensemble.py:57-67): d_pbar holds
                                   dL/d(sum of logits) [B,C]; every bank row receives it unchanged. */
};

RBNN_API int rbnn_abi_version(void);
RBNN_API const char* rbnn_last_error(void);

/* ---- a1: the network f_w  (NN.__init__/set_model, model_nn.py:36-124) -----------------
 * Creates the handle for one architecture on CUDA device `device`.  `hidden` must be a
 * power of two >= 16 (model_nn.py:39-40); conv needs a 1x28x28-shaped input (model_nn.py:95).
 * Only LeakyReLU(0.01) is implemented (every saved model uses it, model_bnn.py:36-66). */
RBNN_API int rbnn_net_create(rbnn_net** out, int arch, int in_ch, int in_h, int in_w, int hidden,
                    int n_classes, int device);
RBNN_API int rbnn_net_destroy(rbnn_net* net);
/* Parameters per posterior sample, in basenet.state_dict() order (model_bnn.py:124). */
RBNN_API int64_t rbnn_net_param_count(const rbnn_net* net);
RBNN_API int rbnn_net_set_precision(rbnn_net* net, int prec);
RBNN_API int rbnn_net_get_precision(const rbnn_net* net);
/* relu / sigm / tanh (model_nn.py:66-73) run on the FP32 CUDA-core engine for arch fc / fc2: every tensor-core kernel
 * and the conv kernels fuse LeakyReLU(0.01) and its guard band.  Selecting another activation switches the handle to
 * RBNN_PREC_FP32; rbnn_net_set_precision then refuses the tensor-core modes, and arch conv is refused here. */
RBNN_API int rbnn_net_set_activation(rbnn_net* net, int act);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
RBNN_API int64_t rbnn_net_launch_count(const rbnn_net* net);
/* Inputs on the 8-bit pixel grid.  Every image set the reference loads is uint8 / 255 (utils.py:102-103, 129-130,
 * 190-191); scaled by 255 * 2^j such inputs are exact in fp16, the low half of their F16X3 split is zero and the fused
 * forward of arch fc issues two tensor-core passes instead of three (detected on the device per call, bit-level test, no
 * tolerance on the data; RBNN_XGRID=0 in the environment disables it).  Returns 1 when the last F16X3 forward of this
 * handle took that route, 0 when not, -1 on error.  Synchronises the device (a test / bench query, not for the hot path). */
RBNN_API int rbnn_net_input_grid(rbnn_net* net);
/* Device timing of the two dominant kernel classes with CUDA events recorded on the launch
 * stream (1 = first-layer forward GEMM, 2 = input-gradient GEMM).  read() synchronises the
 * device, returns the summed kernel time and launch count since the last read, and resets. */
RBNN_API int rbnn_net_timing_enable(rbnn_net* net, int on);
RBNN_API int rbnn_net_timing_read(rbnn_net* net, int cls, double* total_ms, int64_t* launches);

/* ---- a2/a4: the posterior-sample bank ---------------------------------------------------
 * The bank is `capacity` rows of P fp32 parameters resident in HBM, row s = the weights of
 * posterior sample s in state_dict order, plus kernel-ready copies derived from it. */
RBNN_API int rbnn_bank_reserve(rbnn_net* net, int capacity);
RBNN_API int rbnn_bank_capacity(const rbnn_net* net);
/* BNN.load HMC branch / stacked state dicts (model_bnn.py:184-190): copy `count` rows
 * from `weights` ([count, P], host if is_device==0 else device) into rows [s0, s0+count). */
RBNN_API int rbnn_bank_upload(rbnn_net* net, const float* weights, int s0, int count, int is_device, void* stream);
/* BNN.guide (model_bnn.py:121-130): rows [s0, s0+count) <- loc + softplus(rho) * eps where
 * eps of row s0+i is the Philox4x32-10 stream of GLOBAL sample index sample_index0 + i*sample_index_stride under
 * `seed` (independent of how samples are sharded over GPUs). d_loc/d_rho: [P] device fp32. */
RBNN_API int rbnn_bank_sample_diag(rbnn_net* net, const float* d_loc, const float* d_rho, uint64_t seed,
                          int64_t sample_index0, int64_t sample_index_stride, int s0, int count, void* stream);
/* The same draw with a DEVICE-resident addend: row s0+i uses global sample index sample_index0 + *d_index_offset +
 * i*sample_index_stride, *d_index_offset being read when the kernel runs (NULL = 0).  This is what lets one captured
 * CUDA graph of a PGD iteration draw FRESH posterior samples on every replay, as the reference's attacks do
 * (adversarialAttacks.py:95-105 calls net.forward, i.e. a new guide draw, model_bnn.py:230-232, in every iteration): the
 * caller advances the device counter between replays. */
RBNN_API int rbnn_bank_sample_diag_at(rbnn_net* net, const float* d_loc, const float* d_rho, uint64_t seed,
                             int64_t sample_index0, int64_t sample_index_stride, int s0, int count,
                             const int64_t* d_index_offset, void* stream);
/* A NEW posterior is about to be installed in this handle (other guide parameters / another stored bank; BNN.load,
 * model_bnn.py:167-196): forget what was derived from the old weights -- the frozen power-of-two operand scale of the
 * F16X3 engine, the kernel-ready copies, the kept forward.  Synchronous.  Uploads and draws that follow fix a new scale
 * from the rows they touch (a range overflow under a frozen scale is otherwise caught and repaired inside the call that
 * re-lays the rows). */
RBNN_API int rbnn_bank_invalidate(rbnn_net* net);
/* Counter bumped whenever a device buffer owned by the handle is (re)allocated.  A CUDA graph captured over calls on
 * this handle must be re-captured when the value has changed (its nodes hold the old addresses). */
RBNN_API int64_t rbnn_net_alloc_epoch(const rbnn_net* net);
/* Copy rows [s0, s0+count) back to host (synchronous; test support). */
RBNN_API int rbnn_bank_download(rbnn_net* net, float* h_out, int s0, int count);

/* ---- a3/a5/a11: BNN.forward (model_bnn.py:198-258) --------------------------------------
 * d_out_sum[B, C] <- sum over bank rows s in [s0, s1) of softmax(f_{w_s}(x)).  The caller
 * divides by the GLOBAL number of samples after the cross-GPU allreduce.  d_x: [B, D]. */
RBNN_API int rbnn_forward_probs_sum(rbnn_net* net, const float* d_x, int B, int s0, int s1,
                           float* d_out_sum, void* stream);
/* Two-phase form of the attack gradient (adversarialAttacks.py:74-78 evaluates net.forward ONCE and differentiates
 * it): the same sum as rbnn_forward_probs_sum, and the per-sample logits and LeakyReLU masks of this call stay in the
 * handle (~104 B per sample x input), so that rbnn_input_grad_sum_kept can rebuild the gradient without a second
 * forward pass (arch conv keeps the pooled conv1 map, its arg-max indices, the refined second-layer activations and the
 * logits, fc2 the refined hidden activations).  When the engine has no such route (FP32 engine, batch too large for one pass) the call
 * is a plain forward and rbnn_keep_valid() returns 0.  The kept data is invalidated when the bank rows it used are
 * overwritten or the precision changes. */
RBNN_API int rbnn_forward_probs_sum_keep(rbnn_net* net, const float* d_x, int B, int s0, int s1,
                                float* d_out_sum, void* stream);
RBNN_API int rbnn_keep_valid(const rbnn_net* net);
/* d_out_sum[B, D] <- sum over the kept rows of dL_s/dx for the kept inputs; head = GRAD_OF_MEAN / UPSTREAM (with
 * d_pbar as in rbnn_input_grad_sum) or MEAN_OF_GRADS. */
RBNN_API int rbnn_input_grad_sum_kept(rbnn_net* net, int head, const int32_t* d_labels, const float* d_pbar,
                             float* d_out_sum, void* stream);
/* avg_posterior=True (model_bnn.py:206-216): LOGITS of bank row s. d_out: [B, C]. */
RBNN_API int rbnn_forward_logits(rbnn_net* net, const float* d_x, int B, int s, float* d_out, void* stream);
/* Ensemble_NN.forward (model_ensemble.py:57-67): d_out_sum[B, C] <- sum over bank rows s in [s0, s1) of the LOGITS
 * f_{w_s}(x); the caller divides by the ensemble size.  Deterministic NN.forward (model_nn.py:126-141) is the
 * one-row case. */
RBNN_API int rbnn_forward_logits_sum(rbnn_net* net, const float* d_x, int B, int s0, int s1, float* d_out_sum,
                            void* stream);

/* ---- a6/a7/a8/a9: input gradients ---------------------------------------------------------
 * d_out_sum[B, D] <- sum over rows s in [s0, s1) of dL_s/dx with L chosen by `head`:
 *   MEAN_OF_GRADS: L_s = CE(softmax(softmax(z_s)), y)           (lossGradients.py:33-36)
 *   GRAD_OF_MEAN : L   = CE(softmax(pbar), y), pbar = d_pbar[B,C] = the mean over ALL samples
 *                  of softmax(z_s) (already allreduced and divided by the global S); the
 *                  per-sample term is p_s*(g-<p_s,g>) with g = softmax(pbar)-e_y; the caller
 *                  multiplies the final sum by 1/S                (adversarialAttacks.py:74-78)
 *   LOGITS_CE    : L = CE(z_s, y) for the single row s0           (model_bnn.py:206-216)
 *   UPSTREAM     : like GRAD_OF_MEAN but g = d_pbar[B,C] is given (any loss on BNN.forward's output)
 *   LOGITS_UPSTREAM: dL/dz_s = d_pbar[B,C] for every row (any loss on the SUM / mean of the logits: Ensemble_NN,
 *                  deterministic NN; model_ensemble.py:57-67).
 * Sum reduction over the batch (the reference batch is always 1, so no 1/B).
 * d_labels: [B] int32 class indices. d_pbar may be NULL unless head is GRAD_OF_MEAN / UPSTREAM / LOGITS_UPSTREAM. */
RBNN_API int rbnn_input_grad_sum(rbnn_net* net, int head, const float* d_x, const int32_t* d_labels, int B,
                        int s0, int s1, const float* d_pbar, float* d_out_sum, void* stream);

/* ---- a8/a9: the attack update (adversarialAttacks.py:81-82, :103-105) ----------------------
 * FGSM : out = clamp(x + eps*sign(g*scale), 0, 1)
 * PGD  : out = clamp(x0 + clamp(x + alpha_b*sign(g) - x0, -eps, eps), 0, 1), alpha per image
 *        (d_alpha[B]; 2/image.max() or 2/225, adversarialAttacks.py:89,91).  n = B*D. */
RBNN_API int rbnn_fgsm_step(const float* d_x, const float* d_grad, float eps, float* d_out, int64_t n, void* stream);
RBNN_API int rbnn_pgd_step(const float* d_x, const float* d_x0, const float* d_grad, const float* d_alpha,
                  float eps, float* d_out, int B, int D, void* stream);
/* alpha_b = 2 / max_d x[b, d]  (adversarialAttacks.py:89). */
RBNN_API int rbnn_pgd_alpha(const float* d_x, float* d_alpha, int B, int D, void* stream);

/* ---- a11/a12: evaluation (adversarialAttacks.py:30-62, :179, :186) -------------------------
 * rob[n] = 1 - max_c |softmax(o0[n,:]) - softmax(o1[n,:])|_c ; d_minmax[2] receives the min and
 * max of the L-inf differences so the host can raise the reference's range ValueError (:48-49). */
RBNN_API int rbnn_softmax_robustness(const float* d_o0, const float* d_o1, int N, int C, float* d_rob,
                            float* d_minmax, void* stream);
/* *d_count += #{n : argmax_c out[n,c] == labels[n]} (first max wins, as torch.argmax). */
RBNN_API int rbnn_count_correct(const float* d_out, const int32_t* d_labels, int N, int C, int64_t* d_count,
                       void* stream);

/* ---- e2e convenience: the whole of loss_gradients (lossGradients.py:52-66) from HOST buffers.
 * h_x [B, D] and h_labels [B] are copied to the device, the expected loss gradient over bank
 * rows [s0, s1) is evaluated and h_out [B, D] <- (1/n_samples_global) * sum.  Synchronous. */
RBNN_API int rbnn_loss_gradients_host(rbnn_net* net, const float* h_x, const int32_t* h_labels, int B,
                             int s0, int s1, int n_samples_global, float* h_out);

#ifdef __cplusplus
}
#endif
#endif /* RBNN_H_ */
