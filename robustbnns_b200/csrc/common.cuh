// Shared declarations for librbnn.so (see include/rbnn.h for the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/rbnn.h"

namespace rbnn {

constexpr float kLeakySlope = 0.01f;  // nn.LeakyReLU() default (model_nn.py:68-69)

void set_error(const char* fmt, ...);

#define RBNN_CUDA(expr)                                                                 \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      rbnn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

#define RBNN_CHECK(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      rbnn::set_error(__VA_ARGS__);    \
      return 1;                        \
    }                                  \
  } while (0)

#define RBNN_TRY(expr)        \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

// Offsets (in floats) of each state_dict tensor inside one bank row.
struct ParamLayout {
  // fc / fc2: w1 [H,D] b1 [H] (w2 [H,H] b2 [H])? wo [C,H] bo [C]
  // conv    : cw1 [32,1,5,5] cb1 [32] cw2 [H,32,5,5] cb2 [H] wo [C,49H] bo [C]
  int64_t w1 = 0, b1 = 0, w2 = 0, b2 = 0, wo = 0, bo = 0;
  int64_t cw1 = 0, cb1 = 0, cw2 = 0, cb2 = 0;
  int64_t P = 0;
};

}  // namespace rbnn

// Derived, kernel-ready copies of one weight matrix of the bank for the tcgen05 path (tc_fc.cu):
// K-major operands for the forward GEMM ([S, R, C]) and, transposed, for the backward GEMM ([S, C, R]).
struct TcMat {
  int64_t off = 0;          // offset of the [R, C] matrix inside a bank row
  int R = 0, C = 0;
  int ld = 0;               // row pitch (elements) of the forward copies: C rounded up to whole 128-byte lines
  int perm25 = 0;           // conv2 filters [H][32][5][5]: the forward copy stores a row's 800 columns tap-major,
                            // (ky, kx, c) instead of (c, ky, kx) -- the K order of the implicit GEMM (tc_conv.cu)
  float *hi = nullptr, *lo = nullptr;      // tf32 split:  w = hi + lo, hi = rn_tf32(w)
  float *thi = nullptr, *tlo = nullptr;    // the same, transposed
  void *bf = nullptr, *tbf = nullptr;      // bf16 variants
  void *h_hi = nullptr, *h_lo = nullptr;   // F16X3: fp16 split of s_w * w (s_w = TcScales::s_w1, a power of two)
  void *th_hi = nullptr, *th_lo = nullptr; // the same, transposed
};

// Device-resident operand-range bookkeeping of the F16X3 engine (fp16 has 5 exponent bits: operands are scaled by
// powers of two so that their largest element sits in [2^8, 2^9)).  The host never reads it.
struct TcScales {
  unsigned maxw1_bits;   // max |W1| over every bank row re-laid so far (float bits; monotone)
  unsigned maxwo_bits;   // max |Wo| likewise (bounds dH)
  int frozen;            // s_w1 has been fixed (the derived copies of clean rows depend on it)
  float s_w1;            // power-of-two scale of the fp16 copies of W1 / W1^T
};

struct TcBank {
  int capacity = 0;
  int mode = -1;            // precision the copies were built for
  int nmat = 0;
  TcMat mat[2];
  float* wnorm = nullptr;   // [capacity] max_j ||W1_s[j,:]||_2 (guard band of the fused forward kernel)
  TcScales* scales = nullptr;     // device (F16X3)
  unsigned* xgrid_last = nullptr; // device (F16X3): 1 = the last forward found its inputs on the pixel grid (two-pass forward)
  int* overflow_host = nullptr;   // mapped pinned flag: a later row exceeded the fp16 range under the frozen s_w1
  int* overflow_dev = nullptr;    // device alias of overflow_host
  uint8_t* dirty = nullptr; // host flags per row (derived copies stale)
  int frozen_host = 0;      // the F16X3 weight scale has been fixed on the device (a freeze_scales launch is enqueued)
};

// Kept forward of the two-phase attack gradient (tc_gemm.cuh): per-sample logits and LeakyReLU masks of the last
// rbnn_forward_probs_sum_keep call, valid until the bank rows it used change or the precision is switched.
struct KeepCache {
  float* logits = nullptr; size_t logits_cap = 0;       // [S][B][C] floats
  uint32_t* masks = nullptr; size_t masks_cap = 0;      // tc::keep_mask_words(B, S) words
  float* call_sc = nullptr;                             // F16X3: the call's 4 scale scalars, alive between the phases
  unsigned* max_bits = nullptr;                         // F16X3: [0] max|x|, [1] max|d_pbar|
  // unfused FC route (fc2, or arch fc with RBNN_TC_UNFUSED): the hidden activations H1 [, H2] of every kept unit, fp32
  float* fc_h = nullptr; size_t fc_cap = 0;
  // conv engine (tc_conv.cu): pooled conv1 map, its arg-max indices, the refined A2 and the logits of every kept unit
  char* conv_buf = nullptr; size_t conv_cap = 0;
  int valid = 0, B = 0, s0 = 0, s1 = 0;
};

struct rbnn_net {
  int arch = 0, in_ch = 1, in_h = 28, in_w = 28, D = 784, H = 512, C = 10;
  int device = 0;
  int prec = RBNN_PREC_FP32;
  int act = RBNN_ACT_LEAKY;   // hidden-layer activation (model_nn.py:66-75); anything but LeakyReLU: FP32 engine, arch fc / fc2
  rbnn::ParamLayout L;
  // bank
  int capacity = 0;
  float* bank = nullptr;      // [capacity, P]
  float* woutp = nullptr;     // conv only: [capacity][49H (pos, h)][CP] output weights, classes innermost and zero-padded to
                              // CP = conv_class_pitch(C) so that one (pos, h) entry is CP/4 16-byte loads (sampler.cu)
  float* sigma = nullptr;     // [P] scratch: softplus(rho) of the last sample_diag call
  // workspace arena
  char* ws = nullptr;
  size_t ws_bytes = 0;
  size_t ws_budget = (size_t)12 << 30;   // upper bound of the activation workspace (HBM is 180 GB)
  int64_t launches = 0;
  int64_t alloc_epoch = 0;    // bumped whenever a device buffer of this handle is (re)allocated: captured CUDA graphs of
                              // earlier launches hold stale pointers afterwards (rbnn_net_alloc_epoch)
  int sm_count = 148;
  int cc_major = 0;           // compute capability major of `device` (10 = Blackwell: tcgen05 engine usable)
  TcBank tc;
  KeepCache keep;
  int conv1_smem_set = 0;     // conv1_bwd_sum_kernel's dynamic shared-memory attribute has been set on this device
  int sum_logits = 0;         // 1 while rbnn_forward_logits_sum runs: the forward passes accumulate raw logits, not softmax rows
  int tc_unfused = 0;         // 1: arch fc takes the unfused GEMM -> head route (RBNN_TC_UNFUSED=1; A/B testing)
  // optional per-kernel-class device timing (bench.py's roofline leg): event pairs on the launch stream
  int timing = 0;
  std::vector<cudaEvent_t> ev[3][2];   // [class][begin/end]; class 1 = first-layer forward GEMM, 2 = input-grad GEMM
};

namespace rbnn {

int ws_reserve(rbnn_net* net, size_t bytes);

struct Arena {
  char* base;
  size_t off = 0, cap;
  Arena(rbnn_net* n) : base(n->ws), cap(n->ws_bytes) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};

inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

// ---- gemm_simt.cu -----------------------------------------------------------------------
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_LEAKY = 2, EPI_MASK = 3 };

struct GemmArgs {
  const float* A; int64_t lda, sAz;      // A[z][m][k], k contiguous
  const float* B; int64_t ldb, sBz;      // NT: B[z][n][k]   NN: B[z][k][n]
  float* C; int64_t ldc, sCz;            // C[z][m][n]
  const float* bias; int64_t sbz;        // bias[z][n]             (EPI_BIAS*)
  const float* mask; int64_t ldm, sMz;   // activation[z][m][n]    (EPI_MASK: acc *= act>0 ? 1 : slope)
  int M, N, K, Z;
  int epi;
  int act;         // RBNN_ACT_*: EPI_BIAS_LEAKY applies it, EPI_MASK multiplies by its derivative (from the stored activation)
  int b_kn;        // 0 = NT, 1 = NN
  int reduce_z;    // 1: C[m][n] (+)= sum_z ...  (C is a single [M,N] matrix)
  int accumulate;  // reduce_z only: add to what C already holds
  int tag;         // timing class (0 = untimed)
};
int timing_begin(rbnn_net* net, int cls, cudaStream_t st);
int timing_end(rbnn_net* net, int cls, cudaStream_t st);
int gemm_simt(rbnn_net* net, const GemmArgs& a, cudaStream_t st);

// ---- head.cu ----------------------------------------------------------------------------
int head_probs_accumulate(rbnn_net* net, const float* logits, int Z, int B, int C, float* out_sum,
                          cudaStream_t st);
int head_dlogits(rbnn_net* net, int head, const float* logits, const int32_t* labels, const float* pbar,
                 int Z, int B, int C, float* dlogits, cudaStream_t st);

// ---- sampler.cu -------------------------------------------------------------------------
// d_index_offset (may be nullptr): device int64 added to every sample index when the kernel RUNS -- lets a captured
// CUDA graph draw fresh samples on every replay (the caller advances the device counter between replays)
int sample_diag(rbnn_net* net, const float* d_loc, const float* d_rho, uint64_t seed, int64_t sample_index0,
                int64_t stride, int s0, int count, cudaStream_t st, const int64_t* d_index_offset = nullptr);
int sample_sigma(rbnn_net* net, const float* d_rho, cudaStream_t st);      // net->sigma = softplus(rho)
int sample_diag_from(rbnn_net* net, const float* d_loc, uint64_t seed, int64_t sample_index0, int64_t stride, int s0,
                     int count, int64_t elem0, cudaStream_t st,             // elements [elem0, P) of the rows
                     const int64_t* d_index_offset = nullptr);
int conv_permute_wout(rbnn_net* net, int s0, int count, cudaStream_t st);
inline int conv_class_pitch(int C) { return C <= 4 ? 4 : (C <= 12 ? 12 : (C <= 16 ? 16 : 32)); }

// ---- conv.cu ----------------------------------------------------------------------------
int conv1_pool_fwd(rbnn_net* net, const float* x, const float* bank, int s0, int Z, int B, float* p1,
                   uint8_t* idx1, cudaStream_t st, float* p1lo = nullptr, unsigned* max_bits = nullptr);
int im2col_conv2(rbnn_net* net, const float* p1, int ZB, float* col, cudaStream_t st);
int pool2_fwd(rbnn_net* net, const float* a2, int ZB, int H, float* p2, cudaStream_t st);
int pool2_bwd(rbnn_net* net, const float* a2, const float* dp2, int ZB, int H, float* dz2, cudaStream_t st,
              float* dz2_lo = nullptr);
// f16_scale != nullptr: hi / lo are fp16 arrays holding the split of *f16_scale * P1 (F16X3), else tf32-split fp32
int p1_split_hwc(rbnn_net* net, const float* p1, int ZB, void* hi, void* lo, const float* f16_scale, cudaStream_t st);
// unit_max: [Z * B] float bits, max |conv2 pre-activation| of every (sample, image) unit (tc::GemmDesc::group_max)
int conv2_refine(rbnn_net* net, float* a2, const float* p1, int s0, int Z, int B, float eps, cudaStream_t st,
                 const float* p1lo, const unsigned* unit_max);
int col2im_conv2(rbnn_net* net, const float* dcol, const float* p1, int ZB, float* g1, cudaStream_t st);
// partial != nullptr && parts > 1: the sample range is cut into `parts` slices (partial: [parts][B][784] floats)
int conv1_bwd_sum(rbnn_net* net, const float* g1, const uint8_t* idx1, const float* bank, int s0, int Z, int B,
                  float* dx_sum, int accumulate, cudaStream_t st, float* partial = nullptr, int parts = 1);
int conv1_bwd_parts(const rbnn_net* net, int Z, int B);
// MaxPool2d(2, stride 1) + Linear(49H, C) and their input gradient, fused (the pooled map / its gradient stay on chip)
int pool2_logits_chunks(const rbnn_net* net);
int pool2_logits(rbnn_net* net, const float* a2, int s0, int Z, int B, float* logits, float* partial, cudaStream_t st);
// dz2_lo == nullptr: fp32 result; f16_scale == nullptr: tf32 hi / lo (fp32 arrays); else fp16 hi / lo of *f16_scale * dZ2
int pool2_bwd_fused(rbnn_net* net, const float* a2, const float* dlogits, int s0, int Z, int B, void* dz2,
                    void* dz2_lo, const float* f16_scale, cudaStream_t st);

// ---- attack.cu --------------------------------------------------------------------------
int scale_inplace(rbnn_net* net, float* p, float scale, int64_t n, cudaStream_t st);
int add_inplace(rbnn_net* net, float* dst, const float* src, int64_t n, cudaStream_t st);

// ---- tc_fc.cu (tcgen05 FC path) -----------------------------------------------------------
int tc_supported(const rbnn_net* net);
int tc_f16x3_supported(const rbnn_net* net);
void tc_bank_free(rbnn_net* net);
int tc_fc_input_grad_sum(rbnn_net* net, int head, const float* d_x, const int32_t* d_labels, int B, int s0, int s1,
                         const float* d_pbar, float* d_out_sum, cudaStream_t st);
// out_sum != nullptr: out_sum[B,C] += sum_s softmax(logits_s); out_logits != nullptr (one row): logits of row s0
int tc_fc_forward(rbnn_net* net, const float* d_x, int B, int s0, int s1, float* d_out_sum, float* d_out_logits,
                  cudaStream_t st);
// two-phase attack gradient: forward that keeps logits + masks (net->keep.valid tells whether it could), then the
// gradient of a loss of the mean prediction from the kept data
int tc_fc_forward_keep(rbnn_net* net, const float* d_x, int B, int s0, int s1, float* d_out_sum, cudaStream_t st);
int tc_fc_grad_kept(rbnn_net* net, int head, const int32_t* d_labels, const float* d_pbar, float* d_out_sum,
                    cudaStream_t st);
void tc_keep_free(rbnn_net* net);
// derived tensor-core operand copies of bank rows [s0, s1) brought up to date (lazily, per dirty row)
int tc_bank_refresh(rbnn_net* net, int s0, int s1, cudaStream_t st);
// K-sample fused with the operand re-layout (arch fc, F16X3, weight scale already fixed): draws rows [s0, s0+count) and
// writes the bank AND the fp16 hi/lo operand copies of W1 in one pass.  Returns 0 and sets *done = 1 when it ran;
// *done = 0 means "not applicable, use sample_diag + the lazy refresh".
int tc_sample_relayout_f16(rbnn_net* net, const float* d_loc, const float* d_rho, uint64_t seed, int64_t sample_index0,
                           int64_t stride, int s0, int count, cudaStream_t st, int* done,
                           const int64_t* d_index_offset = nullptr);
// a new posterior is being installed: forget the frozen F16X3 operand scale, mark every derived copy stale
int tc_bank_invalidate(rbnn_net* net);

// F16X3 operand ranges (device-resident): *bits = max(*bits, max|p|) as float bits; out[0..3] = s_x, 1/(s_x s_w),
// s_d, 1/(s_d s_w) with s_d from the bound dh_factor * max|g| * max|Wo| (call_scales_kernel, tc_fc.cu)
int tc_maxabs(rbnn_net* net, const float* p, int64_t count, unsigned* bits, cudaStream_t st);
int tc_call_scales(rbnn_net* net, const unsigned* xmax_bits, const unsigned* gmax_bits, float dh_factor, float* out,
                   cudaStream_t st);

// ---- tc_conv.cu (tcgen05 conv path, TF32X3 / F16X3) -------------------------------------------
int tc_conv_supported(const rbnn_net* net);
int tc_conv_input_grad_sum(rbnn_net* net, int head, const float* d_x, const int32_t* d_labels, int B, int s0, int s1,
                           const float* d_pbar, float* d_out_sum, cudaStream_t st);
int tc_conv_forward(rbnn_net* net, const float* d_x, int B, int s0, int s1, float* d_out_sum, float* d_out_logits,
                    cudaStream_t st);
// two-phase attack gradient for the conv net: the forward keeps P1 / arg-max indices / refined A2 / logits per unit
// (net->keep.valid tells whether it could), the gradient pass starts from them -- no second forward
int tc_conv_forward_keep(rbnn_net* net, const float* d_x, int B, int s0, int s1, float* d_out_sum, cudaStream_t st);
int tc_conv_grad_kept(rbnn_net* net, int head, const int32_t* d_labels, const float* d_pbar, float* d_out_sum,
                      cudaStream_t st);

}  // namespace rbnn
