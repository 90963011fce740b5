// tcgen05 grouped GEMM for sm_100a: the tensor-core engine behind RBNN_PREC_TF32X3 / RBNN_PREC_BF16.
//
//   per posterior sample z:   C_z[M,N] = epi( A_z[M,K] . B_z[N,K]^T )            (reduce_z = 0)
//   over a range of samples:  C  [M,N] = sum_z A_z[M,K] . B_z[N,K]^T            (reduce_z = 1)
//
// It stands in for nn.Linear forward (model_nn.py:80,89) and autograd's input gradient of nn.Linear
// for every posterior sample at once; the sum over samples (lossGradients.py:40) is a concatenated-K
// accumulation in TMEM.  Both operands are K-major; operand tiles are TMA-staged into 128B-swizzled
// shared memory, the accumulator lives in TMEM (double buffered, 2 x 256 columns), one thread issues
// tcgen05.mma, four epilogue warps drain TMEM with tcgen05.ld and apply bias / LeakyReLU / mask.
//
// Precision modes
//   TF32X3 : every fp32 operand is pre-split into hi = rn_tf32(x) and lo = x - hi (both fp32 arrays);
//            D += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi with kind::tf32 => fp32-class accuracy (~2^-21).
//   BF16   : single kind::f16 pass on bf16 operands (throughput mode, not parity grade).
//   F16X3  : every fp32 operand is scaled by a power of two (so that its largest element sits near 2^9) and
//            pre-split into fp16 hi = rn_f16(s x) and lo = rn_f16(s x - hi); the same 3-term sum with kind::f16
//            MMAs on fp16 operands => the same ~22 significant bits as TF32X3 at twice the MMA rate and half the
//            operand bytes.  The epilogue multiplies the accumulator by *unscale = 1 / (s_A s_B) (device scalar).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace rbnn {
namespace tc {

constexpr int kBM = 128;            // tile rows = TMEM lanes
constexpr int kBNMax = 256;         // tile columns (UMMA N) upper bound; runtime BN is a multiple of 16
constexpr int kThreads = 192;       // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kTmemCols = 512;      // 2 accumulator stages x 256 fp32 columns

enum { EPI_NONE = 0, EPI_BIAS_LEAKY = 1, EPI_MASK = 2, EPI_BIAS = 3 };
enum { MODE_TF32X3 = 0, MODE_BF16 = 1, MODE_F16X3 = 2 };
enum { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };   // element type of a TMA map

// One GEMM operand: [Z][rows][K] with K contiguous.  `lo` is the residual array (TF32X3: fp32, F16X3: fp16).
struct Operand {
  const void* hi = nullptr;
  const void* lo = nullptr;
  int64_t rows = 0;      // M for A, N for B
  int64_t ld = 0;        // elements between consecutive rows (>= K, multiple of 16 bytes)
  int64_t zstride = 0;   // elements between consecutive z (0 = shared by all z)
};

struct GemmDesc {
  int mode = MODE_TF32X3;
  int M = 0, N = 0, K = 0, Z = 1;
  int BN = 256;          // tile width
  int kblock_bytes = 128; // bytes of K per pipeline stage row: 128 (default) or 64 (twice the stages, measured slower)
  int pair = 0;          // 1: CTA pairs (tcgen05 cta_group::2, 256-row tiles; measured +5% only); 0: one CTA per 128-row tile
  Operand A, B;
  int reduce_z = 0;      // 1: sum over z; the z range is cut into `slots` contiguous pieces -> out[slot]
  int slots = 1;
  int epi = EPI_NONE;
  const float* unscale = nullptr;   // F16X3: device scalar the accumulator is multiplied by before the epilogue (1 / operand scales)
  const float* bias = nullptr; int64_t bias_zstride = 0;                 // bias[z][n]
  const float* act = nullptr; int64_t act_zstride = 0; int64_t act_ld = 0; // EPI_MASK: acc *= act>0 ? 1 : slope
  // outputs, [z or slot][m][n] with leading dimension out_ld (elements) and z stride out_zstride:
  float* out = nullptr;       // fp32 result (or tf32 hi part when out_lo != nullptr)
  float* out_lo = nullptr;    // optional: tf32 residual -> the result is written pre-split for the next GEMM
  void* out_bf = nullptr;     // optional: bf16 copy of the result
  int64_t out_ld = 0, out_zstride = 0;
  // EPI_BIAS_LEAKY, optional: group_max[z * ceil(M / 64) + m / 64] = max(.., |pre-activation|) as float bits (atomicMax;
  // zero it first) -- the per-image maximum the guard band of the conv net scales with, taken where the accumulator
  // is in registers anyway instead of in a separate pass over the output
  unsigned* group_max = nullptr;
  int sm_count = 148;
  int pair_relay = 1;      // pair mode: 1 = own-barrier TMA + relayed full signal, 0 = cta_group::2 TMA onto the leader's barrier
  int spin_wait = 0;       // 1: poll mbarriers with test_wait instead of the suspending try_wait
  int debug_skip_mma = 0;  // harness only: run the TMA / barrier pipeline without issuing MMAs
  // Implicit-GEMM A operand of the second convolution (model_nn.py:101, Conv2d(32, H, 5) on the pooled 32x12x12 map):
  // conv_images > 0 says A.hi / A.lo are channels-last activations [Z][conv_images][12][12][32] (fp32, tf32-split) and
  // the GEMM row m = image * 64 + oy * 8 + ox is the 5x5x32 patch at output position (oy, ox): K-block kb = (ky, kx)
  // is ONE 5-D TMA box {32 channels, 8 x, 8 y, 2 images} at offset (kx, ky) -- no im2col matrix is ever written.
  // K = 800 ordered (ky, kx, c); B rows must be stored in that order.  Single CTAs; TF32X3 (128-byte K-blocks) or
  // F16X3 with kblock_bytes = 64 (32 fp16 channels; activations [..][32] fp16 hi/lo).
  int conv_images = 0;
};

// Enqueues the GEMM on `st`.  Returns 0 on success; on failure fills *err.
int gemm(const GemmDesc& d, cudaStream_t st, std::string* err);

// 3-D TMA map over [Z][rows][K] (K innermost, `ld` / `zstride` in elements): box = one 128-byte K-block x box_rows x 1,
// SWIZZLE_128B / SWIZZLE_64B (kb_bytes = 128 / 64), zero fill out of bounds.  zstride == 0 => a single z.
int make_map(CUtensorMap* map, const void* base, int dtype, int64_t K, int64_t rows, int64_t Z, int64_t ld,
             int64_t zstride, int box_rows, int kb_bytes, std::string* err);

// Fused forward + head kernel for arch fc (tc_fused.cu): for every (posterior sample z, 128-input tile)
//   H = leaky(X . W1_z^T + b1_z) in TMEM -> logits = H . Wo_z^T + bo_z -> loss head -> dlogits
//   dH = (dlogits . Wo_z) (.) leaky'(H)   written K-major, pre-split (tf32 or scaled fp16 hi/lo) or bf16, for the backward GEMM
// without ever writing H to HBM.  Units whose pre-activation lies inside the guard band are queued on a
// worklist together with the sign that was assumed; fused_fixup() re-evaluates them exactly and patches dH.
struct FusedDesc {
  int mode = MODE_TF32X3;
  int B = 0, D = 0, H = 0, C = 0, Z = 0;
  Operand X, W1;                  // X: [B, D] (zstride 0); W1: [Z][H][D]
  int head = -1;                  // RBNN_HEAD_* ; -1 = write logits only; -2 = keep mode: logits + LeakyReLU masks
  uint32_t* maskbuf = nullptr;    // keep mode: keep_mask_words(B, Z) words
  const float* bank = nullptr;    // bank row of sample z: bank + (z_row0 + z) * P
  int64_t P = 0, b1_off = 0, wo_off = 0, bo_off = 0;
  int z_row0 = 0;
  const int32_t* labels = nullptr;
  const float* pbar = nullptr;
  const float* x = nullptr;       // fp32 inputs [B, D] (exact re-evaluation)
  const float* xnorm = nullptr;   // [B]   ||x_b||_2
  const float* wnorm = nullptr;   // [capacity] max_j ||W1_s[j,:]||_2, indexed by bank row
  float eps = 0.f;                // guard = eps * xnorm[b] * wnorm[row]; 0 disables the worklist
  void* dh_hi = nullptr; void* dh_lo = nullptr; void* dh_bf = nullptr;   // [Z][B][H]; hi/lo: fp32 (TF32X3) or fp16 (F16X3)
  const float* unscale = nullptr;   // F16X3: device scalar 1 / (s_X s_W1) applied to the forward accumulator
  const float* dh_scale = nullptr;  // F16X3: device scalar dH is multiplied by before the fp16 hi/lo split
  const unsigned* xlo_zero = nullptr;   // F16X3: device flag, != 0 = X.lo is all zero (inputs on the pixel grid): X.lo is not
                                        // loaded and the X_lo . W1_hi MMAs are not issued (two passes instead of three)
  float* logits = nullptr;        // [Z][B][C] (head == -1)
  unsigned long long* worklist = nullptr;   // num_items * kWorkPerItem slots
  int kblock_bytes = 128;
  int sm_count = 148;
};
constexpr int kWorkPerItem = 448;
bool fused_supported(int H, int C);
size_t fused_worklist_slots(int B, int Z);
int fused_forward_head(const FusedDesc& d, cudaStream_t st, std::string* err);
int fused_fixup(const FusedDesc& d, cudaStream_t st, std::string* err);

// Two-phase evaluation for the attacks (gradient of a loss of the MEAN prediction): the forward pass runs once in keep
// mode (per-sample logits + LeakyReLU masks stay in HBM, ~104 B per unit), the mean is all-reduced, and the gradient
// pass rebuilds dH from the kept data with the head and pass 2 of the fused kernel -- no second forward GEMM.
size_t keep_mask_words(int B, int Z);
int fused_keep_fixup(const FusedDesc& d, uint32_t* masks, cudaStream_t st, std::string* err);   // flips wrongly assumed mask bits
struct KeptDesc {
  int mode = MODE_TF32X3;
  int B = 0, H = 0, C = 0, Z = 0;
  const float* bank = nullptr; int64_t P = 0, wo_off = 0; int z_row0 = 0;
  int head = 0; const int32_t* labels = nullptr; const float* pbar = nullptr;
  const float* logits = nullptr;      // [Z][B][C]
  const uint32_t* masks = nullptr;    // keep_mask_words(B, Z)
  void* dh_hi = nullptr; void* dh_lo = nullptr; void* dh_bf = nullptr;   // [Z][B][H]
  const float* dh_scale = nullptr;    // F16X3
  int sm_count = 148;
};
int dh_from_kept(const KeptDesc& d, cudaStream_t st, std::string* err);

// Dynamic shared memory the kernel asks for (same for both modes).
size_t smem_bytes();

}  // namespace tc
}  // namespace rbnn
