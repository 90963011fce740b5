// The tcgen05 engine for the convolutional net (arch conv, model_nn.py:98-106), precisions RBNN_PREC_TF32X3 and
// RBNN_PREC_F16X3 (power-of-two-scaled fp16 hi/lo operands: 32 channels = one 64-byte K-block row, scales device-resident).
// Per chunk of posterior samples z and inputs b:
//   conv1 + LeakyReLU + MaxPool2d(2)          direct CUDA-core kernel (K = 25, < 1 % of the FLOPs)        conv.cu
//   conv2 = Conv2d(32, H, 5) + bias + LeakyReLU   IMPLICIT GEMM on tcgen05: A tiles are 5-D TMA boxes of the
//                                                 channels-last pooled map (one box per filter tap, no im2col
//                                                 matrix), B = W2_z as [H][(ky, kx, c)], fp32 accumulators in TMEM
//   guard-band refinement                     exact re-evaluation of the entries whose LeakyReLU sign or pooling
//                                             arg-max the tensor-core rounding could flip                     conv.cu
//   MaxPool2d(2, stride 1) + Linear(49H, C) (fused), loss head, their input gradients (fused)   CUDA cores (3 % of the FLOPs)
//   conv2 dgrad  dcol = dZ2 . W2_z  (K = H)   tcgen05 GEMM over the transposed weight copies, then col2im + conv1 backward
// This is lossGradients.py:29-40 / adversarialAttacks.py:74-78 for the conv BNN, input gradients only.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace rbnn {

// guard band of the conv2 refinement: 2^-13 of the image's largest |pre-activation|, ~22x the measured TF32x3 error
// (5.5e-6 of the output maximum at K = 800, tc_gemm_test.cu); ~11x for the difference of two entries (pooling ties)
constexpr float kConvGuardEps = 1.0f / 8192.0f;

int tc_conv_supported(const rbnn_net* n) {
  return n->arch == RBNN_ARCH_CONV && n->cc_major == 10 && (n->H % 16) == 0 && n->H <= 2048;
}

namespace {

struct ConvTcBufs {
  float *p1 = nullptr, *p1lo = nullptr, *a2 = nullptr, *logits = nullptr, *lpart = nullptr;   // p1lo: fp32 residual of P1 (forward only)
  void *p1h = nullptr, *p1l = nullptr, *dzh = nullptr, *dzl = nullptr;     // tf32 split (fp32 arrays) or fp16 hi / lo
  float *dlogits = nullptr, *dcol = nullptr, *g1 = nullptr, *partial = nullptr;
  float* call_sc = nullptr;        // F16X3: [0] s_p, [1] 1/(s_p s_w), [2] s_d, [3] 1/(s_d s_w)
  unsigned* max_bits = nullptr;    // F16X3: [0] max|P1| of the chunk, [1] max|d_pbar|
  unsigned* unit_max = nullptr;    // [ZB] max |conv2 pre-activation| per unit (float bits), from the GEMM epilogue
  uint8_t* idx1 = nullptr;
  int parts = 1;
};

// views into KeepCache::conv_buf for `units` (sample, image) units
struct KeepView {
  float *p1 = nullptr, *a2 = nullptr, *logits = nullptr;
  uint8_t* idx1 = nullptr;
};

size_t keep_bytes(const rbnn_net* n, size_t units) {
  return pad256(units * 4608 * 4) + pad256(units * 64 * n->H * 4) + pad256(units * n->C * 4) + pad256(units * 4608);
}

KeepView keep_view(const rbnn_net* n, size_t units) {
  KeepView v;
  char* p = n->keep.conv_buf;
  v.p1 = reinterpret_cast<float*>(p); p += pad256(units * 4608 * 4);
  v.a2 = reinterpret_cast<float*>(p); p += pad256(units * 64 * n->H * 4);
  v.logits = reinterpret_cast<float*>(p); p += pad256(units * n->C * 4);
  v.idx1 = reinterpret_cast<uint8_t*>(p);
  return v;
}

constexpr size_t kConvKeepBudget = (size_t)16 << 30;      // kept forward: at most 16 GB of HBM

size_t bytes_per_zb(const rbnn_net* n, bool grad) {
  const size_t H = n->H, C = n->C;
  size_t per = 4 * 4608 * 4 + 4608 + 64 * H * 4 + C * 4 + (H / 128 + 1) * C * 4 + 8;
  if (grad) per += C * 4 + 2 * 64 * H * 4 + 64 * 800 * 4 + 4608 * 4;
  return per + 64;
}

// fwd / grad: which halves of the evaluation need workspace; kv != nullptr: P1, idx1, A2 and the logits live in the
// kept-forward storage (unit offset unit0) instead of the arena
void carve(rbnn_net* n, Arena& ar, int Z, int B, bool fwd, bool grad, ConvTcBufs& c, const KeepView* kv = nullptr,
           size_t unit0 = 0) {
  const size_t H = n->H, C = n->C, ZB = (size_t)Z * B;
  c.call_sc = ar.take<float>(4);
  c.max_bits = ar.take<unsigned>(2);
  if (kv) {
    c.p1 = kv->p1 + unit0 * 4608; c.idx1 = kv->idx1 + unit0 * 4608;
    c.a2 = kv->a2 + unit0 * 64 * H; c.logits = kv->logits + unit0 * C;
  } else {
    c.p1 = ar.take<float>((size_t)ZB * 4608);
    c.idx1 = ar.take<uint8_t>((size_t)ZB * 4608);
    c.a2 = ar.take<float>((size_t)ZB * 64 * H);
    c.logits = ar.take<float>((size_t)ZB * C);
  }
  if (fwd) {
    c.p1h = ar.take<float>((size_t)ZB * 4608);      // sized for fp32; the fp16 variant uses half of it
    c.p1l = ar.take<float>((size_t)ZB * 4608);
    c.p1lo = ar.take<float>((size_t)ZB * 4608);
    c.lpart = ar.take<float>((size_t)pool2_logits_chunks(n) * ZB * C);
    c.unit_max = ar.take<unsigned>(ZB);
  }
  if (grad) {
    c.dlogits = ar.take<float>((size_t)ZB * C);
    c.dzh = ar.take<float>((size_t)ZB * 64 * H);
    c.dzl = ar.take<float>((size_t)ZB * 64 * H);
    c.dcol = ar.take<float>((size_t)ZB * 64 * 800);
    c.g1 = ar.take<float>((size_t)ZB * 4608);
    c.parts = conv1_bwd_parts(n, Z, B);
    if (c.parts > 1) c.partial = ar.take<float>((size_t)c.parts * B * 784);
  }
}

int run_tc(rbnn_net* n, tc::GemmDesc& d, int tag, cudaStream_t st) {
  d.mode = n->prec == RBNN_PREC_F16X3 ? tc::MODE_F16X3 : tc::MODE_TF32X3;
  d.sm_count = n->sm_count;
  std::string err;
  RBNN_TRY(timing_begin(n, tag, st));
  if (tc::gemm(d, st, &err)) {
    set_error("%s", err.c_str());
    return 1;
  }
  n->launches++;
  RBNN_TRY(timing_end(n, tag, st));
  return 0;
}

// gmax: device [B, C] tensor whose magnitude bounds the head's g (UPSTREAM / LOGITS_UPSTREAM heads), else nullptr;
// dh_factor: |dZ2| <= dh_factor * max|g| * max|Wout| (4 pooling windows x sum_c |dlogits_c|)
int forward_chunk(rbnn_net* n, const float* x, int B, int z0, int Z, ConvTcBufs& c, float* logits, const float* gmax,
                  float dh_factor, cudaStream_t st) {
  const int H = n->H;
  const int64_t P = n->L.P;
  const float* rows = n->bank + (int64_t)z0 * P;
  const TcMat& m = n->tc.mat[0];
  const bool f16 = n->prec == RBNN_PREC_F16X3;
  // F16X3: operand ranges of this chunk -> the call's power-of-two scales (device-resident, no read-back); max|P1| is
  // reduced by the conv1 kernel itself
  if (f16) RBNN_CUDA(cudaMemsetAsync(c.max_bits, 0, 2 * sizeof(unsigned), st));
  RBNN_TRY(conv1_pool_fwd(n, x, n->bank, z0, Z, B, c.p1, c.idx1, st, c.p1lo, f16 ? c.max_bits : nullptr));
  if (f16) {
    if (gmax) RBNN_TRY(tc_maxabs(n, gmax, (int64_t)B * n->C, c.max_bits + 1, st));
    RBNN_TRY(tc_call_scales(n, c.max_bits, gmax ? c.max_bits + 1 : nullptr, dh_factor, c.call_sc, st));
  }
  RBNN_TRY(p1_split_hwc(n, c.p1, Z * B, c.p1h, c.p1l, f16 ? c.call_sc : nullptr, st));
  tc::GemmDesc g;
  g.M = B * 64; g.N = H; g.K = f16 ? 960 : 800; g.Z = Z; g.BN = std::min(H, 256);
  g.conv_images = B;
  g.A.hi = c.p1h; g.A.lo = c.p1l; g.A.rows = g.M; g.A.ld = g.K; g.A.zstride = (int64_t)B * 4608;
  // CTA pairs (each CTA stages half of the filter tile): the single-CTA tile needs ~62 bytes per clock and SM from L2
  static const int conv_pair = getenv("RBNN_CONV_PAIR") ? atoi(getenv("RBNN_CONV_PAIR")) : 3;   // bit 0: conv2, bit 1: dgrad
  if (f16) {
    g.kblock_bytes = 128;                           // two filter taps (2 x 32 fp16 channels) per K-block
    g.pair = (conv_pair & 1) && (B % 4 == 0) && (g.BN % 32 == 0);
    g.B.hi = reinterpret_cast<const uint16_t*>(m.h_hi) + (int64_t)z0 * H * m.ld;
    g.B.lo = reinterpret_cast<const uint16_t*>(m.h_lo) + (int64_t)z0 * H * m.ld;
    g.unscale = c.call_sc + 1;
  } else {
    g.B.hi = m.hi + (int64_t)z0 * H * m.ld; g.B.lo = m.lo + (int64_t)z0 * H * m.ld;
  }
  g.B.rows = H; g.B.ld = m.ld; g.B.zstride = (int64_t)H * m.ld;
  g.epi = tc::EPI_BIAS_LEAKY;
  g.bias = rows + n->L.cb2; g.bias_zstride = P;
  g.out = c.a2; g.out_ld = H; g.out_zstride = (int64_t)B * 64 * H;
  RBNN_CUDA(cudaMemsetAsync(c.unit_max, 0, (size_t)Z * B * sizeof(unsigned), st));
  g.group_max = c.unit_max;                          // GEMM rows of an image are one group of 64
  RBNN_TRY(run_tc(n, g, 1, st));
  static const float guard_eps = getenv("RBNN_CONV_GUARD_LOG2") ? ldexpf(1.f, -atoi(getenv("RBNN_CONV_GUARD_LOG2"))) : kConvGuardEps;   // experiments
  RBNN_TRY(conv2_refine(n, c.a2, c.p1, z0, Z, B, guard_eps, st, c.p1lo, c.unit_max));
  RBNN_TRY(pool2_logits(n, c.a2, z0, Z, B, logits, c.lpart, st));
  return 0;
}

// loss head + input gradients of one chunk, starting from its logits, refined A2, P1 and arg-max indices
int backward_chunk(rbnn_net* n, int head, const int32_t* labels, const float* pbar, int B, int z0, int Z, ConvTcBufs& c,
                   float* out_sum, bool first, cudaStream_t st) {
  const int H = n->H, C = n->C;
  const TcMat& m = n->tc.mat[0];
  const bool f16 = n->prec == RBNN_PREC_F16X3;
  RBNN_TRY(head_dlogits(n, head, c.logits, labels, pbar, Z, B, C, c.dlogits, st));
  RBNN_TRY(pool2_bwd_fused(n, c.a2, c.dlogits, z0, Z, B, c.dzh, c.dzl, f16 ? c.call_sc + 2 : nullptr, st));
  // dcol[z][b * 64 + pos][c * 25 + ky * 5 + kx] = sum_h dZ2[z][b * 64 + pos][h] W2_z[h][c][ky][kx]
  tc::GemmDesc d;
  static const int dgrad_bn = getenv("RBNN_CONV_DGRAD_BN") ? atoi(getenv("RBNN_CONV_DGRAD_BN")) : 160;   // experiments
  d.M = B * 64; d.N = 800; d.K = H; d.Z = Z; d.BN = dgrad_bn;
  static const int conv_pair = getenv("RBNN_CONV_PAIR") ? atoi(getenv("RBNN_CONV_PAIR")) : 3;
  d.pair = (conv_pair & 2) ? 1 : 0;
  d.A.hi = c.dzh; d.A.lo = c.dzl; d.A.rows = d.M; d.A.ld = H; d.A.zstride = (int64_t)B * 64 * H;
  if (f16) {
    d.B.hi = reinterpret_cast<const uint16_t*>(m.th_hi) + (int64_t)z0 * 800 * H;
    d.B.lo = reinterpret_cast<const uint16_t*>(m.th_lo) + (int64_t)z0 * 800 * H;
    d.unscale = c.call_sc + 3;
  } else {
    d.B.hi = m.thi + (int64_t)z0 * 800 * H; d.B.lo = m.tlo + (int64_t)z0 * 800 * H;
  }
  d.B.rows = 800; d.B.ld = H; d.B.zstride = (int64_t)800 * H;
  d.epi = tc::EPI_NONE;
  d.out = c.dcol; d.out_ld = 800; d.out_zstride = (int64_t)B * 64 * 800;
  RBNN_TRY(run_tc(n, d, 2, st));
  RBNN_TRY(col2im_conv2(n, c.dcol, c.p1, Z * B, c.g1, st));
  RBNN_TRY(conv1_bwd_sum(n, c.g1, c.idx1, n->bank, z0, Z, B, out_sum, first ? 0 : 1, st, c.partial, c.parts));
  return 0;
}

// |dZ2| <= dh_factor * max|g| * max|Wout|: 4 pooling windows x sum_c |dlogits_c| (<= 2 max|g|, or C max|g| when the
// head hands g to the logits unchanged)
float dh_factor_of(const rbnn_net* n, int head) { return head == RBNN_HEAD_LOGITS_UPSTREAM ? 4.f * (float)n->C : 8.f; }
bool g_given(int head) { return head == RBNN_HEAD_UPSTREAM || head == RBNN_HEAD_LOGITS_UPSTREAM; }

int grad_pass(rbnn_net* n, int head, const float* x, const int32_t* labels, int B, int s0, int s1, const float* pbar,
              float* out_sum, cudaStream_t st) {
  const size_t per = bytes_per_zb(n, true) * (size_t)B + 8192 + (size_t)B * 784 * 4;   // + this sample's share of the conv1-backward partials (<= 1 slice per sample)
  const int zc = (int)std::max<size_t>(1, std::min<size_t>(n->ws_budget / per, (size_t)(s1 - s0)));
  RBNN_TRY(ws_reserve(n, per * zc));
  bool first = true;
  for (int z0 = s0; z0 < s1; z0 += zc) {
    const int Z = std::min(zc, s1 - z0);
    Arena ar(n);
    ConvTcBufs c;
    carve(n, ar, Z, B, true, true, c);
    RBNN_TRY(forward_chunk(n, x, B, z0, Z, c, c.logits, g_given(head) ? pbar : nullptr, dh_factor_of(n, head), st));
    RBNN_TRY(backward_chunk(n, head, labels, pbar, B, z0, Z, c, out_sum, first, st));
    first = false;
  }
  return 0;
}

int probs_pass(rbnn_net* n, const float* x, int B, int s0, int s1, float* out_sum, float* out_logits, cudaStream_t st) {
  const int C = n->C;
  const size_t per = bytes_per_zb(n, false) * (size_t)B + 8192;
  const int zc = (int)std::max<size_t>(1, std::min<size_t>(n->ws_budget / per, (size_t)(s1 - s0)));
  RBNN_TRY(ws_reserve(n, per * zc));
  for (int z0 = s0; z0 < s1; z0 += zc) {
    const int Z = std::min(zc, s1 - z0);
    Arena ar(n);
    ConvTcBufs c;
    carve(n, ar, Z, B, true, false, c);
    float* lg = out_logits ? out_logits : c.logits;
    RBNN_TRY(forward_chunk(n, x, B, z0, Z, c, lg, nullptr, 8.f, st));
    if (out_sum) RBNN_TRY(head_probs_accumulate(n, lg, Z, B, C, out_sum, st));
  }
  return 0;
}

// inputs per pass such that one posterior sample fits the workspace budget
int batch_rows(const rbnn_net* n, int B, bool grad) {
  const size_t rows = std::max<size_t>(1, n->ws_budget / (bytes_per_zb(n, grad) + 64));
  return (int)std::min<size_t>(rows, (size_t)B);
}

}  // namespace

int tc_conv_input_grad_sum(rbnn_net* n, int head, const float* x, const int32_t* labels, int B, int s0, int s1,
                           const float* pbar, float* out_sum, cudaStream_t st) {
  RBNN_CHECK(tc_conv_supported(n) && (n->prec == RBNN_PREC_TF32X3 || n->prec == RBNN_PREC_F16X3),
             "the tcgen05 conv engine runs TF32X3 / F16X3 on sm_100 only");
  RBNN_TRY(tc_bank_refresh(n, s0, s1, st));
  const int bc = batch_rows(n, B, true);
  for (int b0 = 0; b0 < B; b0 += bc) {
    const int nb = std::min(bc, B - b0);
    RBNN_TRY(grad_pass(n, head, x + (int64_t)b0 * n->D, labels + b0, nb, s0, s1, pbar ? pbar + (int64_t)b0 * n->C : nullptr,
                       out_sum + (int64_t)b0 * n->D, st));
  }
  return 0;
}

// Phase 1 of the two-phase attack gradient (adversarialAttacks.py:74-78 evaluates net.forward once and differentiates it):
// out_sum[B, C] += sum_s softmax(logits_s) as tc_conv_forward, and P1, its arg-max indices, the refined A2 and the
// logits of every unit stay in net->keep.conv_buf.  Falls back to the plain forward (keep.valid = 0) when the kept
// data would exceed kConvKeepBudget or the batch does not fit one pass.
int tc_conv_forward_keep(rbnn_net* n, const float* x, int B, int s0, int s1, float* out_sum, cudaStream_t st) {
  RBNN_CHECK(tc_conv_supported(n) && (n->prec == RBNN_PREC_TF32X3 || n->prec == RBNN_PREC_F16X3),
             "the tcgen05 conv engine runs TF32X3 / F16X3 on sm_100 only");
  KeepCache& k = n->keep;
  k.valid = 0;
  const size_t units = (size_t)(s1 - s0) * B;
  const size_t need = keep_bytes(n, units);
  if (need > kConvKeepBudget || batch_rows(n, B, false) < B || batch_rows(n, B, true) < B)
    return tc_conv_forward(n, x, B, s0, s1, out_sum, nullptr, st);
  RBNN_TRY(tc_bank_refresh(n, s0, s1, st));
  if (k.conv_cap < need) {
    RBNN_CUDA(cudaDeviceSynchronize());
    cudaFree(k.conv_buf);
    k.conv_buf = nullptr; k.conv_cap = 0;
    RBNN_CUDA(cudaMalloc(&k.conv_buf, need));
    n->alloc_epoch++;
    k.conv_cap = need;
  }
  const KeepView kv = keep_view(n, units);
  const size_t per = bytes_per_zb(n, false) * (size_t)B + 8192;
  const int zc = (int)std::max<size_t>(1, std::min<size_t>(n->ws_budget / per, (size_t)(s1 - s0)));
  RBNN_TRY(ws_reserve(n, per * zc));
  for (int z0 = s0; z0 < s1; z0 += zc) {
    const int Z = std::min(zc, s1 - z0);
    Arena ar(n);
    ConvTcBufs c;
    carve(n, ar, Z, B, true, false, c, &kv, (size_t)(z0 - s0) * B);
    RBNN_TRY(forward_chunk(n, x, B, z0, Z, c, c.logits, nullptr, 8.f, st));
    RBNN_TRY(head_probs_accumulate(n, c.logits, Z, B, n->C, out_sum, st));
  }
  k.valid = 1; k.B = B; k.s0 = s0; k.s1 = s1;
  return 0;
}

// Phase 2: out_sum[B, D] = sum_s dL/dx from the kept forward
int tc_conv_grad_kept(rbnn_net* n, int head, const int32_t* labels, const float* pbar, float* out_sum, cudaStream_t st) {
  KeepCache& k = n->keep;
  RBNN_CHECK(k.valid && k.conv_buf, "no kept forward: call rbnn_forward_probs_sum_keep first (and check rbnn_keep_valid)");
  RBNN_CHECK(head != RBNN_HEAD_LOGITS_CE, "LOGITS_CE has no kept route");
  const int B = k.B, s0 = k.s0, s1 = k.s1;
  const bool f16 = n->prec == RBNN_PREC_F16X3;
  const KeepView kv = keep_view(n, (size_t)(s1 - s0) * B);
  const size_t per = bytes_per_zb(n, true) * (size_t)B + 8192 + (size_t)B * 784 * 4;
  const int zc = (int)std::max<size_t>(1, std::min<size_t>(n->ws_budget / per, (size_t)(s1 - s0)));
  RBNN_TRY(ws_reserve(n, per * zc));
  bool first = true;
  for (int z0 = s0; z0 < s1; z0 += zc) {
    const int Z = std::min(zc, s1 - z0);
    Arena ar(n);
    ConvTcBufs c;
    carve(n, ar, Z, B, false, true, c, &kv, (size_t)(z0 - s0) * B);
    if (f16) {          // the dZ2 range of THIS head ([0], [1] of the scalars belong to the forward and are not used here)
      RBNN_CUDA(cudaMemsetAsync(c.max_bits, 0, 2 * sizeof(unsigned), st));
      if (g_given(head)) RBNN_TRY(tc_maxabs(n, pbar, (int64_t)B * n->C, c.max_bits + 1, st));
      RBNN_TRY(tc_call_scales(n, c.max_bits, g_given(head) ? c.max_bits + 1 : nullptr, dh_factor_of(n, head), c.call_sc, st));
    }
    RBNN_TRY(backward_chunk(n, head, labels, pbar, B, z0, Z, c, out_sum, first, st));
    first = false;
  }
  return 0;
}

int tc_conv_forward(rbnn_net* n, const float* x, int B, int s0, int s1, float* out_sum, float* out_logits,
                    cudaStream_t st) {
  RBNN_CHECK(tc_conv_supported(n) && (n->prec == RBNN_PREC_TF32X3 || n->prec == RBNN_PREC_F16X3),
             "the tcgen05 conv engine runs TF32X3 / F16X3 on sm_100 only");
  RBNN_TRY(tc_bank_refresh(n, s0, s1, st));
  const int bc = batch_rows(n, B, false);
  for (int b0 = 0; b0 < B; b0 += bc) {
    const int nb = std::min(bc, B - b0);
    RBNN_TRY(probs_pass(n, x + (int64_t)b0 * n->D, nb, s0, s1, out_sum ? out_sum + (int64_t)b0 * n->C : nullptr,
                        out_logits ? out_logits + (int64_t)b0 * n->C : nullptr, st));
  }
  return 0;
}

}  // namespace rbnn
