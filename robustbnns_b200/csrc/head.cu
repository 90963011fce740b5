// Class-dimension head: softmax of the logits, posterior-mean accumulation, and d(loss)/d(logits)
// for the three loss definitions of the reference.  C (classes) is 2..32; one thread owns one
// (sample, input) row and keeps the C values in registers.
//
//   BNN.forward                  p_s = softmax(z_s); out = mean_s p_s         model_bnn.py:134,254,257
//   loss_gradient                L_s = CE(p_s, y) = -log softmax(p_s)[y]      lossGradients.py:33-34
//   fgsm/pgd                     L   = CE(mean_s p_s, y)                      adversarialAttacks.py:74-76
//   avg_posterior                L   = CE(z, y) on logits                     model_bnn.py:206-216
//
// With q = softmax(p) (the "double softmax"), g = q - e_y and p = softmax(z):
//   dL/dz = p * (g - <p, g>)          (softmax Jacobian applied to g)
// evaluated in exactly this form -- the one torch's softmax backward uses (grad - sum(grad*out)) * out -- so that
// the fp32 rounding behaviour of saturated rows follows the reference's (a cancellation-free rewrite is more
// accurate but moves the sign of near-zero gradients away from the reference's golden vectors).
#include "common.cuh"

namespace rbnn {

constexpr int kMaxC = 32;

template <int C_MAX>
__device__ __forceinline__ void softmax_inplace(float (&v)[C_MAX], int C) {
  float mx = v[0];
#pragma unroll
  for (int c = 1; c < C_MAX; ++c)
    if (c < C) mx = fmaxf(mx, v[c]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < C_MAX; ++c)
    if (c < C) { v[c] = expf(v[c] - mx); sum += v[c]; }
  const float inv = 1.f / sum;
#pragma unroll
  for (int c = 0; c < C_MAX; ++c)
    if (c < C) v[c] *= inv;
}

// out_sum[b][c] += sum_z softmax(logits[z][b][:])[c].  One warp per input: lane l sums samples l, l+32, ... in order,
// then a fixed butterfly over the lanes => deterministic, and B x 32 threads instead of B (the attack loop calls this
// with B ~ 1000: one thread per input left most of the GPU idle).
// RAW: accumulate the logits themselves (Ensemble_NN.forward averages logits, model_ensemble.py:62-66).
template <int C_MAX, bool RAW>
__global__ void probs_accumulate_kernel(const float* __restrict__ logits, int Z, int B, int C,
                                        float* __restrict__ out_sum) {
  const int lane = threadIdx.x & 31;
  const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (b >= B) return;
  float acc[C_MAX];
#pragma unroll
  for (int c = 0; c < C_MAX; ++c) acc[c] = 0.f;
  for (int z = lane; z < Z; z += 32) {
    float v[C_MAX];
    const float* row = logits + ((int64_t)z * B + b) * C;
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) v[c] = (c < C) ? __ldg(row + c) : 0.f;
    if (!RAW) softmax_inplace<C_MAX>(v, C);
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) acc[c] += v[c];
  }
#pragma unroll
  for (int c = 0; c < C_MAX; ++c)
    if (c < C) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
  if (lane < C) {
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c == lane) v = acc[c];
    out_sum[(int64_t)b * C + lane] += v;
  }
}

template <int C_MAX>
__global__ void dlogits_kernel(int head, const float* __restrict__ logits, const int32_t* __restrict__ labels,
                               const float* __restrict__ pbar, int Z, int B, int C, float* __restrict__ dlogits) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (int64_t)Z * B) return;
  const int b = (int)(r % B);
  if (head == RBNN_HEAD_LOGITS_UPSTREAM) {       // loss of the MEAN LOGITS: every sample receives d_pbar as is
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) dlogits[r * C + c] = __ldg(pbar + (int64_t)b * C + c);
    return;
  }
  const int y = labels[b];
  float p[C_MAX], g[C_MAX];
#pragma unroll
  for (int c = 0; c < C_MAX; ++c) p[c] = (c < C) ? __ldg(logits + r * C + c) : 0.f;
  softmax_inplace<C_MAX>(p, C);
  if (head == RBNN_HEAD_LOGITS_CE) {
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) dlogits[r * C + c] = p[c] - (c == y ? 1.f : 0.f);
    return;
  }
  if (head == RBNN_HEAD_MEAN_OF_GRADS) {
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) g[c] = p[c];
  } else {
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) g[c] = (c < C) ? __ldg(pbar + (int64_t)b * C + c) : 0.f;
  }
  if (head != RBNN_HEAD_UPSTREAM) {
    softmax_inplace<C_MAX>(g, C);
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) g[c] -= (c == y ? 1.f : 0.f);
  }
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < C_MAX; ++c)
    if (c < C) dot = fmaf(p[c], g[c], dot);
#pragma unroll
  for (int c = 0; c < C_MAX; ++c)
    if (c < C) dlogits[r * C + c] = p[c] * (g[c] - dot);
}

int head_probs_accumulate(rbnn_net* net, const float* logits, int Z, int B, int C, float* out_sum,
                          cudaStream_t st) {
  RBNN_CHECK(C >= 1 && C <= kMaxC, "head: n_classes %d not in [1,%d]", C, kMaxC);
  const int thr = 128;                      // 4 inputs per block
  const unsigned blocks = (unsigned)(((int64_t)B * 32 + thr - 1) / thr);
  if (net->sum_logits) {
    if (C <= 16) probs_accumulate_kernel<16, true><<<blocks, thr, 0, st>>>(logits, Z, B, C, out_sum);
    else probs_accumulate_kernel<32, true><<<blocks, thr, 0, st>>>(logits, Z, B, C, out_sum);
  } else {
    if (C <= 16) probs_accumulate_kernel<16, false><<<blocks, thr, 0, st>>>(logits, Z, B, C, out_sum);
    else probs_accumulate_kernel<32, false><<<blocks, thr, 0, st>>>(logits, Z, B, C, out_sum);
  }
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

int head_dlogits(rbnn_net* net, int head, const float* logits, const int32_t* labels, const float* pbar, int Z,
                 int B, int C, float* dlogits, cudaStream_t st) {
  RBNN_CHECK(C >= 1 && C <= kMaxC, "head: n_classes %d not in [1,%d]", C, kMaxC);
  RBNN_CHECK((head != RBNN_HEAD_GRAD_OF_MEAN && head != RBNN_HEAD_UPSTREAM && head != RBNN_HEAD_LOGITS_UPSTREAM) ||
                 pbar != nullptr, "head: needs d_pbar");
  const int thr = 128;
  const int64_t rows = (int64_t)Z * B;
  const unsigned blocks = (unsigned)((rows + thr - 1) / thr);
  if (C <= 16)
    dlogits_kernel<16><<<blocks, thr, 0, st>>>(head, logits, labels, pbar, Z, B, C, dlogits);
  else
    dlogits_kernel<32><<<blocks, thr, 0, st>>>(head, logits, labels, pbar, Z, B, C, dlogits);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rbnn
