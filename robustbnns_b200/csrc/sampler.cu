// K-sample: materialise diagonal-Gaussian posterior samples straight into the HBM bank.
//
// Replaces BNN.guide's per-(image, sample) Pyro draw (model_bnn.py:121-130):
//     w = loc + softplus(scale) * eps,   eps ~ N(0, 1) independently per scalar
// with one counter-based draw per posterior sample: element i of GLOBAL sample index g uses
// Philox4x32-10 counter (i/4, g, 0x52424E4E, 0) under key = the 64-bit seed, so a bank row does
// not depend on which GPU, chunk or launch produced it (sample sharding).
// torch.nn.Softplus() defaults beta=1, threshold=20 (model_bnn.py:18).
// The numpy restatement used by the tests is oracle/oracle.py::philox_standard_normals.
#include "common.cuh"

namespace rbnn {

__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                              uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ float u01(uint32_t x) {  // ((x>>9)+0.5) * 2^-23, exact in fp32, in (0,1)
  return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-07f;
}

// Box-Muller on the special-function units: the angle is taken in (-pi, pi) (same distribution as (0, 2 pi)), where
// __sinf / __cosf err by <= 2^-21.4 absolutely; __logf errs by ~2 ulp on (0, 1).  |z| <= 5.8, so a normal moves by
// < 3e-6 against libm -- the restatement test (oracle.philox_standard_normals) allows 2e-5 sigma.
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& z0, float& z1) {
  const float r = sqrtf(-2.f * __logf(u01(xa)));
  const float t = 6.283185307179586f * u01(xb);
  const float a = t - 3.14159265358979f;        // cos t = -cos a, sin t = -sin a
  z0 = -r * __cosf(a);
  z1 = -r * __sinf(a);
}

__global__ void softplus_kernel(const float* __restrict__ rho, float* __restrict__ sigma, int64_t P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float x = rho[i];
  sigma[i] = x > 20.f ? x : log1pf(expf(x));
}

// grid: (ceil(P/4/256), count)
__global__ void __launch_bounds__(256)
sample_diag_kernel(const float* __restrict__ loc, const float* __restrict__ sigma, float* __restrict__ bank,
                   int64_t P, uint32_t k0, uint32_t k1, int64_t sample_index0, int64_t stride, int s0, int vec) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = q * 4;
  if (i0 >= P) return;
  const int64_t g = sample_index0 + (int64_t)blockIdx.y * stride;
  uint32_t c0 = (uint32_t)q, c1 = (uint32_t)g, c2 = 0x52424E4Eu, c3 = 0u;
  philox4x32_10(c0, c1, c2, c3, k0, k1);
  float z[4];
  box_muller(c0, c1, z[0], z[1]);
  box_muller(c2, c3, z[2], z[3]);
  float* __restrict__ row = bank + (int64_t)(s0 + blockIdx.y) * P;
  if (vec && i0 + 3 < P) {          // P even and loc 16-byte aligned: 16-byte loads, 8-byte stores (rows are 8-byte aligned)
    const float4 sg = __ldg(reinterpret_cast<const float4*>(sigma + i0));
    const float4 lc = __ldg(reinterpret_cast<const float4*>(loc + i0));
    float2* out = reinterpret_cast<float2*>(row + i0);
    out[0] = make_float2(fmaf(sg.x, z[0], lc.x), fmaf(sg.y, z[1], lc.y));
    out[1] = make_float2(fmaf(sg.z, z[2], lc.z), fmaf(sg.w, z[3], lc.w));
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j < P) row[i0 + j] = fmaf(__ldg(sigma + i0 + j), z[j], __ldg(loc + i0 + j));
}

int sample_diag(rbnn_net* net, const float* d_loc, const float* d_rho, uint64_t seed, int64_t sample_index0,
                int64_t stride, int s0, int count, cudaStream_t st) {
  const int64_t P = net->L.P;
  if (!net->sigma) RBNN_CUDA(cudaMalloc(&net->sigma, P * sizeof(float)));
  softplus_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(d_rho, net->sigma, P);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  const int64_t nq = (P + 3) / 4;
  dim3 grid((unsigned)((nq + 255) / 256), (unsigned)count);
  sample_diag_kernel<<<grid, 256, 0, st>>>(d_loc, net->sigma, net->bank, P, (uint32_t)(seed & 0xFFFFFFFFu),
                                           (uint32_t)(seed >> 32), sample_index0, stride, s0,
                                           ((P & 1) == 0 && (reinterpret_cast<uintptr_t>(d_loc) & 15) == 0) ? 1 : 0);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// conv: the output Linear consumes the pooled activations in (pos, h) order (HWC), the reference
// flattens (h, pos) (CHW, model_nn.py:105-106) => keep a permuted copy of model.7.weight per row, classes innermost
// and zero-padded to a multiple of 4 (conv_class_pitch): the fused pooling kernels fetch an entry with 16-byte loads.
__global__ void permute_wout_kernel(const float* __restrict__ bank, int64_t P, int64_t wo, int C, int CP, int H,
                                    float* __restrict__ woutp, int s0) {
  const int F = 49 * H;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)CP * F) return;
  const int s = s0 + blockIdx.y;
  const int k = (int)(i / CP), c = (int)(i % CP);
  const int p = k / H, h = k % H;
  woutp[((int64_t)s * F + k) * CP + c] = c < C ? __ldg(bank + (int64_t)s * P + wo + (int64_t)c * F + h * 49 + p) : 0.f;
}

int conv_permute_wout(rbnn_net* net, int s0, int count, cudaStream_t st) {
  if (net->arch != RBNN_ARCH_CONV || count <= 0) return 0;
  const int CP = conv_class_pitch(net->C);
  const int64_t n = (int64_t)CP * 49 * net->H;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)count);
  permute_wout_kernel<<<grid, 256, 0, st>>>(net->bank, net->L.P, net->L.wo, net->C, CP, net->H, net->woutp, s0);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rbnn
