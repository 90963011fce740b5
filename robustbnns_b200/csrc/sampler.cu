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
#include "philox.cuh"

namespace rbnn {

__global__ void softplus_kernel(const float* __restrict__ rho, float* __restrict__ sigma, int64_t P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float x = rho[i];
  sigma[i] = x > 20.f ? x : log1pf(expf(x));
}

// grid: (ceil(P/4/256), count)
__global__ void __launch_bounds__(256)
sample_diag_kernel(const float* __restrict__ loc, const float* __restrict__ sigma, float* __restrict__ bank,
                   int64_t P, uint32_t k0, uint32_t k1, int64_t sample_index0, int64_t stride, int s0, int vec,
                   int64_t q0, const int64_t* __restrict__ index_offset) {
  const int64_t q = q0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // q0: first quad (the tail of a row only)
  const int64_t i0 = q * 4;
  if (i0 >= P) return;
  const int64_t g = sample_index0 + (index_offset ? *index_offset : 0) + (int64_t)blockIdx.y * stride;
  float z[4];
  philox_normals4((uint32_t)q, (uint32_t)g, k0, k1, z);
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

int sample_sigma(rbnn_net* net, const float* d_rho, cudaStream_t st) {
  const int64_t P = net->L.P;
  if (!net->sigma) {
    RBNN_CUDA(cudaMalloc(&net->sigma, P * sizeof(float)));
    net->alloc_epoch++;
  }
  softplus_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(d_rho, net->sigma, P);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// elements [elem0, P) of the rows (elem0 % 4 == 0; 0 = whole rows); net->sigma must hold softplus(rho) (sample_sigma)
int sample_diag_from(rbnn_net* net, const float* d_loc, uint64_t seed, int64_t sample_index0, int64_t stride, int s0,
                     int count, int64_t elem0, cudaStream_t st, const int64_t* d_index_offset) {
  const int64_t P = net->L.P;
  const int64_t nq = (P + 3) / 4 - elem0 / 4;
  if (nq <= 0) return 0;
  dim3 grid((unsigned)((nq + 255) / 256), (unsigned)count);
  sample_diag_kernel<<<grid, 256, 0, st>>>(d_loc, net->sigma, net->bank, P, (uint32_t)(seed & 0xFFFFFFFFu),
                                           (uint32_t)(seed >> 32), sample_index0, stride, s0,
                                           ((P & 1) == 0 && (reinterpret_cast<uintptr_t>(d_loc) & 15) == 0) ? 1 : 0, elem0 / 4,
                                           d_index_offset);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

int sample_diag(rbnn_net* net, const float* d_loc, const float* d_rho, uint64_t seed, int64_t sample_index0,
                int64_t stride, int s0, int count, cudaStream_t st, const int64_t* d_index_offset) {
  RBNN_TRY(sample_sigma(net, d_rho, st));
  return sample_diag_from(net, d_loc, seed, sample_index0, stride, s0, count, 0, st, d_index_offset);
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
