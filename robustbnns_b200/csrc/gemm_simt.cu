// Generic batched fp32 GEMM on the CUDA cores (FFMA): the reference-class-rounding engine
// (RBNN_PREC_FP32) and the fallback for every shape the tcgen05 path does not cover
// (conv im2col GEMMs, the narrow class-dimension products, half-moons D=2).
//
// Replaces, per (posterior sample z): nn.Linear forward (model_nn.py:80,82,89,91,106) as
//   NT:  C[z][m][n] = epi( sum_k A[z][m][k] * B[z][n][k] + bias[z][n] )
// and autograd's input-gradient of nn.Linear / the im2col'd nn.Conv2d as
//   NN:  C[z][m][n] = epi( sum_k A[z][m][k] * B[z][k][n] )
// with the LeakyReLU forward / derivative fused into the epilogue, and optionally the sum over
// posterior samples (lossGradients.py:40, model_bnn.py:257) folded in as a concatenated-K loop.
#include "common.cuh"

namespace rbnn {

// hidden-layer activation and its derivative expressed through the STORED activation value (model_nn.py:66-75)
__device__ __forceinline__ float act_fwd(int act, float v) {
  switch (act) {
    case RBNN_ACT_RELU: return v > 0.f ? v : 0.f;
    case RBNN_ACT_SIGM: return 1.f / (1.f + expf(-v));
    case RBNN_ACT_TANH: return tanhf(v);
    default: return v > 0.f ? v : v * kLeakySlope;
  }
}
__device__ __forceinline__ float act_grad(int act, float a) {
  switch (act) {
    case RBNN_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case RBNN_ACT_SIGM: return a * (1.f - a);
    case RBNN_ACT_TANH: return 1.f - a * a;
    default: return a > 0.f ? 1.f : kLeakySlope;
  }
}


template <int BM, int BN, int BK, int TM, int TN, bool BKN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(GemmArgs a, int z_per_block, float* __restrict__ partial) {
  constexpr int NTHR = (BM / TM) * (BN / TN);
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int zb = blockIdx.z * z_per_block;
  const int ze = min(zb + z_per_block, a.Z);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int z = zb; z < ze; ++z) {
    const float* __restrict__ A = a.A + (int64_t)z * a.sAz;
    const float* __restrict__ B = a.B + (int64_t)z * a.sBz;
    for (int k0 = 0; k0 < a.K; k0 += BK) {
      for (int i = tid; i < BM * BK; i += NTHR) {
        const int m = i / BK, k = i % BK;
        float v = 0.f;
        if (m0 + m < a.M && k0 + k < a.K) v = __ldg(A + (int64_t)(m0 + m) * a.lda + (k0 + k));
        As[k][m] = v;
      }
      if (!BKN) {
        for (int i = tid; i < BN * BK; i += NTHR) {
          const int n = i / BK, k = i % BK;
          float v = 0.f;
          if (n0 + n < a.N && k0 + k < a.K) v = __ldg(B + (int64_t)(n0 + n) * a.ldb + (k0 + k));
          Bs[k][n] = v;
        }
      } else {
        for (int i = tid; i < BN * BK; i += NTHR) {
          const int k = i / BN, n = i % BN;
          float v = 0.f;
          if (n0 + n < a.N && k0 + k < a.K) v = __ldg(B + (int64_t)(k0 + k) * a.ldb + (n0 + n));
          Bs[k][n] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float ar[TM], br[TN];
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          const float4 v = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
          ar[i] = v.x; ar[i + 1] = v.y; ar[i + 2] = v.z; ar[i + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + j]);
          br[j] = v.x; br[j + 1] = v.y; br[j + 2] = v.z; br[j + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
      __syncthreads();
    }
    if (!a.reduce_z) {
      float* __restrict__ C = a.C + (int64_t)z * a.sCz;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int n = n0 + tx * TN + j;
          if (n >= a.N) continue;
          float v = acc[i][j];
          if (a.epi == EPI_BIAS || a.epi == EPI_BIAS_LEAKY) v += __ldg(a.bias + (int64_t)z * a.sbz + n);
          if (a.epi == EPI_BIAS_LEAKY) v = act_fwd(a.act, v);
          if (a.epi == EPI_MASK) v *= act_grad(a.act, __ldg(a.mask + (int64_t)z * a.sMz + (int64_t)m * a.ldm + n));
          C[(int64_t)m * a.ldc + n] = v;
          acc[i][j] = 0.f;
        }
      }
    }
  }
  if (a.reduce_z) {
    float* __restrict__ C = (gridDim.z == 1) ? a.C : partial + (int64_t)blockIdx.z * a.M * a.ldc;
    const bool add = (gridDim.z == 1) && a.accumulate;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m >= a.M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n >= a.N) continue;
        const int64_t o = (int64_t)m * a.ldc + n;
        C[o] = add ? C[o] + acc[i][j] : acc[i][j];
      }
    }
  }
}

// out[i] (+)= sum_p partial[p][i], fixed order -> deterministic
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nparts, int64_t n,
                                       float* __restrict__ out, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = accumulate ? out[i] : 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * n + i];
  out[i] = s;
}

template <bool BKN>
static int launch(rbnn_net* net, const GemmArgs& a, cudaStream_t st) {
  constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, 1);
  int zpb = 1;
  float* partial = nullptr;
  if (a.reduce_z) {
    // split the sample loop over blockIdx.z until the grid covers the SMs ~2x
    const int tiles = grid.x * grid.y;
    int zsplit = 1;
    while (tiles * zsplit < 2 * net->sm_count && zsplit * 2 <= a.Z && zsplit < 64) zsplit *= 2;
    if (zsplit > 1) {
      const size_t need = (size_t)zsplit * a.M * a.ldc * sizeof(float);
      // partials live at the END of the workspace arena (callers bump-allocate from the front)
      if (net->ws_bytes < need) { set_error("gemm_simt: workspace too small for split-z partials"); return 1; }
      partial = reinterpret_cast<float*>(net->ws + ((net->ws_bytes - need) & ~(size_t)255));
      if (a.ldc != a.N) zsplit = 1, partial = nullptr;  // partial layout assumes dense C
    }
    zpb = (a.Z + zsplit - 1) / zsplit;
    grid.z = (a.Z + zpb - 1) / zpb;
    if (grid.z == 1) partial = nullptr;
  } else {
    grid.z = a.Z;
  }
  if (a.tag) RBNN_TRY(timing_begin(net, a.tag, st));
  gemm_simt_kernel<BM, BN, BK, TM, TN, BKN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(a, zpb, partial);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  if (a.tag) RBNN_TRY(timing_end(net, a.tag, st));
  if (a.reduce_z && grid.z > 1) {
    const int64_t n = (int64_t)a.M * a.N;
    reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, grid.z, n, a.C, a.accumulate);
    net->launches++;
    RBNN_CUDA(cudaGetLastError());
  }
  return 0;
}

int gemm_simt(rbnn_net* net, const GemmArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0 || a.Z <= 0) return 0;
  return a.b_kn ? launch<true>(net, a, st) : launch<false>(net, a, st);
}

}  // namespace rbnn
