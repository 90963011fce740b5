// K-attack / K-rob: the elementwise attack updates and the evaluation reductions.
//   fgsm  : clamp(x + eps*sign(g), 0, 1)                              adversarialAttacks.py:81-82
//   pgd   : clamp(x0 + clamp(x + alpha*sign(g) - x0, -eps, eps), 0, 1) adversarialAttacks.py:103-105
//   alpha : 2 / image.max() per image                                 adversarialAttacks.py:89
//   rob   : 1 - max_c |softmax(o0) - softmax(o1)|                      adversarialAttacks.py:36-46,60
//   count : sum(argmax(out) == label)                                 adversarialAttacks.py:179,186
// All HBM-bound streaming kernels: 128-bit accesses where the layout allows, one pass over the data.
#include "common.cuh"

namespace rbnn {

__device__ __forceinline__ float sgn(float g) { return (float)(g > 0.f) - (float)(g < 0.f); }  // sign(0) = 0
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void fgsm_step_kernel(const float* __restrict__ x, const float* __restrict__ g, float eps,
                                 float* __restrict__ out, int64_t n) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const float4 xv = *reinterpret_cast<const float4*>(x + i4);
    const float4 gv = *reinterpret_cast<const float4*>(g + i4);
    float4 o;
    o.x = clampf(xv.x + eps * sgn(gv.x), 0.f, 1.f);
    o.y = clampf(xv.y + eps * sgn(gv.y), 0.f, 1.f);
    o.z = clampf(xv.z + eps * sgn(gv.z), 0.f, 1.f);
    o.w = clampf(xv.w + eps * sgn(gv.w), 0.f, 1.f);
    *reinterpret_cast<float4*>(out + i4) = o;
  } else {
    for (int64_t i = i4; i < n; ++i) out[i] = clampf(x[i] + eps * sgn(g[i]), 0.f, 1.f);
  }
}

__global__ void pgd_step_kernel(const float* __restrict__ x, const float* __restrict__ x0,
                                const float* __restrict__ g, const float* __restrict__ alpha, float eps,
                                float* __restrict__ out, int B, int D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * D) return;
  const float a = __ldg(alpha + i / D);
  const float orig = x0[i];
  const float pert = x[i] + a * sgn(g[i]);
  const float eta = clampf(pert - orig, -eps, eps);
  out[i] = clampf(orig + eta, 0.f, 1.f);
}

__global__ void pgd_alpha_kernel(const float* __restrict__ x, float* __restrict__ alpha, int D) {
  __shared__ float red[32];
  const float* row = x + (int64_t)blockIdx.x * D;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < D; i += blockDim.x) m = fmaxf(m, row[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) alpha[blockIdx.x] = 2.f / m;
  }
}

__global__ void minmax_init_kernel(float* mm) {
  mm[0] = INFINITY;
  mm[1] = 0.f;
}

template <int C_MAX>
__global__ void softmax_robustness_kernel(const float* __restrict__ o0, const float* __restrict__ o1, int N, int C,
                                          float* __restrict__ rob, float* __restrict__ minmax) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float d = 0.f;
  if (n < N) {
    float a[C_MAX], b[C_MAX];
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) {
        a[c] = __ldg(o0 + (int64_t)n * C + c);
        b[c] = __ldg(o1 + (int64_t)n * C + c);
        ma = fmaxf(ma, a[c]);
        mb = fmaxf(mb, b[c]);
      }
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) {
        a[c] = expf(a[c] - ma); sa += a[c];
        b[c] = expf(b[c] - mb); sb += b[c];
      }
    const float ia = 1.f / sa, ib = 1.f / sb;
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) d = fmaxf(d, fabsf(a[c] * ia - b[c] * ib));
    rob[n] = 1.f - d;
  }
  // block min/max of the differences (non-negative floats order like their bit patterns)
  float lo = n < N ? d : INFINITY, hi = n < N ? d : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(reinterpret_cast<int*>(minmax), __float_as_int(lo));
    atomicMax(reinterpret_cast<int*>(minmax) + 1, __float_as_int(hi));
  }
}

template <int C_MAX>
__global__ void count_correct_kernel(const float* __restrict__ out, const int32_t* __restrict__ labels, int N, int C,
                                     unsigned long long* __restrict__ count) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (n < N) {
    float best = __ldg(out + (int64_t)n * C);
    int bi = 0;
#pragma unroll
    for (int c = 1; c < C_MAX; ++c)
      if (c < C) {
        const float v = __ldg(out + (int64_t)n * C + c);
        if (v > best) { best = v; bi = c; }
      }
    ok = (bi == labels[n]);
  }
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

__global__ void scale_kernel(float* __restrict__ p, float s, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] *= s;
}

__global__ void add_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

int scale_inplace(rbnn_net* net, float* p, float scale, int64_t n, cudaStream_t st) {
  scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, scale, n);
  if (net) net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

int add_inplace(rbnn_net* net, float* dst, const float* src, int64_t n, cudaStream_t st) {
  add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, src, n);
  if (net) net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rbnn

using namespace rbnn;

// The stateless entry points have no handle: they launch on the device that OWNS the buffers (the caller's current
// device may be another one), restoring the caller's device afterwards.
namespace {
struct PtrDeviceGuard {
  int prev = -1;
  explicit PtrDeviceGuard(const void* p) {
    cudaPointerAttributes a;
    if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) {
      int cur = -1;
      cudaGetDevice(&cur);
      if (cur != a.device) { prev = cur; cudaSetDevice(a.device); }
    } else {
      cudaGetLastError();
    }
  }
  ~PtrDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

extern "C" int rbnn_fgsm_step(const float* d_x, const float* d_grad, float eps, float* d_out, int64_t n,
                              void* stream) {
  if (n <= 0) return 0;
  RBNN_CHECK(((uintptr_t)d_x | (uintptr_t)d_grad | (uintptr_t)d_out) % 16 == 0, "fgsm_step: buffers must be 16-byte aligned");
  const int64_t n4 = (n + 3) / 4;
  PtrDeviceGuard dg(d_x);
  fgsm_step_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_grad, eps, d_out, n);
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rbnn_pgd_step(const float* d_x, const float* d_x0, const float* d_grad, const float* d_alpha,
                             float eps, float* d_out, int B, int D, void* stream) {
  const int64_t n = (int64_t)B * D;
  if (n <= 0) return 0;
  PtrDeviceGuard dg(d_x);
  pgd_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_x0, d_grad, d_alpha, eps,
                                                                                d_out, B, D);
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rbnn_pgd_alpha(const float* d_x, float* d_alpha, int B, int D, void* stream) {
  if (B <= 0) return 0;
  PtrDeviceGuard dg(d_x);
  pgd_alpha_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(d_x, d_alpha, D);
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rbnn_softmax_robustness(const float* d_o0, const float* d_o1, int N, int C, float* d_rob,
                                       float* d_minmax, void* stream) {
  RBNN_CHECK(C >= 1 && C <= 32, "softmax_robustness: n_classes %d not in [1,32]", C);
  cudaStream_t st = (cudaStream_t)stream;
  PtrDeviceGuard dg(d_minmax);
  minmax_init_kernel<<<1, 1, 0, st>>>(d_minmax);
  if (N > 0) {
    if (C <= 16)
      softmax_robustness_kernel<16><<<(N + 127) / 128, 128, 0, st>>>(d_o0, d_o1, N, C, d_rob, d_minmax);
    else
      softmax_robustness_kernel<32><<<(N + 127) / 128, 128, 0, st>>>(d_o0, d_o1, N, C, d_rob, d_minmax);
  }
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rbnn_count_correct(const float* d_out, const int32_t* d_labels, int N, int C, int64_t* d_count,
                                  void* stream) {
  RBNN_CHECK(C >= 1 && C <= 32, "count_correct: n_classes %d not in [1,32]", C);
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  PtrDeviceGuard dg(d_out);
  if (C <= 16)
    count_correct_kernel<16><<<(N + 127) / 128, 128, 0, st>>>(d_out, d_labels, N, C, (unsigned long long*)d_count);
  else
    count_correct_kernel<32><<<(N + 127) / 128, 128, 0, st>>>(d_out, d_labels, N, C, (unsigned long long*)d_count);
  RBNN_CUDA(cudaGetLastError());
  return 0;
}
