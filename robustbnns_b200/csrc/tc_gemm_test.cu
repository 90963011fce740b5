// Stand-alone bring-up / timing harness for the tcgen05 GEMM (tc_gemm.cu).  Not part of librbnn.so.
//   build/tc_gemm_test            correctness on small ragged shapes vs a double-precision host reference
//   build/tc_gemm_test bench      + timing of the headline shapes (10 000 x 784 x 512 per posterior sample)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tc_gemm.cuh"

using namespace rbnn::tc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__host__ __device__ inline uint32_t hash32(uint64_t i, uint32_t seed) {
  uint64_t x = i * 0x9E3779B97F4A7C15ull + seed;
  x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29; x *= 0x94D049BB133111EBull; x ^= x >> 32;
  return (uint32_t)x;
}
__host__ __device__ inline float val(uint64_t i, uint32_t seed, float scale) {
  return ((float)(hash32(i, seed) & 0xFFFFFF) * (1.0f / 16777216.0f) - 0.5f) * scale;
}
__host__ __device__ inline float tf32_rn(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & ~0x1FFFu;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

__global__ void fill_kernel(float* hi, float* lo, __nv_bfloat16* bf, int64_t n, uint32_t seed, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = val(i, seed, scale);
    uint32_t u = __float_as_uint(v);
    u = (u + 0x1000u) & ~0x1FFFu;
    const float h = __uint_as_float(u);
    hi[i] = h;
    lo[i] = v - h;
    if (bf) bf[i] = __float2bfloat16(v);
  }
}

// F16X3 operands: fp16 hi/lo of s * v (s a power of two)
__global__ void fill_f16_kernel(__half* hi, __half* lo, int64_t n, uint32_t seed, float scale, float s) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = val(i, seed, scale) * s;
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}

struct Dev {
  float *hi = nullptr, *lo = nullptr;
  __nv_bfloat16* bf = nullptr;
  __half *h16 = nullptr, *l16 = nullptr;
  void init_f16(float s) {
    CK(cudaMalloc(&h16, n * 2));
    CK(cudaMalloc(&l16, n * 2));
    fill_f16_kernel<<<1184, 256>>>(h16, l16, n, seed, scale, s);
    CK(cudaGetLastError());
  }
  int64_t n = 0;
  uint32_t seed;
  float scale;
  void init(int64_t n_, uint32_t seed_, float scale_, bool want_bf) {
    n = n_; seed = seed_; scale = scale_;
    CK(cudaMalloc(&hi, n * 4));
    CK(cudaMalloc(&lo, n * 4));
    if (want_bf) CK(cudaMalloc(&bf, n * 2));
    fill_kernel<<<1184, 256>>>(hi, lo, bf, n, seed, scale);
    CK(cudaGetLastError());
  }
  float at(int64_t i) const { return val(i, seed, scale); }
  void free_() { cudaFree(hi); cudaFree(lo); cudaFree(bf); cudaFree(h16); cudaFree(l16); }
};

struct Case {
  const char* name;
  int mode, M, N, K, Z, BN, reduce, slots, a_per_z, epi, split_out;
};
static int g_kblock = 64, g_pair = 1, g_skip = 0, g_relay = 1, g_spin = 0, g_ldpad = 0, g_noout = 0;

static float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }

// returns max |err| / max |ref| over the checked entries
static double run_case(const Case& c, bool full_check, int timing_iters, double* ms_out) {
  const bool bf = c.mode == MODE_BF16, f16 = c.mode == MODE_F16X3;
  Dev A, B, bias, act;
  const int64_t a_z = c.a_per_z ? c.Z : 1;
  const int per_line = c.mode == MODE_TF32X3 ? 32 : 64;          // TC_LDPAD=1: operand rows start on 128-byte lines
  const int64_t ldk = g_ldpad ? (c.K + per_line - 1) / per_line * per_line : c.K;
  A.init(a_z * c.M * ldk, 11, 2.0f, bf);
  B.init((int64_t)c.Z * c.N * ldk, 22, 0.2f, bf);
  const float sA = 256.f, sB = 4096.f;     // |A| < 1 -> < 2^8, |B| < 0.1 -> < 2^9
  float* unscale = nullptr;
  if (f16) {
    A.init_f16(sA);
    B.init_f16(sB);
    const float u = 1.f / (sA * sB);
    CK(cudaMalloc(&unscale, 4));
    CK(cudaMemcpy(unscale, &u, 4, cudaMemcpyHostToDevice));
  }
  bias.init((int64_t)c.Z * c.N + 1, 33, 1.0f, false);
  const int out_z = c.reduce ? c.slots : c.Z;
  act.init((int64_t)out_z * c.M * c.N, 44, 1.0f, false);
  float *out = nullptr, *out_lo = nullptr;
  const int64_t on = (int64_t)out_z * c.M * c.N;
  CK(cudaMalloc(&out, on * 4));
  CK(cudaMemset(out, 0xFF, on * 4));
  if (c.split_out) { CK(cudaMalloc(&out_lo, on * 4)); CK(cudaMemset(out_lo, 0xFF, on * 4)); }

  GemmDesc d;
  d.kblock_bytes = g_kblock;
  d.pair = g_pair;
  d.debug_skip_mma = g_skip;
  d.pair_relay = g_relay;
  d.spin_wait = g_spin;
  d.mode = c.mode; d.M = c.M; d.N = c.N; d.K = c.K; d.Z = c.Z; d.BN = c.BN;
  d.A.hi = bf ? (void*)A.bf : (void*)A.hi; d.A.lo = A.lo; d.A.rows = c.M; d.A.ld = ldk;
  d.unscale = unscale;
  d.A.zstride = c.a_per_z ? (int64_t)c.M * ldk : 0;
  d.B.hi = bf ? (void*)B.bf : (void*)B.hi; d.B.lo = B.lo; d.B.rows = c.N; d.B.ld = ldk; d.B.zstride = (int64_t)c.N * ldk;
  if (f16) { d.A.hi = A.h16; d.A.lo = A.l16; d.B.hi = B.h16; d.B.lo = B.l16; }
  d.reduce_z = c.reduce; d.slots = c.slots; d.epi = c.epi;
  d.bias = bias.hi + 1; d.bias_zstride = c.N;     // +1: bias rows are only 4-byte aligned in the bank
  d.act = act.hi; d.act_zstride = (int64_t)c.M * c.N; d.act_ld = c.N;
  d.out = g_noout ? nullptr : out; d.out_lo = out_lo; d.out_ld = c.N; d.out_zstride = (int64_t)c.M * c.N;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  d.sm_count = prop.multiProcessorCount;
  std::string err;
  if (gemm(d, 0, &err)) { printf("%s: gemm failed: %s\n", c.name, err.c_str()); exit(3); }
  CK(cudaDeviceSynchronize());

  if (timing_iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < timing_iters; ++i) gemm(d, 0, &err);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / timing_iters;
  }

  std::vector<float> h(on), hl;
  CK(cudaMemcpy(h.data(), out, on * 4, cudaMemcpyDeviceToHost));
  if (c.split_out) { hl.resize(on); CK(cudaMemcpy(hl.data(), out_lo, on * 4, cudaMemcpyDeviceToHost)); }
  std::vector<float> hb((int64_t)c.Z * c.N), hact;
  CK(cudaMemcpy(hb.data(), bias.hi + 1, hb.size() * 4, cudaMemcpyDeviceToHost));
  if (c.epi == EPI_MASK) { hact.resize(on); CK(cudaMemcpy(hact.data(), act.hi, on * 4, cudaMemcpyDeviceToHost)); }

  double max_err = 0, max_ref = 0;
  const int64_t nchk = full_check ? on : 4096;
  for (int64_t q = 0; q < nchk; ++q) {
    const int64_t idx = full_check ? q : (int64_t)(hash32(q, 777) % (uint64_t)on);
    const int zz = (int)(idx / ((int64_t)c.M * c.N));
    const int m = (int)((idx / c.N) % c.M), n = (int)(idx % c.N);
    const int z0 = c.reduce ? (int)((long long)zz * c.Z / c.slots) : zz;
    const int z1 = c.reduce ? (int)((long long)(zz + 1) * c.Z / c.slots) : zz + 1;
    double acc = 0;
    for (int z = z0; z < z1; ++z) {
      const int64_t ao = (c.a_per_z ? (int64_t)z * c.M * ldk : 0) + (int64_t)m * ldk;
      const int64_t bo = ((int64_t)z * c.N + n) * ldk;
      for (int k = 0; k < c.K; ++k) {
        float a = A.at(ao + k), b = B.at(bo + k);
        if (bf) { a = bf16_round(a); b = bf16_round(b); }
        acc += (double)a * (double)b;
      }
    }
    if (c.epi == EPI_BIAS_LEAKY || c.epi == EPI_BIAS) acc += hb[(int64_t)zz * c.N + n];
    if (c.epi == EPI_BIAS_LEAKY) acc = acc > 0 ? acc : acc * 0.01;
    if (c.epi == EPI_MASK) acc = hact[idx] > 0.f ? acc : acc * 0.01;
    double got = h[idx];
    if (c.split_out) got += hl[idx];
    max_err = fmax(max_err, fabs(got - acc));
    max_ref = fmax(max_ref, fabs(acc));
  }
  A.free_(); B.free_(); bias.free_(); act.free_();
  cudaFree(out); cudaFree(out_lo); cudaFree(unscale);
  return max_err / fmax(max_ref, 1e-30);
}

// Implicit-GEMM operand of the 5x5 convolution (GemmDesc::conv_images): A = channels-last maps [Z][images][12][12][32],
// B = [Z][H][800] with K ordered (ky, kx, c); out[z][image * 64 + oy * 8 + ox][h] vs a double-precision host reference.
static double run_conv_case(int images, int Z, int H, int BN, int timing_iters, double* ms_out, int mode = MODE_TF32X3) {
  Dev A, B, bias;
  A.init((int64_t)Z * images * 144 * 32, 55, 2.0f, false);
  B.init((int64_t)Z * H * 800, 66, 0.2f, false);
  float* unscale = nullptr;
  if (mode == MODE_F16X3) {
    A.init_f16(256.f);
    B.init_f16(4096.f);
    const float u = 1.f / (256.f * 4096.f);
    CK(cudaMalloc(&unscale, 4));
    CK(cudaMemcpy(unscale, &u, 4, cudaMemcpyHostToDevice));
  }
  bias.init((int64_t)Z * H + 1, 77, 1.0f, false);
  const int M = images * 64;
  const int64_t on = (int64_t)Z * M * H;
  float* out = nullptr;
  CK(cudaMalloc(&out, on * 4));
  CK(cudaMemset(out, 0xFF, on * 4));
  GemmDesc d;
  d.mode = mode; d.M = M; d.N = H; d.K = 800; d.Z = Z; d.BN = BN;
  d.conv_images = images;
  d.A.hi = A.hi; d.A.lo = A.lo; d.A.rows = M; d.A.ld = 800; d.A.zstride = (int64_t)images * 144 * 32;
  d.B.hi = B.hi; d.B.lo = B.lo; d.B.rows = H; d.B.ld = 800; d.B.zstride = (int64_t)H * 800;
  if (mode == MODE_F16X3) {
    d.kblock_bytes = 64;
    d.A.hi = A.h16; d.A.lo = A.l16; d.B.hi = B.h16; d.B.lo = B.l16;
    d.unscale = unscale;
  }
  d.epi = EPI_BIAS_LEAKY;
  d.bias = bias.hi + 1; d.bias_zstride = H;
  d.out = out; d.out_ld = H; d.out_zstride = (int64_t)M * H;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  d.sm_count = prop.multiProcessorCount;
  std::string err;
  if (gemm(d, 0, &err)) { printf("conv: gemm failed: %s\n", err.c_str()); exit(3); }
  CK(cudaDeviceSynchronize());
  if (timing_iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < timing_iters; ++i) gemm(d, 0, &err);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / timing_iters;
  }
  std::vector<float> h(on), hb((int64_t)Z * H);
  CK(cudaMemcpy(h.data(), out, on * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), bias.hi + 1, hb.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  const int64_t nchk = on <= 400000 ? on : 8192;
  for (int64_t q = 0; q < nchk; ++q) {
    const int64_t idx = nchk == on ? q : (int64_t)(hash32(q, 778) % (uint64_t)on);
    const int z = (int)(idx / ((int64_t)M * H));
    const int m = (int)((idx / H) % M), n = (int)(idx % H);
    const int img = m / 64, oy = (m % 64) / 8, ox = m % 8;
    double acc = 0;
    for (int ky = 0; ky < 5; ++ky)
      for (int kx = 0; kx < 5; ++kx)
        for (int c = 0; c < 32; ++c)
          acc += (double)A.at((((int64_t)z * images + img) * 144 + (oy + ky) * 12 + ox + kx) * 32 + c) *
                 (double)B.at(((int64_t)z * H + n) * 800 + (ky * 5 + kx) * 32 + c);
    acc += hb[(int64_t)z * H + n];
    acc = acc > 0 ? acc : acc * 0.01;
    max_err = fmax(max_err, fabs((double)h[idx] - acc));
    max_ref = fmax(max_ref, fabs(acc));
  }
  A.free_(); B.free_(); bias.free_();
  cudaFree(out); cudaFree(unscale);
  return max_err / fmax(max_ref, 1e-30);
}

int main(int argc, char** argv) {
  const bool bench = argc > 1 && !strcmp(argv[1], "bench");
  int fails = 0;
  const Case small[] = {
      {"tf32x3 fwd ragged M, K tail, bias+leaky", MODE_TF32X3, 300, 512, 784, 3, 256, 0, 1, 0, EPI_BIAS_LEAKY, 0},
      {"tf32x3 bwd reduce_z, BN=208, N tail", MODE_TF32X3, 300, 784, 512, 5, 208, 1, 2, 1, EPI_NONE, 0},
      {"tf32x3 mask epilogue, split output", MODE_TF32X3, 200, 512, 512, 2, 256, 0, 1, 1, EPI_MASK, 1},
      {"tf32x3 tiny M=7, H=64", MODE_TF32X3, 7, 64, 784, 4, 64, 0, 1, 0, EPI_BIAS, 0},
      {"tf32x3 one tile, many z (phase wrap)", MODE_TF32X3, 128, 256, 64, 9, 256, 1, 1, 1, EPI_NONE, 0},
      {"f16x3 fwd ragged M, K tail, bias+leaky", MODE_F16X3, 300, 512, 784, 3, 256, 0, 1, 0, EPI_BIAS_LEAKY, 0},
      {"f16x3 bwd reduce_z, BN=208, N tail", MODE_F16X3, 300, 784, 512, 5, 208, 1, 2, 1, EPI_NONE, 0},
      {"f16x3 tiny M=7, H=64", MODE_F16X3, 7, 64, 784, 4, 64, 0, 1, 0, EPI_BIAS, 0},
      {"f16x3 one tile, many z (phase wrap)", MODE_F16X3, 128, 256, 64, 9, 256, 1, 1, 1, EPI_NONE, 0},
      {"f16x3 bwd reduce_z, BN=160 (3 stages)", MODE_F16X3, 300, 784, 512, 5, 160, 1, 2, 1, EPI_NONE, 0},
      {"f16x3 bwd reduce_z, BN=112 (3 stages)", MODE_F16X3, 300, 784, 512, 5, 112, 1, 2, 1, EPI_NONE, 0},
      {"tf32x3 fwd BN=128 (3 stages)", MODE_TF32X3, 300, 512, 784, 3, 128, 0, 1, 0, EPI_BIAS_LEAKY, 0},
      {"bf16 fwd", MODE_BF16, 300, 512, 784, 3, 256, 0, 1, 0, EPI_BIAS_LEAKY, 0},
      {"bf16 bwd reduce_z", MODE_BF16, 300, 784, 512, 5, 208, 1, 2, 1, EPI_NONE, 0},
  };
  const char* cm = getenv("TC_CFGS");
  const int cfg_mask = cm ? atoi(cm) : 15;
  g_skip = getenv("TC_SKIP_MMA") ? atoi(getenv("TC_SKIP_MMA")) : 0;
  g_relay = getenv("TC_RELAY") ? atoi(getenv("TC_RELAY")) : 1;
  g_spin = getenv("TC_SPIN") ? atoi(getenv("TC_SPIN")) : 0;
  g_ldpad = getenv("TC_LDPAD") ? atoi(getenv("TC_LDPAD")) : 0;
  g_noout = getenv("TC_NOOUT") ? atoi(getenv("TC_NOOUT")) : 0;    // timing only: the epilogue drains TMEM but stores nothing
  for (int cfg = 0; cfg < 4; ++cfg) {
  if (!((cfg_mask >> cfg) & 1)) continue;
  g_kblock = (cfg & 1) ? 128 : 64;
  g_pair = (cfg & 2) ? 0 : 1;
  printf("---- K-block %d bytes, %s ----\n", g_kblock, g_pair ? "CTA pairs (cta_group::2)" : "single CTA");
  const int only_mode = getenv("TC_MODE") ? atoi(getenv("TC_MODE")) : -1;   // restrict to one precision mode
  for (const Case& c : small) {
    if (only_mode >= 0 && c.mode != only_mode) continue;
    double ms = 0;
    const double e = run_case(c, true, 0, &ms);
    const double tol = 3e-5;   // tensor-core fp32 accumulation truncates: ~5e-6 (K=784) .. 2e-5 (K=2048) of the output max
    printf("%-45s rel err %.3e  %s\n", c.name, e, e < tol ? "ok" : "FAIL");
    if (!(e < tol) && !g_skip) fails++;
  }
  if (!bench) continue;
  {
    const Case big[] = {
        {"tf32x3 fwd 10000x512x784 Z=148 BN=256", MODE_TF32X3, 10000, 512, 784, 148, 256, 0, 1, 0, EPI_BIAS_LEAKY, 0},
        {"tf32x3 fwd 10000x512x784 Z=148 BN=128", MODE_TF32X3, 10000, 512, 784, 148, 128, 0, 1, 0, EPI_BIAS_LEAKY, 0},
        {"tf32x3 bwd 10000x784x512 Z=148 BN=208 s37", MODE_TF32X3, 10000, 784, 512, 148, 208, 1, 37, 1, EPI_NONE, 0},
        {"tf32x3 bwd 10000x784x512 Z=148 BN=256 s37", MODE_TF32X3, 10000, 784, 512, 148, 256, 1, 37, 1, EPI_NONE, 0},
        {"tf32x3 bwd 10000x784x512 Z=148 BN=112 s21", MODE_TF32X3, 10000, 784, 512, 148, 112, 1, 21, 1, EPI_NONE, 0},
        {"f16x3 fwd 10000x512x784 Z=148 BN=256", MODE_F16X3, 10000, 512, 784, 148, 256, 0, 1, 0, EPI_BIAS_LEAKY, 0},
        {"f16x3 fwd 10000x512x768 Z=148 BN=256", MODE_F16X3, 10000, 512, 768, 148, 256, 0, 1, 0, EPI_BIAS_LEAKY, 0},
        {"f16x3 fwd 10000x512x784 Z=148 BN=160", MODE_F16X3, 10000, 512, 784, 148, 160, 0, 1, 0, EPI_BIAS_LEAKY, 0},
        {"f16x3 bwd 10000x784x512 Z=148 BN=208 s37", MODE_F16X3, 10000, 784, 512, 148, 208, 1, 37, 1, EPI_NONE, 0},
        {"f16x3 bwd 10000x784x512 Z=148 BN=256 s37", MODE_F16X3, 10000, 784, 512, 148, 256, 1, 37, 1, EPI_NONE, 0},
        {"f16x3 bwd 10000x784x512 Z=148 BN=160 s37", MODE_F16X3, 10000, 784, 512, 148, 160, 1, 37, 1, EPI_NONE, 0},
        {"f16x3 bwd 10000x784x512 Z=148 BN=112 s37", MODE_F16X3, 10000, 784, 512, 148, 112, 1, 37, 1, EPI_NONE, 0},
        {"f16x3 bwd 10000x784x512 Z=148 BN=128 s37", MODE_F16X3, 10000, 784, 512, 148, 128, 1, 37, 1, EPI_NONE, 0},
        {"f16x3 fwd 10000x512x784 Z=148 BN=128", MODE_F16X3, 10000, 512, 784, 148, 128, 0, 1, 0, EPI_BIAS_LEAKY, 0},
        {"tf32x3 bwd 10000x784x512 Z=148 BN=160 s37", MODE_TF32X3, 10000, 784, 512, 148, 160, 1, 37, 1, EPI_NONE, 0},
        {"bf16 fwd 10000x512x784 Z=148 BN=256", MODE_BF16, 10000, 512, 784, 148, 256, 0, 1, 0, EPI_BIAS_LEAKY, 0},
        {"bf16 bwd 10000x784x512 Z=148 BN=208 s37", MODE_BF16, 10000, 784, 512, 148, 208, 1, 37, 1, EPI_NONE, 0},
    };
    for (const Case& c : big) {
      if (only_mode >= 0 && c.mode != only_mode) continue;
      double ms = 0;
      const double e = run_case(c, false, 3, &ms);
      const double flop = 2.0 * c.M * c.N * (double)c.K * c.Z;
      printf("%-45s rel err %.3e  %.3f ms  %.1f TFLOP/s (algorithmic, 1 pass)\n", c.name, e, ms, flop / ms * 1e-9);
      const double tol = 3e-5;
      if (!(e < tol) && !g_skip) fails++;
    }
  }
  }
  if (!getenv("TC_NO_CONV")) {
    printf("---- implicit-GEMM conv operand (5-D TMA boxes), TF32X3, single CTA ----\n");
    const int cc[][4] = {{5, 3, 64, 64}, {2, 2, 512, 256}, {7, 2, 128, 128}, {1, 4, 32, 32}};
    for (const auto& c : cc) {
      double ms = 0;
      const double e = run_conv_case(c[0], c[1], c[2], c[3], 0, &ms);
      printf("conv images=%d Z=%d H=%d BN=%d                      rel err %.3e  %s\n", c[0], c[1], c[2], c[3], e, e < 3e-5 ? "ok" : "FAIL");
      if (!(e < 3e-5)) fails++;
    }
    printf("---- implicit-GEMM conv operand, F16X3 (64-byte K-blocks = 32 fp16 channels) ----\n");
    for (const auto& c : cc) {
      double ms = 0;
      const double e = run_conv_case(c[0], c[1], c[2], c[3], 0, &ms, MODE_F16X3);
      printf("conv f16x3 images=%d Z=%d H=%d BN=%d                rel err %.3e  %s\n", c[0], c[1], c[2], c[3], e, e < 3e-5 ? "ok" : "FAIL");
      if (!(e < 3e-5)) fails++;
    }
    if (bench) {
      double ms16 = 0;
      const double e16 = run_conv_case(100, 50, 512, 256, 3, &ms16, MODE_F16X3);
      printf("conv f16x3 images=100 Z=50 H=512 BN=256 (cfg4)     rel err %.3e  %.3f ms  %.1f TFLOP/s (algorithmic, 1 pass)\n", e16, ms16,
             2.0 * 6400 * 512 * 800.0 * 50 / ms16 * 1e-9);
      if (!(e16 < 3e-5)) fails++;
      double ms = 0;
      const double e = run_conv_case(100, 50, 512, 256, 3, &ms);
      printf("conv images=100 Z=50 H=512 BN=256 (cfg4)           rel err %.3e  %.3f ms  %.1f TFLOP/s (algorithmic, 1 pass)\n", e, ms,
             2.0 * 6400 * 512 * 800.0 * 50 / ms * 1e-9);
      if (!(e < 3e-5)) fails++;
    }
  }
  printf(fails ? "FAILED (%d)\n" : "ALL OK\n", fails);
  return fails ? 1 : 0;
}
