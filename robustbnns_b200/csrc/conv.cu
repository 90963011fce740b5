// Convolutional architecture (model_nn.py:98-106), per posterior sample z and input b:
//   x[1,28,28] -Conv2d(1,32,5)-> [32,24,24] -LeakyReLU-> -MaxPool2d(2)-> P1[32,12,12]
//     -Conv2d(32,H,5)-> [H,8,8] -LeakyReLU-> A2 -MaxPool2d(2,stride=1)-> P2[H,7,7] -Flatten-> Linear(49H,C)
// conv1 (K=25, <1% of the FLOPs) and both pools are direct kernels; conv2 (97% of the FLOPs) is
// lowered to GEMM through an explicit im2col (round 1; implicit-GEMM on tcgen05 is the next step).
// Internal activation layouts after conv2 are position-major / channel-minor (HWC) so the GEMM
// output is consumed as is; model.7.weight is permuted once per bank row to match (sampler.cu).
// Max-pool ties route to the FIRST maximum in window scan order, as torch's max_pool2d does.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace rbnn {

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : v * kLeakySlope; }

// ---- conv1 + leaky + maxpool(2): one block per (z, b) -------------------------------------
// The 25-term dot products are accumulated in fp64 (exact fp32 products), and the two decisions taken here -- the sign
// under LeakyReLU and the arg-max of the 2x2 pooling window -- are taken on those fp64 values, i.e. as exact arithmetic
// (and the fp64 oracle) takes them.  In fp32 about one (sample, image) unit in 250 of the conv-512 net has a first-layer
// near-tie that rounding decides the other way, and such a unit moves by 1e-3..1e-2 (profiles/r2_conv_cfg4_oracle.json).
// P1 is stored as fp32 plus, when p1lo != nullptr, its fp32 residual, so that the exact re-evaluation of conv2's
// near-ties (conv2_refine_kernel) sees the pooled map to ~2^-48 instead of 2^-24.
// Thread = (channel c = tid / 8, 9 pairs of horizontally adjacent pooled outputs): the 25 filter taps stay in registers
// for the whole block and one 8-wide input row feeds both outputs of the pair -- 48 shared-memory loads per 200 DFMAs
// (one output per thread with the taps re-read from shared memory was load-bound: 61 loads per 100 DFMAs).
// max_bits != nullptr: max|P1| of the launch is folded in (float bits, atomicMax) -- the F16X3 operand range.
__global__ void __launch_bounds__(256)
conv1_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ bank, int64_t P, int64_t cw1,
                      int64_t cb1, int s0, int B, float* __restrict__ p1, float* __restrict__ p1lo,
                      uint8_t* __restrict__ idx1, unsigned* __restrict__ max_bits) {
  __shared__ double xs[28 * 28];
  __shared__ float red[8];
  const int zb = blockIdx.x;
  const int z = zb / B, b = zb % B;
  const float* row = bank + (int64_t)(s0 + z) * P;
  for (int i = threadIdx.x; i < 784; i += blockDim.x) xs[i] = (double)__ldg(x + (int64_t)b * 784 + i);
  const int c = threadIdx.x >> 3, sub = threadIdx.x & 7;
  double wk[25];
#pragma unroll
  for (int k = 0; k < 25; ++k) wk[k] = (double)__ldg(row + cw1 + c * 25 + k);
  const double bias = (double)__ldg(row + cb1 + c);
  __syncthreads();
  float amax = 0.f;
#pragma unroll 1
  for (int it = 0; it < 9; ++it) {
    const int pr = sub + 8 * it;                       // pair index 0..71 of this channel
    const int py = pr / 6, px0 = (pr - py * 6) * 2;
    double acc[2][2][2];
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) acc[o][dy][dx] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {                      // one input row at a time: it feeds window rows dy = i - ky
      double r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = xs[(2 * py + i) * 28 + 2 * px0 + j];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int ky = i - dy;
        if (ky < 0 || ky > 4) continue;
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx)
              acc[o][dy][dx] = fma(r[2 * o + dx + kx], wk[ky * 5 + kx], acc[o][dy][dx]);
      }
    }
    float hi2[2], lo2[2];
    uint8_t bi2[2];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      double best = 0.0;
      int bi = 0;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          double v = acc[o][dy][dx] + bias;
          v = v > 0.0 ? v : v * (double)kLeakySlope;
          if ((dy == 0 && dx == 0) || v > best) { best = v; bi = dy * 2 + dx; }
        }
      hi2[o] = (float)best;
      lo2[o] = (float)(best - (double)hi2[o]);
      bi2[o] = (uint8_t)bi;
      amax = fmaxf(amax, fabsf(hi2[o]));
    }
    const int64_t at = (int64_t)zb * 4608 + c * 144 + py * 12 + px0;      // even offset: 8-byte stores
    *reinterpret_cast<float2*>(p1 + at) = make_float2(hi2[0], hi2[1]);
    if (p1lo) *reinterpret_cast<float2*>(p1lo + at) = make_float2(lo2[0], lo2[1]);
    *reinterpret_cast<uchar2*>(idx1 + at) = make_uchar2(bi2[0], bi2[1]);
  }
  if (max_bits) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 1; i < 8; ++i) amax = fmaxf(amax, red[i]);
      atomicMax(max_bits, __float_as_uint(amax));
    }
  }
}

int conv1_pool_fwd(rbnn_net* net, const float* x, const float* bank, int s0, int Z, int B, float* p1,
                   uint8_t* idx1, cudaStream_t st, float* p1lo, unsigned* max_bits) {
  conv1_pool_fwd_kernel<<<Z * B, 256, 0, st>>>(x, bank, net->L.P, net->L.cw1, net->L.cb1, s0, B, p1, p1lo, idx1, max_bits);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- im2col for conv2: col[(zb*64+pos)][c*25+ky*5+kx] = P1[zb][c][oy+ky][ox+kx] -------------
__global__ void im2col_conv2_kernel(const float* __restrict__ p1, int64_t total, float* __restrict__ col) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % 800);
  const int64_t r = i / 800;
  const int pos = (int)(r % 64);
  const int64_t zb = r / 64;
  const int c = k / 25, ky = (k % 25) / 5, kx = k % 5;
  const int oy = pos / 8, ox = pos % 8;
  col[i] = __ldg(p1 + zb * 4608 + c * 144 + (oy + ky) * 12 + (ox + kx));
}

int im2col_conv2(rbnn_net* net, const float* p1, int ZB, float* col, cudaStream_t st) {
  const int64_t total = (int64_t)ZB * 64 * 800;
  im2col_conv2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p1, total, col);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- maxpool(2, stride 1) on A2[zb][8*8][H] -> P2[zb][7*7][H] ------------------------------
__global__ void pool2_fwd_kernel(const float* __restrict__ a2, int64_t total, int H, float* __restrict__ p2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int h = (int)(i % H);
  const int64_t r = i / H;
  const int p = (int)(r % 49);
  const int64_t zb = r / 49;
  const int py = p / 7, px = p % 7;
  const float* base = a2 + (zb * 64 + py * 8 + px) * H + h;
  const float v = fmaxf(fmaxf(__ldg(base), __ldg(base + H)), fmaxf(__ldg(base + 8 * H), __ldg(base + 9 * H)));
  p2[i] = v;
}

int pool2_fwd(rbnn_net* net, const float* a2, int ZB, int H, float* p2, cudaStream_t st) {
  const int64_t total = (int64_t)ZB * 49 * H;
  pool2_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a2, total, H, p2);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// dZ2[zb][pos][h] = leaky'(A2) * sum over the <=4 windows containing pos whose first-max is pos
__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// F16X3 operand split of an already scaled value: hi = rn_f16(v), lo = rn_f16(v - hi)
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// dz2_lo != nullptr: the result is written tf32-split (hi = rn_tf32(v), lo = v - hi) for the tcgen05 dgrad GEMM
__global__ void pool2_bwd_kernel(const float* __restrict__ a2, const float* __restrict__ dp2, int64_t total, int H,
                                 float* __restrict__ dz2, float* __restrict__ dz2_lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int h = (int)(i % H);
  const int64_t r = i / H;
  const int pos = (int)(r % 64);
  const int64_t zb = r / 64;
  const int y = pos / 8, x = pos % 8;
  const float* A = a2 + zb * 64 * H + h;
  float acc = 0.f;
#pragma unroll
  for (int wy = y - 1; wy <= y; ++wy)
#pragma unroll
    for (int wx = x - 1; wx <= x; ++wx) {
      if (wy < 0 || wy > 6 || wx < 0 || wx > 6) continue;
      float best = __ldg(A + (wy * 8 + wx) * H);
      int by = wy, bx = wx;
#pragma unroll
      for (int j = 1; j < 4; ++j) {
        const int yy = wy + (j >> 1), xx = wx + (j & 1);
        const float v = __ldg(A + (yy * 8 + xx) * H);
        if (v > best) { best = v; by = yy; bx = xx; }
      }
      if (by == y && bx == x) acc += __ldg(dp2 + (zb * 49 + wy * 7 + wx) * H + h);
    }
  const float a = __ldg(A + pos * H);
  const float v = a > 0.f ? acc : acc * kLeakySlope;
  if (dz2_lo) {
    const float h = tf32_rn(v);
    dz2[i] = h;
    dz2_lo[i] = v - h;
  } else {
    dz2[i] = v;
  }
}

int pool2_bwd(rbnn_net* net, const float* a2, const float* dp2, int ZB, int H, float* dz2, cudaStream_t st,
              float* dz2_lo) {
  const int64_t total = (int64_t)ZB * 64 * H;
  pool2_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a2, dp2, total, H, dz2, dz2_lo);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- col2im (gather) + leaky' of the pooled conv1 activation ---------------------------------
// G1[zb][c][iy][ix] = leaky'(P1) * sum_{ky,kx} dcol[(zb*64 + (iy-ky)*8 + (ix-kx))][c*25+ky*5+kx]
// One block per (image, group of 4 channels): the image's 64 x 100 slice of dcol is staged in shared memory with
// coalesced 400-byte runs, then gathered (the direct gather strides 3200 bytes between neighbouring threads).
__global__ void __launch_bounds__(256)
col2im_conv2_kernel(const float* __restrict__ dcol, const float* __restrict__ p1, float* __restrict__ g1) {
  __shared__ __align__(16) float t[64 * 100];
  const int64_t zb = blockIdx.x;
  const int cg = blockIdx.y;
  // 16-byte loads: a position's 100 floats of this channel group start at a multiple of 400 bytes (the kernel was
  // instruction-bound on scalar loads and their index arithmetic at half the HBM rate)
  for (int i = threadIdx.x; i < 1600; i += blockDim.x) {
    const int pos = i / 25, j = i - pos * 25;
    reinterpret_cast<float4*>(t)[i] = __ldg(reinterpret_cast<const float4*>(dcol + (zb * 64 + pos) * 800 + cg * 100) + j);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 4 * 144; o += blockDim.x) {
    const int cl = o / 144, r = o - cl * 144, iy = r / 12, ix = r - iy * 12;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const int oy = iy - ky;
      if (oy < 0 || oy > 7) continue;
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int ox = ix - kx;
        if (ox < 0 || ox > 7) continue;
        acc += t[(oy * 8 + ox) * 100 + cl * 25 + ky * 5 + kx];
      }
    }
    const int64_t gi = zb * 4608 + (cg * 4 + cl) * 144 + r;
    g1[gi] = __ldg(p1 + gi) > 0.f ? acc : acc * kLeakySlope;
  }
}

int col2im_conv2(rbnn_net* net, const float* dcol, const float* p1, int ZB, float* g1, cudaStream_t st) {
  col2im_conv2_kernel<<<dim3(ZB, 8), 256, 0, st>>>(dcol, p1, g1);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- maxpool1 backward + conv1 input gradient, summed over the chunk's samples ---------------
// dx[b][iy][ix] (+)= sum_z sum_c sum_{ky,kx} R_z[c][iy-ky][ix-kx] * cw1_z[c][ky][kx],  where R is G1 routed back through
// MaxPool2d(2): R[c][oy][ox] = G1[c][oy/2][ox/2] if idx1 picked (oy, ox) in its 2x2 cell, else 0   (model_nn.py:98-100).
// Per (sample, image) the routed map of 16 channels at a time is scattered into shared memory (x padded by 4 zeros on
// both sides, row pitch 33 => conflict-free), then a thread owns a strip of 7 output pixels of one row: per (channel,
// ky) it loads 11 map entries + 5 weights and issues 35 FMAs (the first version tested idx1 for every (pixel, tap,
// channel): 6 instructions per FMA).  gridDim.y > 1: block (b, zi) sums its slice of the samples into
// partial[zi][b][784] (reduced in a fixed order by conv1_reduce_kernel); gridDim.y == 1: straight into dx[b][784].
constexpr int kC1Pitch = 33, kC1Chan = 16;
constexpr int kC1SmemFloats = kC1Chan * 24 * kC1Pitch + 800;

__global__ void __launch_bounds__(128)
conv1_bwd_sum_kernel(const float* __restrict__ g1, const uint8_t* __restrict__ idx1, const float* __restrict__ bank,
                     int64_t P, int64_t cw1, int s0, int Z, int B, float* __restrict__ dx, int accumulate) {
  extern __shared__ float c1sm[];
  float* R = c1sm;                                   // [16][24][33]
  float* ws = c1sm + kC1Chan * 24 * kC1Pitch;        // [32][25]
  const int b = blockIdx.x;
  const int z_begin = (int)((long long)blockIdx.y * Z / gridDim.y), z_end = (int)((long long)(blockIdx.y + 1) * Z / gridDim.y);
  dx += (int64_t)blockIdx.y * B * 784;
  const int t = threadIdx.x, iy = t >> 2, strip = t & 3;
  const bool active = iy < 28;
  float acc[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) acc[j] = 0.f;
  for (int i = t; i < kC1Chan * 24 * kC1Pitch; i += blockDim.x) R[i] = 0.f;
  int prev[kC1Chan * 144 / 128];                     // where this thread's cells were scattered last time (-1: nowhere)
#pragma unroll
  for (int q = 0; q < kC1Chan * 144 / 128; ++q) prev[q] = -1;
  for (int z = z_begin; z < z_end; ++z) {
    const int64_t zb = (int64_t)z * B + b;
    __syncthreads();                                 // the previous sample's readers are done with ws / R
    for (int i = t; i < 800; i += blockDim.x) ws[i] = __ldg(bank + (int64_t)(s0 + z) * P + cw1 + i);
#pragma unroll 1
    for (int half = 0; half < 32 / kC1Chan; ++half) {
#pragma unroll
      for (int q = 0; q < kC1Chan * 144 / 128; ++q) {
        const int i = t + q * 128;                   // (local channel, cell): always handled by this thread
        const int c = i / 144, cell = i - c * 144, cy = cell / 12, cx = cell - 12 * cy;
        const int64_t gi = zb * 4608 + (half * kC1Chan + c) * 144 + cell;
        const int sub = idx1[gi];
        const int o = (c * 24 + 2 * cy + (sub >> 1)) * kC1Pitch + 2 * cx + (sub & 1) + 4;
        if (prev[q] >= 0) R[prev[q]] = 0.f;
        R[o] = __ldg(g1 + gi);
        prev[q] = o;
      }
      __syncthreads();
      if (active) {
#pragma unroll 2
        for (int c = 0; c < kC1Chan; ++c) {
#pragma unroll
          for (int ky = 0; ky < 5; ++ky) {
            const int oy = iy - ky;
            if (oy < 0 || oy > 23) continue;
            const float* __restrict__ row = R + (c * 24 + oy) * kC1Pitch + strip * 7;
            const float* __restrict__ w = ws + (half * kC1Chan + c) * 25 + ky * 5;
            float r[11], wv[5];
#pragma unroll
            for (int k = 0; k < 11; ++k) r[k] = row[k];
#pragma unroll
            for (int k = 0; k < 5; ++k) wv[k] = w[k];
#pragma unroll
            for (int j = 0; j < 7; ++j)
#pragma unroll
              for (int kx = 0; kx < 5; ++kx) acc[j] = fmaf(r[j - kx + 4], wv[kx], acc[j]);
          }
        }
      }
      __syncthreads();                               // readers done before the next scatter rewrites R
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int64_t o = (int64_t)b * 784 + iy * 28 + strip * 7 + j;
      dx[o] = accumulate ? dx[o] + acc[j] : acc[j];
    }
  }
}

// dx[i] = (accumulate ? dx[i] : 0) + sum_p partial[p][i], fixed order
__global__ void conv1_reduce_kernel(const float* __restrict__ partial, int parts, int64_t n, float* __restrict__ dx,
                                    int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = accumulate ? dx[i] : 0.f;
  for (int p = 0; p < parts; ++p) s += __ldg(partial + (int64_t)p * n + i);
  dx[i] = s;
}

// slices of the sample range one image's blocks are cut into (fills the SMs when the batch is small)
int conv1_bwd_parts(const rbnn_net* net, int Z, int B) {
  const int want = (4 * net->sm_count + B - 1) / B;
  return std::max(1, std::min(std::min(want, Z), 32));
}

int conv1_bwd_sum(rbnn_net* net, const float* g1, const uint8_t* idx1, const float* bank, int s0, int Z, int B,
                  float* dx_sum, int accumulate, cudaStream_t st, float* partial, int parts) {
  const size_t smem = (size_t)kC1SmemFloats * sizeof(float);
  if (!net->conv1_smem_set) {       // per handle (= per device): the opt-in above 48 KB of dynamic shared memory
    RBNN_CUDA(cudaFuncSetAttribute(conv1_bwd_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    net->conv1_smem_set = 1;
  }
  if (!partial || parts <= 1) {
    conv1_bwd_sum_kernel<<<B, 128, smem, st>>>(g1, idx1, bank, net->L.P, net->L.cw1, s0, Z, B, dx_sum, accumulate);
    net->launches++;
    RBNN_CUDA(cudaGetLastError());
    return 0;
  }
  conv1_bwd_sum_kernel<<<dim3(B, parts), 128, smem, st>>>(g1, idx1, bank, net->L.P, net->L.cw1, s0, Z, B, partial, 0);
  const int64_t n = (int64_t)B * 784;
  conv1_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, parts, n, dx_sum, accumulate);
  net->launches += 2;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- MaxPool2d(2, stride 1) + Linear(49H, C), fused ------------------------------------------------
// logits[zb][c] = bo_z[c] + sum_{p, h} max(A2 window p)[h] * woutp_z[c][p*H + h]      (model_nn.py:103-106)
// The pooled map is never written.  A block owns kPoolImgs images of one sample (every output weight it loads is used
// for all of them); a thread owns channels h, h + 128, ... and walks the 7x7 windows row by row with the two A2 rows a
// window row touches in registers, so every A2 entry is loaded once (the position-major form re-read each entry for
// its 4 windows and thrashed L1: 21 % hit rate).  Per-thread partial sums, then a fixed-order block reduction.
constexpr int kPoolImgs = 2;

template <int C_MAX>
__global__ void __launch_bounds__(128)
pool2_logits_kernel(const float* __restrict__ a2, const float* __restrict__ woutp, int s0, int B, int H, int C,
                    float* __restrict__ partial) {
  // grid (channel chunks of blockDim.x, image groups, samples): partial[chunk][zb][c], summed by logits_reduce_kernel
  __shared__ float red[4][kPoolImgs][C_MAX];
  const int z = blockIdx.z, b0 = blockIdx.y * kPoolImgs;
  const int64_t F = (int64_t)49 * H;
  const int nimg = min(kPoolImgs, B - b0);
  float acc[kPoolImgs][C_MAX];
#pragma unroll
  for (int q = 0; q < kPoolImgs; ++q)
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) acc[q][c] = 0.f;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < H) {
    const float4* __restrict__ W = reinterpret_cast<const float4*>(woutp + ((int64_t)(s0 + z) * F + h) * C_MAX);
    const float* A[kPoolImgs];
#pragma unroll
    for (int q = 0; q < kPoolImgs; ++q) A[q] = a2 + ((int64_t)z * B + b0 + min(q, nimg - 1)) * 64 * H + h;
    float r0[kPoolImgs][8], r1[kPoolImgs][8];
#pragma unroll
    for (int q = 0; q < kPoolImgs; ++q)
#pragma unroll
      for (int x = 0; x < 8; ++x) r0[q][x] = __ldg(A[q] + (int64_t)x * H);
#pragma unroll 1
    for (int wy = 0; wy < 7; ++wy) {
#pragma unroll
      for (int q = 0; q < kPoolImgs; ++q)
#pragma unroll
        for (int x = 0; x < 8; ++x) r1[q][x] = __ldg(A[q] + (int64_t)((wy + 1) * 8 + x) * H);
#pragma unroll
      for (int wx = 0; wx < 7; ++wx) {
        float v[kPoolImgs];
#pragma unroll
        for (int q = 0; q < kPoolImgs; ++q)
          v[q] = q < nimg ? fmaxf(fmaxf(r0[q][wx], r0[q][wx + 1]), fmaxf(r1[q][wx], r1[q][wx + 1])) : 0.f;
        const float4* __restrict__ w = W + (int64_t)(wy * 7 + wx) * H * (C_MAX / 4);
#pragma unroll
        for (int j = 0; j < C_MAX / 4; ++j) {
          const float4 wv = __ldg(w + j);                 // classes 4j .. 4j+3 of this (window, channel); padding is zero
#pragma unroll
          for (int q = 0; q < kPoolImgs; ++q) {
            acc[q][4 * j + 0] = fmaf(v[q], wv.x, acc[q][4 * j + 0]);
            acc[q][4 * j + 1] = fmaf(v[q], wv.y, acc[q][4 * j + 1]);
            acc[q][4 * j + 2] = fmaf(v[q], wv.z, acc[q][4 * j + 2]);
            acc[q][4 * j + 3] = fmaf(v[q], wv.w, acc[q][4 * j + 3]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kPoolImgs; ++q)
#pragma unroll
        for (int x = 0; x < 8; ++x) r0[q][x] = r1[q][x];
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < kPoolImgs; ++q)
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) {
      float t = acc[q][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) red[warp][q][c] = t;
    }
  __syncthreads();
  const int nwarps = blockDim.x >> 5;
  for (int j = threadIdx.x; j < nimg * C; j += blockDim.x) {
    const int q = j / C, c = j - q * C;
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += red[w][q][c];
    partial[((int64_t)blockIdx.x * gridDim.z * B + (int64_t)z * B + b0 + q) * C + c] = t;
  }
}

// logits[zb][c] = bo_z[c] + sum_chunk partial[chunk][zb][c]   (fixed order)
__global__ void logits_reduce_kernel(const float* __restrict__ partial, int chunks, const float* __restrict__ bank,
                                     int64_t P, int64_t bo_off, int s0, int B, int C, int64_t n, float* __restrict__ logits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  const int z = (int)(i / ((int64_t)B * C));
  float t = 0.f;
  for (int k = 0; k < chunks; ++k) t += __ldg(partial + (int64_t)k * n + i);
  logits[i] = t + __ldg(bank + (int64_t)(s0 + z) * P + bo_off + c);
}

// specialisations on the class pitch of woutp (conv_class_pitch): 4, 12 (MNIST / F-MNIST: 10 classes), 16, 32
#define RBNN_CONV_C_DISPATCH(KERNEL, GRID, BLOCK, ...)                                      \
  do {                                                                                      \
    const int cp_ = conv_class_pitch(net->C);                                               \
    if (cp_ == 4) KERNEL<4><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__);                            \
    else if (cp_ == 12) KERNEL<12><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__);                     \
    else if (cp_ == 16) KERNEL<16><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__);                     \
    else KERNEL<32><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__);                                    \
  } while (0)

int pool2_logits_chunks(const rbnn_net* net) { return (net->H + 127) / 128; }

// partial: pool2_logits_chunks(net) * Z * B * C floats of scratch
int pool2_logits(rbnn_net* net, const float* a2, int s0, int Z, int B, float* logits, float* partial, cudaStream_t st) {
  const int threads = std::min(128, (net->H + 31) / 32 * 32);
  const int chunks = (net->H + threads - 1) / threads;
  dim3 grid(chunks, (B + kPoolImgs - 1) / kPoolImgs, Z);
  RBNN_CONV_C_DISPATCH(pool2_logits_kernel, grid, threads, a2, net->woutp, s0, B, net->H, net->C, partial);
  const int64_t n = (int64_t)Z * B * net->C;
  logits_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, chunks, net->bank, net->L.P, net->L.bo, s0, B,
                                                                    net->C, n, logits);
  net->launches += 2;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- input gradient of Linear(49H, C) + MaxPool2d(2, stride 1) + LeakyReLU, fused ---------------------
// dZ2[zb][pos][h] = leaky'(A2[pos][h]) * sum over the <= 4 windows w containing pos whose FIRST maximum (scan order,
// as torch's max_pool2d) is pos of  sum_c dlogits[zb][c] * woutp_z[c][w*H + h]          -- dP2 is never written.
// A thread owns one channel h of kBwdImgs images and walks the 7x7 windows row by row: every output weight is loaded
// once per block (window-major; the position-major form re-read each weight from L2 four times), the two A2 rows a
// window row touches and their gradient accumulators live in registers (all indices static after unrolling), and a
// position receives its <= 4 contributions in the fixed order (wy-1,wx-1), (wy-1,wx), (wy,wx-1), (wy,wx).
// dz2_lo != nullptr: the result is written tf32-split (hi = rn_tf32(v), lo = v - hi) for the tcgen05 dgrad GEMM.
constexpr int kBwdImgs = 2;

template <int C_MAX>
__global__ void __launch_bounds__(256)
pool2_bwd_fused_kernel(const float* __restrict__ a2, const float* __restrict__ dlogits, const float* __restrict__ woutp,
                       int s0, int B, int H, int C, void* __restrict__ dz2v, void* __restrict__ dz2_lov,
                       const float* __restrict__ f16_scale) {
  float* __restrict__ dz2 = reinterpret_cast<float*>(dz2v);
  float* __restrict__ dz2_lo = reinterpret_cast<float*>(dz2_lov);
  const float f16s = f16_scale ? __ldg(f16_scale) : 0.f;
  // blockDim.y image pairs share a block: they read the same output weights at nearly the same time, so all but the
  // first read of a line hit L1 (one pair per block re-read the block's 301 KB weight slice from L2 for every pair)
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  const int z = blockIdx.z, b0 = (blockIdx.y * blockDim.y + threadIdx.y) * kBwdImgs;
  if (h >= H || b0 >= B) return;
  const int nimg = min(kBwdImgs, B - b0);
  const int64_t F = (int64_t)49 * H;
  const float4* __restrict__ W = reinterpret_cast<const float4*>(woutp + ((int64_t)(s0 + z) * F + h) * C_MAX);
  float dl[kBwdImgs][C_MAX];
  const float* A[kBwdImgs];
  int64_t obase[kBwdImgs];
#pragma unroll
  for (int q = 0; q < kBwdImgs; ++q) {
    const int64_t zb = (int64_t)z * B + b0 + min(q, nimg - 1);
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) dl[q][c] = (q < nimg && c < C) ? __ldg(dlogits + zb * C + c) : 0.f;
    obase[q] = zb * 64 * H + h;
    A[q] = a2 + obase[q];
  }
  float r0[kBwdImgs][8], r1[kBwdImgs][8], acc0[kBwdImgs][8], acc1[kBwdImgs][8];
#pragma unroll
  for (int q = 0; q < kBwdImgs; ++q)
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      r0[q][x] = __ldg(A[q] + (int64_t)x * H);
      acc0[q][x] = 0.f;
      acc1[q][x] = 0.f;
    }
  auto store_row = [&](int row) {
#pragma unroll
    for (int q = 0; q < kBwdImgs; ++q) {
      if (q >= nimg) continue;
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const float v = r0[q][x] > 0.f ? acc0[q][x] : acc0[q][x] * kLeakySlope;
        const int64_t o = obase[q] + (int64_t)(row * 8 + x) * H;
        if (f16_scale) {
          split_f16(v * f16s, reinterpret_cast<__half*>(dz2v)[o], reinterpret_cast<__half*>(dz2_lov)[o]);
        } else if (dz2_lo) {
          const float hv = tf32_rn(v);
          dz2[o] = hv;
          dz2_lo[o] = v - hv;
        } else {
          dz2[o] = v;
        }
      }
    }
  };
#pragma unroll 1
  for (int wy = 0; wy < 7; ++wy) {
#pragma unroll
    for (int q = 0; q < kBwdImgs; ++q)
#pragma unroll
      for (int x = 0; x < 8; ++x) r1[q][x] = __ldg(A[q] + (int64_t)((wy + 1) * 8 + x) * H);
#pragma unroll
    for (int wx = 0; wx < 7; ++wx) {
      float wv[C_MAX];
#pragma unroll
      for (int j = 0; j < C_MAX / 4; ++j) {
        const float4 t = __ldg(W + (int64_t)(wy * 7 + wx) * H * (C_MAX / 4) + j);
        wv[4 * j] = t.x; wv[4 * j + 1] = t.y; wv[4 * j + 2] = t.z; wv[4 * j + 3] = t.w;
      }
#pragma unroll
      for (int q = 0; q < kBwdImgs; ++q) {
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < C_MAX; ++c) d = fmaf(dl[q][c], wv[c], d);
        float best = r0[q][wx];
        int k = 0;
        if (r0[q][wx + 1] > best) { best = r0[q][wx + 1]; k = 1; }
        if (r1[q][wx] > best) { best = r1[q][wx]; k = 2; }
        if (r1[q][wx + 1] > best) { k = 3; }
        acc0[q][wx] += k == 0 ? d : 0.f;
        acc0[q][wx + 1] += k == 1 ? d : 0.f;
        acc1[q][wx] += k == 2 ? d : 0.f;
        acc1[q][wx + 1] += k == 3 ? d : 0.f;
      }
    }
    store_row(wy);
#pragma unroll
    for (int q = 0; q < kBwdImgs; ++q)
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        acc0[q][x] = acc1[q][x];
        acc1[q][x] = 0.f;
        r0[q][x] = r1[q][x];
      }
  }
  store_row(7);
}

int pool2_bwd_fused(rbnn_net* net, const float* a2, const float* dlogits, int s0, int Z, int B, void* dz2,
                    void* dz2_lo, const float* f16_scale, cudaStream_t st) {
  const int threads = std::min(128, (net->H + 31) / 32 * 32);
  static const int pairs_env = getenv("RBNN_POOL_PAIRS") ? atoi(getenv("RBNN_POOL_PAIRS")) : 2;     // image pairs per block
  const int pairs = std::max(1, std::min(pairs_env, 256 / threads));
  const int per_block = kBwdImgs * pairs;
  dim3 grid((net->H + threads - 1) / threads, (B + per_block - 1) / per_block, Z);
  const dim3 block(threads, pairs);
  RBNN_CONV_C_DISPATCH(pool2_bwd_fused_kernel, grid, block, a2, dlogits, net->woutp, s0, B, net->H, net->C, dz2, dz2_lo,
                       f16_scale);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- tcgen05 route (tc_conv.cu): operand preparation and guard-band refinement of conv2 ---------------------
// P1[zb][c][y][x] (fp32, CHW) -> channels-last split copies, one block per image, transposed through shared memory.
//   TF32X3: hi/lo[zb][y][x][32] fp32 -- a pixel's 32 channels are the 128-byte K-block row of the implicit GEMM.
//   F16X3 : hi/lo[zb][y][x][64] fp16 -- the channels of pixel (y, x) FOLLOWED BY those of pixel (y, x + 1) (zeros past
//           the row end), so that one 128-byte row carries two horizontally adjacent filter taps: K-blocks of 64 bytes
//           (one tap) left the tensor pipe 46 % busy (twice the TMA requests and barrier round trips per byte).
__global__ void __launch_bounds__(256)
p1_split_hwc_kernel(const float* __restrict__ p1, void* __restrict__ hi, void* __restrict__ lo,
                    const float* __restrict__ f16_scale) {
  __shared__ float t[32 * 145];
  const int64_t zb = blockIdx.x;
  for (int i = threadIdx.x; i < 4608; i += blockDim.x) t[(i / 144) * 145 + i % 144] = __ldg(p1 + zb * 4608 + i);
  __syncthreads();
  const float sc = f16_scale ? __ldg(f16_scale) : 1.f;
  if (f16_scale) {
    __half* h16 = reinterpret_cast<__half*>(hi) + zb * 9216;
    __half* l16 = reinterpret_cast<__half*>(lo) + zb * 9216;
    for (int o = threadIdx.x; o < 4608; o += blockDim.x) {      // pairs of output elements: 4-byte stores
      const int e = (o & 31) * 2, px = o >> 5;                   // element pair e, e + 1 of the 64-element row of pixel px
      const int c = e & 31, src = px + (e >> 5);                 // second half: the next pixel of the same image row
      const bool in = (e < 32) || (px % 12) != 11;
      __half2 hh, ll;
      split_f16(in ? t[c * 145 + src] * sc : 0.f, hh.x, ll.x);
      split_f16(in ? t[(c + 1) * 145 + src] * sc : 0.f, hh.y, ll.y);
      reinterpret_cast<__half2*>(h16)[px * 32 + (e >> 1)] = hh;
      reinterpret_cast<__half2*>(l16)[px * 32 + (e >> 1)] = ll;
    }
    return;
  }
  for (int o = threadIdx.x; o < 4608; o += blockDim.x) {
    const int c = o & 31, px = o >> 5;
    const float v = t[c * 145 + px];
    const float h = tf32_rn(v);
    reinterpret_cast<float*>(hi)[zb * 4608 + o] = h;
    reinterpret_cast<float*>(lo)[zb * 4608 + o] = v - h;
  }
}

int p1_split_hwc(rbnn_net* net, const float* p1, int ZB, void* hi, void* lo, const float* f16_scale, cudaStream_t st) {
  p1_split_hwc_kernel<<<ZB, 256, 0, st>>>(p1, hi, lo, f16_scale);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// The input gradient is discontinuous in the conv2 pre-activations in two ways: LeakyReLU (the sign of the
// pre-activation) and MaxPool2d(2, stride 1) (which of the 4 window entries is the largest, model_nn.py:102-103).
// The tensor-core GEMM errs by ~5e-6 of the output maximum, so every A2 entry whose pre-activation is within
// guard = eps * max|pre| (of this image; the maximum comes from the GEMM epilogue, GemmDesc::group_max) of ZERO, or that
// is one of two or more entries of a pooling window within guard of that window's maximum (a near-tie for the arg-max),
// is recomputed exactly -- fp64 accumulation of the exact products over the 5x5x32 patch of P1 = hi + lo, CUDA cores --
// and rewritten in place.  Afterwards every sign and every window arg-max agrees with exact arithmetic.  Three phases:
//   1  flags from the untouched GEMM output (=> the recomputed set, hence the result, is deterministic)
//   2  re-evaluation, one warp per flagged entry; warps take 32-word chunks of the flag bitmap from a shared counter
//      (the ~100 flagged entries of a unit are unevenly spread), all 25 weight loads of an entry are in flight at once
//   3  fp32 ties: two entries of one window whose exact values differ but round to the same fp32 number
// One block per (sample, image); A2 is [zb][64][H].  Dynamic shared memory (refine_smem_bytes).
constexpr int kTieCap = 640;
static size_t refine_smem_bytes(int H) {
  const int hw = (H + 31) / 32;
  return (size_t)kTieCap * 8 + 2 * 4608 * 4 + (size_t)64 * hw * 4 + kTieCap * 4 + 800 * 2 + 16;
}

__global__ void __launch_bounds__(256, 4)
conv2_refine_kernel(float* __restrict__ a2, const float* __restrict__ p1, const float* __restrict__ p1lo,
                    const unsigned* __restrict__ unit_max, const float* __restrict__ bank, int64_t P, int64_t cw2,
                    int64_t cb2, int s0, int B, int H, float eps) {
  extern __shared__ __align__(16) unsigned char rsm[];
  const int hw = (H + 31) >> 5;                 // 32-channel words per position
  double* tie_v = reinterpret_cast<double*>(rsm);                        // [kTieCap] exact post-LeakyReLU values
  double* pd = tie_v + kTieCap;                                          // [4608] P1 = hi + lo as fp64 (converted once)
  unsigned* fl = reinterpret_cast<unsigned*>(pd + 4608);                 // [64 * hw] recompute flags, bit = channel
  int* tie_at = reinterpret_cast<int*>(fl + 64 * hw);                    // [kTieCap] position * H + channel
  int* ctr = reinterpret_cast<int*>(tie_at + kTieCap);                   // [0] tie count, [1] next chunk of phase 2
  const int zb = blockIdx.x, z = zb / B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* A = a2 + (int64_t)zb * 64 * H;
  for (int i = threadIdx.x; i < 1152; i += blockDim.x) {
    const float4 h = __ldg(reinterpret_cast<const float4*>(p1 + (int64_t)zb * 4608) + i);
    const float4 l = p1lo ? __ldg(reinterpret_cast<const float4*>(p1lo + (int64_t)zb * 4608) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<double2*>(pd)[2 * i] = make_double2((double)h.x + (double)l.x, (double)h.y + (double)l.y);
    reinterpret_cast<double2*>(pd)[2 * i + 1] = make_double2((double)h.z + (double)l.z, (double)h.w + (double)l.w);
  }
  // this lane's 25 filter elements k = lane + 32 i are the same for every entry: their offsets inside the 32x12x12 map,
  // two per register (the fp64 conversions of P1 and the table lookups were 40 % of the instructions of phase 2)
  unsigned toff[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    const int k0 = lane + 64 * i, k1 = lane + 64 * i + 32;
    const int c0 = k0 / 25, r0 = k0 - 25 * c0, c1 = k1 / 25, r1 = k1 - 25 * c1;
    const unsigned o0 = (unsigned)(c0 * 144 + (r0 / 5) * 12 + r0 % 5);
    const unsigned o1 = k1 < 800 ? (unsigned)(c1 * 144 + (r1 / 5) * 12 + r1 % 5) : 0u;
    toff[i] = o0 | (o1 << 16);
  }
  if (threadIdx.x < 2) ctr[threadIdx.x] = 0;
  const float guard = eps * __uint_as_float(__ldg(unit_max + zb));
  const float* __restrict__ wrow = bank + (int64_t)(s0 + z) * P;
  // phase 1.  A thread owns a channel and walks the 8 rows with rows y-1, y, y+1 (as pre-activations) in registers; per
  // window row it derives thr[wx] = (two or more of the window's entries lie within guard of its maximum) ? maximum -
  // guard : +inf, and an entry is flagged when it exceeds the threshold of any window it belongs to, or sits within
  // guard of zero.
  for (int h = threadIdx.x; h < hw * 32; h += blockDim.x) {
    const bool ok = h < H;
    const float* __restrict__ Ah = A + (ok ? h : 0);
    float rb[8], rc[8], thrA[7], thrB[7];
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const float v = Ah[x * H];
      rb[x] = v > 0.f ? v : 100.f * v;
    }
#pragma unroll
    for (int x = 0; x < 7; ++x) thrA[x] = __int_as_float(0x7f800000);
#pragma unroll 1
    for (int y = 0; y < 8; ++y) {
      if (y < 7) {
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const float v = Ah[((y + 1) * 8 + x) * H];
          rc[x] = v > 0.f ? v : 100.f * v;
        }
#pragma unroll
        for (int wx = 0; wx < 7; ++wx) {
          const float top = fmaxf(fmaxf(rb[wx], rb[wx + 1]), fmaxf(rc[wx], rc[wx + 1])) - guard;
          const int near = (rb[wx] > top) + (rb[wx + 1] > top) + (rc[wx] > top) + (rc[wx + 1] > top);
          thrB[wx] = near >= 2 ? top : __int_as_float(0x7f800000);
        }
      } else {
#pragma unroll
        for (int wx = 0; wx < 7; ++wx) thrB[wx] = __int_as_float(0x7f800000);
      }
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const float pre = rb[x];
        bool flag = fabsf(pre) < guard;
        if (x > 0) flag = flag || pre > thrA[x - 1] || pre > thrB[x - 1];
        if (x < 7) flag = flag || pre > thrA[x] || pre > thrB[x];
        const unsigned any = __ballot_sync(0xffffffffu, flag && ok);
        if (lane == 0) fl[(y * 8 + x) * hw + (h >> 5)] = any;
      }
#pragma unroll
      for (int x = 0; x < 8; ++x) rb[x] = rc[x];
#pragma unroll
      for (int wx = 0; wx < 7; ++wx) thrA[wx] = thrB[wx];
    }
  }
  __syncthreads();
  // phase 2
  const int nwords = 64 * hw, nchunks = (nwords + 31) >> 5;
  for (;;) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(&ctr[1], 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk >= nchunks) break;
    const int my_wi = chunk * 32 + lane;
    const unsigned my_bits = my_wi < nwords ? fl[my_wi] : 0u;
    unsigned live = __ballot_sync(0xffffffffu, my_bits != 0u);
    while (live) {
      const int wl = __ffs(live) - 1;
      live &= live - 1;
      unsigned any = __shfl_sync(0xffffffffu, my_bits, wl);
      const int wi = chunk * 32 + wl;
      const int pos = wi / hw, h0 = (wi - pos * hw) << 5;
      const int poff = (pos >> 3) * 12 + (pos & 7);
      while (any) {
        const int src = __ffs(any) - 1;
        any &= any - 1;
        const int hh = h0 + src;
        const float* __restrict__ w = wrow + cw2 + (int64_t)hh * 800 + lane;
        float wv[25];
#pragma unroll
        for (int i = 0; i < 25; ++i) wv[i] = __ldg(w + 32 * i);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 25; ++i) {
          const int t = poff + (int)((i & 1) ? (toff[i >> 1] >> 16) : (toff[i >> 1] & 0xffffu));
          s = fma(pd[t], (double)wv[i], s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s += (double)__ldg(wrow + cb2 + hh);
        s = s > 0.0 ? s : s * (double)kLeakySlope;
        if (lane == 0) {
          A[pos * H + hh] = (float)s;
          const int slot = atomicAdd(&ctr[0], 1);
          if (slot < kTieCap) { tie_at[slot] = pos * H + hh; tie_v[slot] = s; }
        }
      }
    }
  }
  // phase 3: two entries of one pooling window whose exact values differ but round to the SAME fp32 number would be
  // ordered by the pooling kernels' tie rule instead of by their values (7 windows of the 1.2e8 of the conv-512 cfg4
  // workload, each moving its unit's gradient by ~1e-3).  Both were re-evaluated above (a near-tie flags every entry
  // within the band of the window maximum), so the exact values are at hand: the larger one is raised by one fp32 ulp
  // per window-mate it has to beat.  The order of the list (atomic slots) does not enter the result.
  __syncthreads();
  const int nt = min(ctr[0], kTieCap);
  for (int i = threadIdx.x; i < nt; i += blockDim.x) {
    const int at = tie_at[i], pos = at / H, hh = at - pos * H, y = pos >> 3, x = pos & 7;
    const double v = tie_v[i];
    const float f = (float)v;
    int bump = 0;
    for (int j = 0; j < nt; ++j) {
      const int d = tie_at[j] - at;              // same channel: d = (pos2 - pos) * H
      if (d == 0 || d % H) continue;
      const int pos2 = pos + d / H;
      const int dy = (pos2 >> 3) - y, dx = (pos2 & 7) - x;
      if (dy < -1 || dy > 1 || dx < -1 || dx > 1) continue;
      bump += ((float)tie_v[j] == f && tie_v[j] < v) ? 1 : 0;
    }
    if (bump) {
      float r = f;
      for (int q = 0; q < bump; ++q) r = nextafterf(r, __int_as_float(0x7f800000));
      A[at] = r;
    }
  }
}

int conv2_refine(rbnn_net* net, float* a2, const float* p1, int s0, int Z, int B, float eps, cudaStream_t st,
                 const float* p1lo, const unsigned* unit_max) {
  RBNN_CHECK(unit_max != nullptr, "conv2_refine needs the per-image maxima of the GEMM epilogue");
  const size_t smem = refine_smem_bytes(net->H);
  static bool attr_done = false;
  if (!attr_done) {
    RBNN_CUDA(cudaFuncSetAttribute(conv2_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_done = true;
  }
  conv2_refine_kernel<<<Z * B, 256, smem, st>>>(a2, p1, p1lo, unit_max, net->bank, net->L.P, net->L.cw2, net->L.cb2, s0, B,
                                                net->H, eps);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rbnn
