// Convolutional architecture (model_nn.py:98-106), per posterior sample z and input b:
//   x[1,28,28] -Conv2d(1,32,5)-> [32,24,24] -LeakyReLU-> -MaxPool2d(2)-> P1[32,12,12]
//     -Conv2d(32,H,5)-> [H,8,8] -LeakyReLU-> A2 -MaxPool2d(2,stride=1)-> P2[H,7,7] -Flatten-> Linear(49H,C)
// conv1 (K=25, <1% of the FLOPs) and both pools are direct kernels; conv2 (97% of the FLOPs) is
// lowered to GEMM through an explicit im2col (round 1; implicit-GEMM on tcgen05 is the next step).
// Internal activation layouts after conv2 are position-major / channel-minor (HWC) so the GEMM
// output is consumed as is; model.7.weight is permuted once per bank row to match (sampler.cu).
// Max-pool ties route to the FIRST maximum in window scan order, as torch's max_pool2d does.
#include "common.cuh"

namespace rbnn {

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : v * kLeakySlope; }

// ---- conv1 + leaky + maxpool(2): one block per (z, b) -------------------------------------
__global__ void __launch_bounds__(256)
conv1_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ bank, int64_t P, int64_t cw1,
                      int64_t cb1, int s0, int B, float* __restrict__ p1, uint8_t* __restrict__ idx1) {
  __shared__ float xs[28 * 28];
  __shared__ float ws[32 * 25];
  __shared__ float bs[32];
  const int zb = blockIdx.x;
  const int z = zb / B, b = zb % B;
  const float* row = bank + (int64_t)(s0 + z) * P;
  for (int i = threadIdx.x; i < 784; i += blockDim.x) xs[i] = __ldg(x + (int64_t)b * 784 + i);
  for (int i = threadIdx.x; i < 800; i += blockDim.x) ws[i] = __ldg(row + cw1 + i);
  if (threadIdx.x < 32) bs[threadIdx.x] = __ldg(row + cb1 + threadIdx.x);
  __syncthreads();
  for (int o = threadIdx.x; o < 32 * 144; o += blockDim.x) {
    const int c = o / 144, py = (o % 144) / 12, px = o % 12;
    float patch[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) patch[i][j] = xs[(2 * py + i) * 28 + 2 * px + j];
    float best = 0.f;
    int bi = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) acc = fmaf(patch[dy + ky][dx + kx], ws[c * 25 + ky * 5 + kx], acc);
        const float v = leaky(acc + bs[c]);
        if ((dy == 0 && dx == 0) || v > best) { best = v; bi = dy * 2 + dx; }
      }
    p1[(int64_t)zb * 4608 + o] = best;
    idx1[(int64_t)zb * 4608 + o] = (uint8_t)bi;
  }
}

int conv1_pool_fwd(rbnn_net* net, const float* x, const float* bank, int s0, int Z, int B, float* p1,
                   uint8_t* idx1, cudaStream_t st) {
  conv1_pool_fwd_kernel<<<Z * B, 256, 0, st>>>(x, bank, net->L.P, net->L.cw1, net->L.cb1, s0, B, p1, idx1);
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
__global__ void col2im_conv2_kernel(const float* __restrict__ dcol, const float* __restrict__ p1, int64_t total,
                                    float* __restrict__ g1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ix = (int)(i % 12), iy = (int)((i / 12) % 12), c = (int)((i / 144) % 32);
  const int64_t zb = i / 4608;
  float acc = 0.f;
#pragma unroll
  for (int ky = 0; ky < 5; ++ky) {
    const int oy = iy - ky;
    if (oy < 0 || oy > 7) continue;
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
      const int ox = ix - kx;
      if (ox < 0 || ox > 7) continue;
      acc += __ldg(dcol + (zb * 64 + oy * 8 + ox) * 800 + c * 25 + ky * 5 + kx);
    }
  }
  g1[i] = __ldg(p1 + i) > 0.f ? acc : acc * kLeakySlope;
}

int col2im_conv2(rbnn_net* net, const float* dcol, const float* p1, int ZB, float* g1, cudaStream_t st) {
  const int64_t total = (int64_t)ZB * 4608;
  col2im_conv2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dcol, p1, total, g1);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- maxpool1 backward + conv1 input gradient, summed over the chunk's samples ---------------
// dx[b][iy][ix] (+)= sum_z sum_c sum_{ky,kx} [idx1 routes (iy-ky, ix-kx)] G1[z][b][c][.][.] * cw1[z][c][ky][kx]
__global__ void __launch_bounds__(256)
conv1_bwd_sum_kernel(const float* __restrict__ g1, const uint8_t* __restrict__ idx1, const float* __restrict__ bank,
                     int64_t P, int64_t cw1, int s0, int Z, int B, float* __restrict__ dx, int accumulate) {
  __shared__ float gs[4608];
  __shared__ uint8_t is[4608];
  __shared__ float ws[800];
  const int b = blockIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int z = 0; z < Z; ++z) {
    const int64_t zb = (int64_t)z * B + b;
    __syncthreads();
    for (int i = threadIdx.x; i < 4608; i += blockDim.x) {
      gs[i] = __ldg(g1 + zb * 4608 + i);
      is[i] = idx1[zb * 4608 + i];
    }
    for (int i = threadIdx.x; i < 800; i += blockDim.x) ws[i] = __ldg(bank + (int64_t)(s0 + z) * P + cw1 + i);
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int pix = threadIdx.x + t * 256;
      if (pix >= 784) break;
      const int iy = pix / 28, ix = pix % 28;
      float a = 0.f;
      for (int ky = 0; ky < 5; ++ky) {
        const int oy = iy - ky;
        if (oy < 0 || oy > 23) continue;
        for (int kx = 0; kx < 5; ++kx) {
          const int ox = ix - kx;
          if (ox < 0 || ox > 23) continue;
          const int pp = (oy >> 1) * 12 + (ox >> 1);
          const int sub = ((oy & 1) << 1) | (ox & 1);
#pragma unroll 8
          for (int c = 0; c < 32; ++c)
            if (is[c * 144 + pp] == sub) a = fmaf(gs[c * 144 + pp], ws[c * 25 + ky * 5 + kx], a);
        }
      }
      acc[t] += a;
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int pix = threadIdx.x + t * 256;
    if (pix >= 784) break;
    const int64_t o = (int64_t)b * 784 + pix;
    dx[o] = accumulate ? dx[o] + acc[t] : acc[t];
  }
}

int conv1_bwd_sum(rbnn_net* net, const float* g1, const uint8_t* idx1, const float* bank, int s0, int Z, int B,
                  float* dx_sum, int accumulate, cudaStream_t st) {
  conv1_bwd_sum_kernel<<<B, 256, 0, st>>>(g1, idx1, bank, net->L.P, net->L.cw1, s0, Z, B, dx_sum, accumulate);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- tcgen05 route (tc_conv.cu): operand preparation and guard-band refinement of conv2 ---------------------
// P1[zb][c][y][x] (fp32, CHW) -> channels-last tf32-split copies hi/lo[zb][y][x][c]: a pixel's 32 channels are the
// 128-byte K-block row of the implicit GEMM.  One block per image, transposed through shared memory.
__global__ void __launch_bounds__(256)
p1_split_hwc_kernel(const float* __restrict__ p1, float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float t[32 * 145];
  const int64_t zb = blockIdx.x;
  for (int i = threadIdx.x; i < 4608; i += blockDim.x) t[(i / 144) * 145 + i % 144] = __ldg(p1 + zb * 4608 + i);
  __syncthreads();
  for (int o = threadIdx.x; o < 4608; o += blockDim.x) {
    const int c = o & 31, px = o >> 5;
    const float v = t[c * 145 + px];
    const float h = tf32_rn(v);
    hi[zb * 4608 + o] = h;
    lo[zb * 4608 + o] = v - h;
  }
}

int p1_split_hwc(rbnn_net* net, const float* p1, int ZB, float* hi, float* lo, cudaStream_t st) {
  p1_split_hwc_kernel<<<ZB, 256, 0, st>>>(p1, hi, lo);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// The input gradient is discontinuous in the conv2 pre-activations in two ways: LeakyReLU (the sign of the
// pre-activation) and MaxPool2d(2, stride 1) (which of the 4 window entries is the largest, model_nn.py:102-103).
// The tensor-core GEMM errs by ~5e-6 of the output maximum, so every A2 entry whose pre-activation is within
// eps * max|pre| (of this image) of ZERO or of one of its 8 spatial NEIGHBOURS in the same channel (the entries it
// shares a pooling window with) is recomputed exactly -- fp64 accumulation of the fp32 products over the 5x5x32 patch,
// CUDA cores -- and rewritten in place.  Afterwards every sign and every window arg-max agrees with exact arithmetic.
// Values written concurrently by other warps differ from the ones they replace by far less than the band, so the
// unsynchronised neighbour reads are harmless.  One block per (sample, image); A2 is [zb][64 positions][H].
__global__ void __launch_bounds__(256)
conv2_refine_kernel(float* __restrict__ a2, const float* __restrict__ p1, const float* __restrict__ bank, int64_t P,
                    int64_t cw2, int64_t cb2, int s0, int B, int H, float eps) {
  __shared__ float ps[4608];
  __shared__ float red[8];
  const int zb = blockIdx.x, z = zb / B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* A = a2 + (int64_t)zb * 64 * H;
  for (int i = threadIdx.x; i < 4608; i += blockDim.x) ps[i] = __ldg(p1 + (int64_t)zb * 4608 + i);
  float m = 0.f;
  for (int i = threadIdx.x; i < 64 * H; i += blockDim.x) {
    const float v = A[i];
    m = fmaxf(m, v > 0.f ? v : -100.f * v);        // |pre-activation| (LeakyReLU inverted)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  const float guard = eps * m;
  const float* __restrict__ wrow = bank + (int64_t)(s0 + z) * P;
  for (int pos = warp; pos < 64; pos += 8) {
    const int y = pos >> 3, x = pos & 7;
    for (int h0 = 0; h0 < H; h0 += 32) {
      const int h = h0 + lane;
      bool flag = false;
      if (h < H) {
        const float v = A[pos * H + h];
        const float pre = v > 0.f ? v : 100.f * v;
        flag = fabsf(pre) < guard;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if ((dy == 0 && dx == 0) || yy < 0 || yy > 7 || xx < 0 || xx > 7) continue;
            const float vn = A[(yy * 8 + xx) * H + h];
            const float pn = vn > 0.f ? vn : 100.f * vn;
            flag = flag || fabsf(pre - pn) < guard;
          }
      }
      unsigned any = __ballot_sync(0xffffffffu, flag);
      while (any) {
        const int src = __ffs(any) - 1;
        any &= any - 1;
        const int hh = h0 + src;
        const float* __restrict__ w = wrow + cw2 + (int64_t)hh * 800;
        double s = 0.0;
        for (int k = lane; k < 800; k += 32) {
          const int c = k / 25, r = k - 25 * c, ky = r / 5, kx = r - 5 * ky;
          s = fma((double)ps[c * 144 + (y + ky) * 12 + x + kx], (double)__ldg(w + k), s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s += (double)__ldg(wrow + cb2 + hh);
        float val = (float)s;
        val = val > 0.f ? val : val * kLeakySlope;
        if (lane == src) A[pos * H + hh] = val;
      }
    }
  }
}

int conv2_refine(rbnn_net* net, float* a2, const float* p1, int s0, int Z, int B, float eps, cudaStream_t st) {
  conv2_refine_kernel<<<Z * B, 256, 0, st>>>(a2, p1, net->bank, net->L.P, net->L.cw2, net->L.cb2, s0, B, net->H, eps);
  net->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rbnn
