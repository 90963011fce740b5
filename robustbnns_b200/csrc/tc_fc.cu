// The tcgen05 engine for the fully connected nets (arch fc / fc2, model_nn.py:77-91): per chunk of
// posterior samples
//   forward   H1_z = leaky(X . W1_z^T + b1_z) [, H2_z = leaky(H1_z . W2_z^T + b2_z)]       tc::gemm, per-z
//   head      logits_z = H_z . Wo_z^T + bo_z -> softmax -> loss head -> dlogits            fc_head_kernel
//             dH_z = (dlogits_z . Wo_z) (.) leaky'(H_z)           (written pre-split for the next GEMM)
//   backward  [dH1_z = (dH2_z . W2_z) (.) leaky'(H1_z)]                                    tc::gemm, per-z
//             dX += sum_z dH1_z . W1_z             (K-concatenated over the samples, in TMEM) tc::gemm, reduce_z
// which is lossGradients.py:29-40 / adversarialAttacks.py:74-78 for all inputs and samples at once,
// input gradients only (no weight gradients).  Operands of the tensor-core GEMMs are kept K-major:
// the bank's weight matrices are re-laid once per refresh as [S,R,C] and transposed [S,C,R] copies,
// split into tf32 hi/lo pairs (RBNN_PREC_TF32X3), into power-of-two-scaled fp16 hi/lo pairs (RBNN_PREC_F16X3,
// arch fc) or rounded to bf16 (RBNN_PREC_BF16).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "philox.cuh"
#include "tc_gemm.cuh"

namespace rbnn {

namespace {

__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// ---- operand preparation --------------------------------------------------------------------------
// x[B][D] -> hi/lo (tf32 split) or bf16 copies with row pitch ld4 * 4 elements (rows start on 128-byte lines, see k_pitch)
__global__ void split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                             __nv_bfloat16* __restrict__ bf, int64_t n4, int d4, int ld4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const int64_t row = i / d4, o = row * ld4 + (i - row * d4);
    if (bf) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&a);
      pk.y = *reinterpret_cast<const uint32_t*>(&b);
      reinterpret_cast<uint2*>(bf)[o] = pk;
    } else {
      float4 h, l;
      h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
      l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
      reinterpret_cast<float4*>(hi)[o] = h;
      reinterpret_cast<float4*>(lo)[o] = l;
    }
  }
}

// bank row s, matrix [R, C] at `off`  ->  [s][R][C] and transposed [s][C][R] copies (32x32 smem tiles)
// grid: (C/32 ceil, R/32 ceil, count), block (32, 8)
// perm25: forward-copy column (c, tap) -> (tap, c) for the conv2 filters (C = 32 * 25), see TcMat::perm25
__global__ void relayout_kernel(const float* __restrict__ bank, int64_t P, int64_t off, int R, int C, int ld, int s0,
                                float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ thi,
                                float* __restrict__ tlo, __nv_bfloat16* __restrict__ bf,
                                __nv_bfloat16* __restrict__ tbf, int perm25) {
  __shared__ float tile[32][33];
  const int s = s0 + blockIdx.z;
  const float* __restrict__ src = bank + (int64_t)s * P + off;
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = blockIdx.y * 32 + i;
    float v = 0.f;
    if (r < R && c < C) {
      v = __ldg(src + (int64_t)r * C + c);
      const int cf = perm25 ? (c % 25) * 32 + c / 25 : c;
      const int64_t o = ((int64_t)s * R + r) * ld + cf;     // forward copy: row pitch ld >= C
      if (bf) {
        bf[o] = __float2bfloat16(v);
      } else {
        const float h = tf32_rn(v);
        hi[o] = h;
        lo[o] = v - h;
      }
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c2 = blockIdx.x * 32 + i;
    if (r2 < R && c2 < C) {
      const float v = tile[threadIdx.x][i];
      const int64_t o = ((int64_t)s * C + c2) * R + r2;
      if (tbf) {
        tbf[o] = __float2bfloat16(v);
      } else {
        const float h = tf32_rn(v);
        thi[o] = h;
        tlo[o] = v - h;
      }
    }
  }
}

// ---- F16X3 operand scaling ---------------------------------------------------------------------------
// fp16 keeps 11 significant bits over 2^-14 .. 2^16 only, so every F16X3 operand is multiplied by a power of two
// (exact) that puts its largest element just below 2^target: hi = rn_f16(s v) carries the leading 11 bits, lo =
// rn_f16(s v - hi) the next 11; elements far below the maximum lose relative (not absolute) precision, which a dot
// product does not see.  Operands whose maximum is known exactly when they are scaled (the inputs of a call; dH, which
// is bounded by 2 max|g| max|Wo|) go to 2^15, the top of the fp16 range -- 2^29 of dynamic range below the maximum
// before the lo part runs out of bits.  The W1 copies persist across calls, so their scale is fixed on first use and
// keeps 2^6 of head room (target 2^9) for rows uploaded later.  All scales live in device memory; nothing is read
// back by the host.
constexpr int kF16TargetExpFrozen = 9, kF16TargetExpExact = 15;

__device__ __forceinline__ float pow2_scale_for(float maxabs, int target_exp) {     // s = 2^k, s * maxabs in [2^(t-1), 2^t)
  if (!(maxabs > 0.f) || !isfinite(maxabs)) return 1.f;
  int e;
  frexpf(maxabs, &e);                                               // maxabs = m 2^e, m in [0.5, 1)
  return ldexpf(1.f, target_exp - e);
}

// *out_bits = max(*out_bits, max_i |v_i|) as float bits (non-negative floats order like unsigned integers);
// `count` segments of `len` floats, segment i at base + i * stride
__global__ void __launch_bounds__(256)
maxabs_kernel(const float* __restrict__ base, int64_t stride, int64_t len, int count, unsigned* __restrict__ out_bits) {
  __shared__ float red[8];
  float m = 0.f;
  const int64_t total = len * count;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t seg = i / len, o = i - seg * len;
    m = fmaxf(m, fabsf(__ldg(base + seg * stride + o)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    if (!(m == m)) m = __int_as_float(0x7f800000);                  // NaN weights -> "infinite" range -> overflow flag
    atomicMax(out_bits, __float_as_uint(m));
  }
}

// contiguous variant (activations of a chunk): 16-byte loads, no per-element segment arithmetic
__global__ void __launch_bounds__(256)
maxabs_flat_kernel(const float* __restrict__ p, int64_t n, unsigned* __restrict__ out_bits) {
  __shared__ float red[8];
  float m = 0.f;
  const int64_t n4 = (reinterpret_cast<uintptr_t>(p) & 15) ? 0 : n / 4;
  const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(p4 + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(p + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    if (!(m == m)) m = __int_as_float(0x7f800000);
    atomicMax(out_bits, __float_as_uint(m));
  }
}

// Inputs on the 8-bit pixel grid.  Every image data set the reference loads is uint8 / 255 (utils.py:102-103, 129-130,
// 190-191), i.e. x = fl(q / 255) with an integer |q| <= 2047.  Scaled by 255 * 2^j such an input IS its fp16 hi part and
// the lo part of the F16X3 split is zero, so the forward GEMM needs two passes (X_hi W_hi + X_hi W_lo) instead of three.
// Detection is exact, not a tolerance on the data: an element qualifies only when fl(255 x) is within two fp32 roundings
// (2^-22 relative) of an integer, so reading it as q / 255 moves it by less than the rounding of the division that
// produced it; anything else (PGD iterates, synthetic floats) keeps the three-pass split.
// bits[0] = max|x| (as maxabs_flat_kernel), bits[2] |= 1 when some element is off the grid.
constexpr float kGridDen = 255.f;
constexpr float kGridRelTol = 2.4e-7f;       // 2^-22
constexpr float kGridMaxQ = 2047.f;          // 11 significant bits: exact in fp16
__device__ __forceinline__ bool off_grid(float v) {
  const float t = v * kGridDen, q = rintf(t);
  return !(fabsf(t - q) <= fabsf(q) * kGridRelTol && fabsf(q) <= kGridMaxQ);
}
__global__ void __launch_bounds__(256)
maxabs_grid_kernel(const float* __restrict__ p, int64_t n, unsigned* __restrict__ bits) {
  __shared__ float red[8];
  __shared__ int red_off[8];
  float m = 0.f;
  bool off = false;
  const int64_t n4 = (reinterpret_cast<uintptr_t>(p) & 15) ? 0 : n / 4;
  const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(p4 + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    off = off || off_grid(v.x) || off_grid(v.y) || off_grid(v.z) || off_grid(v.w);
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(p + i);
    m = fmaxf(m, fabsf(v));
    off = off || off_grid(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const unsigned any_off = __ballot_sync(0xffffffffu, off);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = m; red_off[threadIdx.x >> 5] = any_off != 0u; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int o = red_off[0];
    for (int i = 1; i < 8; ++i) { m = fmaxf(m, red[i]); o |= red_off[i]; }
    if (!(m == m)) { m = __int_as_float(0x7f800000); o = 1; }
    atomicMax(bits, __float_as_uint(m));
    if (o) atomicOr(bits + 2, 1u);
  }
}

// after the maxima of the dirty rows are in: fix s_w1 on first use, afterwards only check that the frozen scale
// still keeps every |W1| inside the fp16 range (2^6 head room); otherwise raise the sticky host-visible flag
__global__ void freeze_scales_kernel(TcScales* sc, int* overflow) {
  const float mw = __uint_as_float(sc->maxw1_bits);
  if (!sc->frozen) {
    sc->s_w1 = pow2_scale_for(mw, kF16TargetExpFrozen);
    sc->frozen = 1;
  }
  if (!(mw * sc->s_w1 < 32768.f)) {
    *overflow = 1;
    __threadfence_system();
  }
}

// per call: [0] s_x, [1] 1 / (s_x s_w1) (forward unscale), [2] dH scale, [3] 1 / (dH scale * s_w1) (backward unscale)
// |dH| <= 2 max|g| max|Wo| with max|g| <= 1 for the built-in heads (g = softmax(.) - e_y) and max|d_pbar| for UPSTREAM
// dh_factor: the bound of the hidden-layer gradient is dh_factor * max|g| * max|Wo| (2 for the FC nets; the conv net
// passes 8: a position collects up to 4 pooling windows)
// xgrid != nullptr (the 4 words of an FC call: [0] max|x|, [1] max|g|, [2] off-grid flag of maxabs_grid_kernel): the
// inputs' scale becomes 255 * 2^j when they lie on the pixel grid, and xgrid[3] = 1 tells split_f16_kernel and the fused
// kernel that the lo part of the inputs is zero (two-pass forward).
__global__ void call_scales_kernel(const TcScales* sc, const unsigned* xmax_bits, const unsigned* gmax_bits,
                                   float* out, float dh_factor, unsigned* xgrid = nullptr) {
  float sx = pow2_scale_for(__uint_as_float(*xmax_bits), kF16TargetExpExact);
  if (xgrid) {
    const float xmax = __uint_as_float(*xmax_bits);
    const bool on = xgrid[2] == 0u && xmax > 0.f && isfinite(xmax);
    if (on) sx = kGridDen * pow2_scale_for(rintf(xmax * kGridDen), kF16TargetExpExact);   // q_max 2^j in [2^14, 2^15)
    xgrid[3] = on ? 1u : 0u;
  }
  const float gmax = gmax_bits ? __uint_as_float(*gmax_bits) : 1.f;
  const float sd = pow2_scale_for(dh_factor * gmax * __uint_as_float(sc->maxwo_bits), kF16TargetExpExact);
  out[0] = sx;
  out[1] = 1.f / (sx * sc->s_w1);
  out[2] = sd;
  out[3] = 1.f / (sd * sc->s_w1);
}

__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// x[n] -> fp16 hi/lo of s_x * x
// xgrid[3] != 0 (inputs on the pixel grid, s = 255 * 2^j): hi = q 2^j exactly, lo = 0
__global__ void split_f16_kernel(const float* __restrict__ x, const float* __restrict__ call_sc,
                                 __half* __restrict__ hi, __half* __restrict__ lo, int64_t n4, int d4, int ld4,
                                 const unsigned* __restrict__ xgrid) {
  const float s = __ldg(call_sc);
  const bool grid = xgrid && __ldg(xgrid + 3) != 0u;
  const float s2 = s / kGridDen;                      // 2^j (exact)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const int64_t row = i / d4, o = row * ld4 + (i - row * d4);
    __half h[4], l[4];
    if (grid) {
      h[0] = __float2half_rn(rintf(v.x * kGridDen) * s2); h[1] = __float2half_rn(rintf(v.y * kGridDen) * s2);
      h[2] = __float2half_rn(rintf(v.z * kGridDen) * s2); h[3] = __float2half_rn(rintf(v.w * kGridDen) * s2);
      l[0] = l[1] = l[2] = l[3] = __float2half_rn(0.f);
    } else {
      split_f16(v.x * s, h[0], l[0]); split_f16(v.y * s, h[1], l[1]);
      split_f16(v.z * s, h[2], l[2]); split_f16(v.w * s, h[3], l[3]);
    }
    reinterpret_cast<uint2*>(hi)[o] = *reinterpret_cast<const uint2*>(h);
    reinterpret_cast<uint2*>(lo)[o] = *reinterpret_cast<const uint2*>(l);
  }
}

// F16X3 twin of relayout_kernel: [s][R][C] and transposed [s][C][R] fp16 hi/lo copies of s_w1 * W
__global__ void relayout_f16_kernel(const float* __restrict__ bank, int64_t P, int64_t off, int R, int C, int ld, int s0,
                                    const TcScales* __restrict__ sc, __half* __restrict__ hi, __half* __restrict__ lo,
                                    __half* __restrict__ thi, __half* __restrict__ tlo, int perm25) {
  __shared__ float tile[32][33];
  const float sw = sc->s_w1;
  const int s = s0 + blockIdx.z;
  const float* __restrict__ src = bank + (int64_t)s * P + off;
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = blockIdx.y * 32 + i;
    float v = 0.f;
    if (r < R && c < C) {
      v = __ldg(src + (int64_t)r * C + c) * sw;
      // conv2 filters: tap-major forward copy (TcMat::perm25); 2 = rows of 5 x 6 taps, the 6th a zero padding tap
      // (K = 960: 128-byte K-blocks of two taps; the padding slots were zeroed when the copies were allocated)
      const int tap = c % 25;
      const int cf = perm25 == 2 ? ((tap / 5) * 6 + tap % 5) * 32 + c / 25 : (perm25 ? tap * 32 + c / 25 : c);
      const int64_t o = ((int64_t)s * R + r) * ld + cf;
      split_f16(v, hi[o], lo[o]);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c2 = blockIdx.x * 32 + i;
    if (r2 < R && c2 < C) {
      const int64_t o = ((int64_t)s * C + c2) * R + r2;
      split_f16(tile[threadIdx.x][i], thi[o], tlo[o]);
    }
  }
}

// K-sample + re-layout in one pass (F16X3, scale already frozen): block = 32 rows of W1 x all C columns, looped in
// 64-column tiles.  Per tile a thread draws the 8 normals of two Philox calls (8 consecutive columns of one row: the
// bank's element order), writes w = loc + sigma * eps to the bank row, splits s_w1 * w into fp16 hi/lo and stores the 8
// values of the forward copy with ONE 16-byte store per array straight from registers; the (hi | lo) pairs go through a
// shared-memory tile from which every thread takes 8 consecutive ROWS of one column for the transposed copy (16-byte
// stores again, 4 lanes = the 64 contiguous bytes of a column's 32 rows).  Row norms (guard band) and max|W1| (range
// check) come from the same registers.  (The first version moved one value per store through 2-byte stores: 123
// instructions per weight, issue-bound at 2.7 TB/s; this one needs ~45.)  grid: (ceil(R/32), count), block 256; C % 8 == 0,
// R % 8 == 0, ld % 8 == 0.
// Thread maps: draw phase  -- warp w, lane l: row (w/2)*8 + l/4, column group (w%2)*4 + l%4  (conflict-free tile writes);
//              transposed  -- warp w, lane l: column w*8 + l/4, rows (l%4)*8 .. +7           (conflict-free tile reads).
__global__ void __launch_bounds__(256)
sample_relayout_f16_kernel(const float* __restrict__ loc, const float* __restrict__ sigma, float* __restrict__ bank,
                           int64_t P, int64_t off, int R, int C, int ld, int s0, int64_t sample_index0, int64_t stride,
                           uint32_t k0, uint32_t k1, TcScales* __restrict__ sc, __half* __restrict__ hi,
                           __half* __restrict__ lo, __half* __restrict__ thi, __half* __restrict__ tlo,
                           float* __restrict__ wnorm, const int64_t* __restrict__ index_offset) {
  __shared__ uint32_t tile[64][33];                      // [column][row]: fp16 hi | lo << 16
  __shared__ float red_n[32][2], red_m[8];
  const float sw = sc->s_w1;
  const int s = s0 + blockIdx.y;
  const uint32_t g = (uint32_t)(sample_index0 + (index_offset ? *index_offset : 0) + (int64_t)blockIdx.y * stride);
  const int r0 = blockIdx.x * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tr = (warp >> 1) * 8 + (lane >> 2), tg = (warp & 1) * 4 + (lane & 3);
  const int tc = warp * 8 + (lane >> 2), trg = (lane & 3) * 8;       // transposed phase: column of the tile, first row
  const int r = r0 + tr;
  float* __restrict__ brow = bank + (int64_t)s * P;
  float racc = 0.f, mabs = 0.f;
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int c = c0 + tg * 8;
    uint32_t hp[4] = {0u, 0u, 0u, 0u}, lp[4] = {0u, 0u, 0u, 0u};     // packed fp16 pairs of s_w1 * w: hi and residual
    if (r < R && c < C) {
      const int64_t i = off + (int64_t)r * C + c;        // multiple of 8
      // the guide's parameters first: their L2 latency hides behind the ~300 instructions of the two Philox calls
      float4 sg0, sg1, lc0, lc1;
      asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sg0.x), "=f"(sg0.y), "=f"(sg0.z), "=f"(sg0.w) : "l"(sigma + i));
      asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sg1.x), "=f"(sg1.y), "=f"(sg1.z), "=f"(sg1.w) : "l"(sigma + i + 4));
      asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(lc0.x), "=f"(lc0.y), "=f"(lc0.z), "=f"(lc0.w) : "l"(loc + i));
      asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(lc1.x), "=f"(lc1.y), "=f"(lc1.z), "=f"(lc1.w) : "l"(loc + i + 4));
      float z[8];
      philox_normals4((uint32_t)(i >> 2), g, k0, k1, *reinterpret_cast<float(*)[4]>(&z[0]));
      philox_normals4((uint32_t)(i >> 2) + 1u, g, k0, k1, *reinterpret_cast<float(*)[4]>(&z[4]));
      float w[8];
      w[0] = fmaf(sg0.x, z[0], lc0.x); w[1] = fmaf(sg0.y, z[1], lc0.y);
      w[2] = fmaf(sg0.z, z[2], lc0.z); w[3] = fmaf(sg0.w, z[3], lc0.w);
      w[4] = fmaf(sg1.x, z[4], lc1.x); w[5] = fmaf(sg1.y, z[5], lc1.y);
      w[6] = fmaf(sg1.z, z[6], lc1.z); w[7] = fmaf(sg1.w, z[7], lc1.w);
      float2* out = reinterpret_cast<float2*>(brow + i);             // bank rows are 8-byte aligned (P even)
#pragma unroll
      for (int j = 0; j < 4; ++j) out[j] = make_float2(w[2 * j], w[2 * j + 1]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { racc = fmaf(w[j], w[j], racc); mabs = fmaxf(mabs, fabsf(w[j])); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = w[2 * j] * sw, b = w[2 * j + 1] * sw;
        const __half2 h = __floats2half2_rn(a, b);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
        hp[j] = *reinterpret_cast<const uint32_t*>(&h);
        lp[j] = *reinterpret_cast<const uint32_t*>(&l);
      }
      const int64_t o = ((int64_t)s * R + r) * ld + c;               // 16-byte aligned: ld % 8 == 0, c % 8 == 0
      *reinterpret_cast<uint4*>(hi + o) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      *reinterpret_cast<uint4*>(lo + o) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {                                    // (hi | lo << 16) of columns tg*8 + 2j, + 2j + 1
      tile[tg * 8 + 2 * j][tr] = __byte_perm(hp[j], lp[j], 0x5410);
      tile[tg * 8 + 2 * j + 1][tr] = __byte_perm(hp[j], lp[j], 0x7632);
    }
    __syncthreads();
    const int c2 = c0 + tc, r2 = r0 + trg;
    if (c2 < C && r2 < R) {
      uint32_t v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = tile[tc][trg + k];
      uint4 th, tl;
      th.x = __byte_perm(v[0], v[1], 0x5410); tl.x = __byte_perm(v[0], v[1], 0x7632);
      th.y = __byte_perm(v[2], v[3], 0x5410); tl.y = __byte_perm(v[2], v[3], 0x7632);
      th.z = __byte_perm(v[4], v[5], 0x5410); tl.z = __byte_perm(v[4], v[5], 0x7632);
      th.w = __byte_perm(v[6], v[7], 0x5410); tl.w = __byte_perm(v[6], v[7], 0x7632);
      const int64_t o = ((int64_t)s * C + c2) * R + r2;              // 16-byte aligned: R % 8 == 0, r2 % 8 == 0
      *reinterpret_cast<uint4*>(thi + o) = th;
      *reinterpret_cast<uint4*>(tlo + o) = tl;
    }
    __syncthreads();
  }
  // row norm: a row's 8 column groups sit in 4 lanes of two neighbouring warps; then the maximum over the block's 32 rows
  racc += __shfl_xor_sync(0xffffffffu, racc, 1);
  racc += __shfl_xor_sync(0xffffffffu, racc, 2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mabs = fmaxf(mabs, __shfl_xor_sync(0xffffffffu, mabs, o));
  if ((lane & 3) == 0) red_n[tr][warp & 1] = racc;
  if (lane == 0) red_m[warp] = mabs;
  __syncthreads();
  if (tid == 0) {
    float m = 0.f, a = 0.f;
    for (int i = 0; i < 32; ++i) m = fmaxf(m, red_n[i][0] + red_n[i][1]);
    for (int i = 0; i < 8; ++i) a = fmaxf(a, red_m[i]);
    const bool bad = !(m == m) || !(a == a);
    atomicMax(reinterpret_cast<unsigned*>(wnorm + s), __float_as_uint(bad ? __int_as_float(0x7f800000) : sqrtf(m)));
    atomicMax(&sc->maxw1_bits, __float_as_uint(bad ? __int_as_float(0x7f800000) : a));
  }
}

// ---- head -------------------------------------------------------------------------------------------
constexpr int kHeadWarps = 8;
constexpr int kHeadRowsPerWarp = 8;
constexpr int kMaxC = 32;

template <int C_MAX>
__device__ __forceinline__ void softmax_c(float (&v)[C_MAX], int C) {
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

// One warp per (sample z, input b) row of the top hidden activations.
//   GRAD = false : logits[z][b][:] = H . Wo_z^T + bo_z
//   GRAD = true  : additionally the loss head (same algebra as head.cu::dlogits_kernel) and
//                  dH[z][b][j] = leaky'(H[z][b][j]) * sum_c dlogits_c Wo_z[c][j], written as tf32 hi/lo or bf16.
// grid: (ceil(B / 64), Z); block 256; dynamic smem: (C*H + C) floats (Wo_z, bo_z of this block's sample)
template <int C_MAX, bool GRAD>
__global__ void __launch_bounds__(kHeadWarps * 32)
fc_head_kernel(int head, const float* __restrict__ Hact, const float* __restrict__ bank, int64_t P, int64_t wo_off,
               int64_t bo_off, int z_row0, const int32_t* __restrict__ labels, const float* __restrict__ pbar,
               int B, int H, int C, float* __restrict__ logits_out, float* __restrict__ dh_hi,
               float* __restrict__ dh_lo, __nv_bfloat16* __restrict__ dh_bf) {
  extern __shared__ float sm[];
  float* wo = sm;               // [C][H]
  float* bo = sm + C * H;       // [C]
  const int z = blockIdx.y;
  const float* __restrict__ row_w = bank + (int64_t)(z_row0 + z) * P;
  for (int i = threadIdx.x; i < C * H; i += blockDim.x) wo[i] = __ldg(row_w + wo_off + i);
  for (int i = threadIdx.x; i < C; i += blockDim.x) bo[i] = __ldg(row_w + bo_off + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b_begin = (blockIdx.x * kHeadWarps + warp) * kHeadRowsPerWarp;
  for (int b = b_begin; b < min(b_begin + kHeadRowsPerWarp, B); ++b) {
    const int64_t row = (int64_t)z * B + b;
    const float* __restrict__ h = Hact + row * H;
    float acc[C_MAX];
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) acc[c] = 0.f;
    for (int j = lane * 4; j < H; j += 128) {
      const float4 hv = __ldg(reinterpret_cast<const float4*>(h + j));
#pragma unroll
      for (int c = 0; c < C_MAX; ++c)
        if (c < C) {
          const float4 w = *reinterpret_cast<const float4*>(wo + c * H + j);
          acc[c] = fmaf(hv.x, w.x, acc[c]);
          acc[c] = fmaf(hv.y, w.y, acc[c]);
          acc[c] = fmaf(hv.z, w.z, acc[c]);
          acc[c] = fmaf(hv.w, w.w, acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        acc[c] += bo[c];
      }
    if (!GRAD) {
      if (lane < C) {
        float v = 0.f;
#pragma unroll
        for (int c = 0; c < C_MAX; ++c)
          if (c == lane) v = acc[c];
        logits_out[row * C + lane] = v;
      }
      continue;
    }
    // ---- loss head: dlogits = softmax-Jacobian applied to g (see head.cu) ----
    const int y = labels[b];
    float g[C_MAX];
    if (head != RBNN_HEAD_LOGITS_UPSTREAM) softmax_c<C_MAX>(acc, C);                 // acc = p = softmax(z)
    if (head == RBNN_HEAD_LOGITS_UPSTREAM) {   // loss of the mean LOGITS (ensembles, deterministic nets): dlogits = d_pbar
#pragma unroll
      for (int c = 0; c < C_MAX; ++c) acc[c] = (c < C) ? __ldg(pbar + (int64_t)b * C + c) : 0.f;
    } else if (head == RBNN_HEAD_LOGITS_CE) {
#pragma unroll
      for (int c = 0; c < C_MAX; ++c) acc[c] = acc[c] - (c == y ? 1.f : 0.f);
    } else {
      if (head == RBNN_HEAD_MEAN_OF_GRADS) {
#pragma unroll
        for (int c = 0; c < C_MAX; ++c) g[c] = acc[c];
      } else {
#pragma unroll
        for (int c = 0; c < C_MAX; ++c) g[c] = (c < C) ? __ldg(pbar + (int64_t)b * C + c) : 0.f;
      }
      if (head != RBNN_HEAD_UPSTREAM) {
        softmax_c<C_MAX>(g, C);
#pragma unroll
        for (int c = 0; c < C_MAX; ++c) g[c] -= (c == y ? 1.f : 0.f);
      }
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < C_MAX; ++c)
        if (c < C) dot = fmaf(acc[c], g[c], dot);
#pragma unroll
      for (int c = 0; c < C_MAX; ++c) acc[c] = acc[c] * (g[c] - dot);
    }
    // ---- dH = (dlogits . Wo) (.) leaky'(H) ----
    for (int j = lane * 4; j < H; j += 128) {
      const float4 hv = __ldg(reinterpret_cast<const float4*>(h + j));
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < C_MAX; ++c)
        if (c < C) {
          const float4 w = *reinterpret_cast<const float4*>(wo + c * H + j);
          d[0] = fmaf(acc[c], w.x, d[0]);
          d[1] = fmaf(acc[c], w.y, d[1]);
          d[2] = fmaf(acc[c], w.z, d[2]);
          d[3] = fmaf(acc[c], w.w, d[3]);
        }
      d[0] = hv.x > 0.f ? d[0] : d[0] * kLeakySlope;
      d[1] = hv.y > 0.f ? d[1] : d[1] * kLeakySlope;
      d[2] = hv.z > 0.f ? d[2] : d[2] * kLeakySlope;
      d[3] = hv.w > 0.f ? d[3] : d[3] * kLeakySlope;
      if (dh_bf) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(d[0], d[1]), bb = __floats2bfloat162_rn(d[2], d[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&a);
        pk.y = *reinterpret_cast<const uint32_t*>(&bb);
        *reinterpret_cast<uint2*>(dh_bf + row * H + j) = pk;
      } else {
        float4 hi4, lo4;
        hi4.x = tf32_rn(d[0]); hi4.y = tf32_rn(d[1]); hi4.z = tf32_rn(d[2]); hi4.w = tf32_rn(d[3]);
        lo4.x = d[0] - hi4.x; lo4.y = d[1] - hi4.y; lo4.z = d[2] - hi4.z; lo4.w = d[3] - hi4.w;
        *reinterpret_cast<float4*>(dh_hi + row * H + j) = hi4;
        *reinterpret_cast<float4*>(dh_lo + row * H + j) = lo4;
      }
    }
  }
}

// Guard-band refinement of a forward GEMM output (TF32X3 mode).  LeakyReLU makes the input gradient
// discontinuous in the pre-activations: a hidden unit whose pre-activation is within the tensor-core
// rounding error of zero may come out with the wrong sign and change its row of the gradient by O(1/H).
// Every unit with |pre-activation| < eps * (row max) is therefore recomputed exactly (fp64 accumulation of
// the fp32 products, CUDA cores) and rewritten in place.  About 1e-3 of the units qualify.
// One warp per (sample z, input b) row; H_hi (+ H_lo when the output is stored tf32-split) hold leaky(pre).
// NCH = ceil(H / 128): the row (4 values per lane and 128-column chunk) stays in registers between the maximum pass
// and the flag pass, so H is read once.
// pre-activation of second-layer unit j of fc2 from the inputs, fp64 throughout (one warp; every lane returns it):
// b2 + sum_i W2[j,i] leaky(b1[i] + <x, W1[i,:]>).  Slow path of the second worklist (it overflowed).
__device__ float exact_unit2_warp(const float* __restrict__ xr, const float* __restrict__ wrow, int D, int H, int64_t w1_off,
                                  int64_t b1_off, const float* __restrict__ w2row, float b2, int lane) {
  double out = 0.0;
  for (int i = 0; i < H; ++i) {
    const float* __restrict__ wr = wrow + w1_off + (int64_t)i * D;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) s = fma((double)__ldg(xr + d), (double)__ldg(wr + d), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s += (double)__ldg(wrow + b1_off + i);
    out = fma(s > 0.0 ? s : s * (double)kLeakySlope, (double)__ldg(w2row + i), out);
  }
  return (float)(out + (double)b2);
}

//
// Second layer of fc2 (wl2 != nullptr): the A operand of the re-evaluation is the FIRST-layer activation as the tensor
// cores left it (~5e-6 of the layer maximum off, except for its own refined units), so the "exact" value still carries
// that error times ||W2_j||: a second-layer unit within eps2 * (row max) of zero is therefore queued on a second
// worklist as (z, b, j) and settled by refine2_kernel from EXACT first-layer values (fp64 all the way from x).
template <int NCH>
__global__ void __launch_bounds__(256)
refine_kernel(float* __restrict__ Hhi, float* __restrict__ Hlo, int Z, int B, int H, const float* __restrict__ a_hi,
              const float* __restrict__ a_lo, int64_t a_zstride, int K, const float* __restrict__ bank, int64_t P,
              int64_t w_off, int64_t b_off, int z_row0, float eps, unsigned long long* __restrict__ wl2,
              unsigned* __restrict__ wl2_count, unsigned wl2_cap, float eps2, const float* __restrict__ x0, int D0,
              int64_t w1_off, int64_t b1_off) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < (int64_t)Z * B; row += nwarps) {
    const int z = (int)(row / B), b = (int)(row % B);
    float* hh = Hhi + row * H;
    float* hl = Hlo ? Hlo + row * H : nullptr;
    float v[NCH][4];
    float m = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = c * 128 + lane * 4;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < H) {
        t = *reinterpret_cast<const float4*>(hh + j);
        if (hl) {
          const float4 l = *reinterpret_cast<const float4*>(hl + j);
          t.x += l.x; t.y += l.y; t.z += l.z; t.w += l.w;
        }
      }
      v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w;
#pragma unroll
      for (int e = 0; e < 4; ++e) m = fmaxf(m, v[c][e] > 0.f ? v[c][e] : -100.f * v[c][e]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float guard = eps * m;
    const float* __restrict__ ah = a_hi + (int64_t)z * a_zstride + (int64_t)b * K;
    const float* __restrict__ al = a_lo ? a_lo + (int64_t)z * a_zstride + (int64_t)b * K : nullptr;
    const float* __restrict__ wrow = bank + (int64_t)(z_row0 + z) * P;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j0 = c * 128, j = j0 + lane * 4;
      unsigned flags = 0u;
      if (j < H) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pre = v[c][e] > 0.f ? v[c][e] : -100.f * v[c][e];
          if (pre < guard) flags |= 1u << e;
        }
      }
      unsigned any = __ballot_sync(0xffffffffu, flags != 0u);
      bool changed = false;
      while (any) {
        const int src = __ffs(any) - 1;
        const unsigned f = __shfl_sync(0xffffffffu, flags, src);
        const int e = __ffs(f) - 1;
        const int jj = j0 + src * 4 + e;
        const float* __restrict__ w = wrow + w_off + (int64_t)jj * K;
        // 8 loads in flight per lane and two accumulation chains: the loop is latency bound (weights come from L2)
        double s = 0.0, s2 = 0.0;
        int d = lane;
        for (; d + 224 < K; d += 256) {
          float a[8], wv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { a[u] = ah[d + 32 * u]; wv[u] = __ldg(w + d + 32 * u); }
          if (al) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] += al[d + 32 * u];
          }
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
            s = fma((double)a[u], (double)wv[u], s);
            s2 = fma((double)a[u + 1], (double)wv[u + 1], s2);
          }
        }
        for (; d < K; d += 32) {
          float a = ah[d];
          if (al) a += al[d];
          s = fma((double)a, (double)__ldg(w + d), s);
        }
        s += s2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s += (double)__ldg(wrow + b_off + jj);
        float val = (float)s;
        if (wl2 && fabsf(val) < eps2 * m) {                     // warp-uniform: every lane holds the reduced sum
          unsigned slot = 0;
          if (lane == 0) slot = atomicAdd(wl2_count + z, 1u);   // list of sample z; order varies from run to run, the set of entries does not
          slot = __shfl_sync(0xffffffffu, slot, 0);
          if (slot < wl2_cap) {
            if (lane == 0)
              wl2[(size_t)z * wl2_cap + slot] =
                  ((unsigned long long)z << 44) | ((unsigned long long)b << 20) | ((unsigned long long)jj << 4);
          } else {                                              // worklist full (pathological inputs): settle it here
            val = exact_unit2_warp(x0 + (int64_t)b * D0, wrow, D0, H, w1_off, b1_off, w + 0, __ldg(wrow + b_off + jj), lane);
          }
        }
        val = val > 0.f ? val : val * kLeakySlope;
        if (lane == src) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q == e) v[c][q] = val;
          flags &= ~(1u << e);
          changed = true;
        }
        any = __ballot_sync(0xffffffffu, flags != 0u);
      }
      if (changed) {
        if (hl) {
          float4 h4, l4;
          h4.x = tf32_rn(v[c][0]); h4.y = tf32_rn(v[c][1]); h4.z = tf32_rn(v[c][2]); h4.w = tf32_rn(v[c][3]);
          l4.x = v[c][0] - h4.x; l4.y = v[c][1] - h4.y; l4.z = v[c][2] - h4.z; l4.w = v[c][3] - h4.w;
          *reinterpret_cast<float4*>(hh + j) = h4;
          *reinterpret_cast<float4*>(hl + j) = l4;
        } else {
          *reinterpret_cast<float4*>(hh + j) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
        }
      }
    }
  }
}

// Second-level refinement of fc2's second layer: every queued unit (z, b, j) is re-evaluated from the inputs in fp64,
//   H1[i] = leaky(b1[i] + <x_b, W1_z[i,:]>) for ALL hidden i (kept in double), pre2 = b2[j] + <H1, W2_z[j,:]>,
// and H2[z, b, j] = leaky(pre2) is rewritten.  The worklist is kept PER SAMPLE (refine_kernel appends to list z), so a
// block takes kR2Group entries of one sample and every W1_z row it loads serves all of them -- with one global list
// the entries of ~10 samples interleaved, almost every entry paid for its own pass over W1_z (1.6 MB at fc2-512) and the
// kernel took 21.6 ms of a 24 ms gradient evaluation (fc2-512, 1000 inputs x 100 samples).  The first-layer part is a
// small fp64 GEMM [entries x D] . [D x H]: a warp takes 4 rows of W1_z at a time, so a shared-memory load of an input
// feeds 4 DFMAs (one per row left the shared-memory pipe 4x oversubscribed), and the 32 partial sums of a lane are
// reduced over the warp by halving (31 exchanges instead of 160).
constexpr int kR2Group = 8, kR2Rows = 4;
__global__ void __launch_bounds__(256)
refine2_kernel(const unsigned long long* __restrict__ wl2, const unsigned* __restrict__ counts, unsigned cap,
               const float* __restrict__ x, int B, int D, int H, const float* __restrict__ bank, int64_t P, int64_t w1_off,
               int64_t b1_off, int64_t w2_off, int64_t b2_off, int z_row0, float* __restrict__ H2) {
  extern __shared__ double r2sm[];
  double* h1 = r2sm;                                  // [kR2Group][H] exact first-layer activations
  double* xs = h1 + (size_t)kR2Group * H;             // [kR2Group / 2][D] pairs: inputs of entries (2 r2, 2 r2 + 1), zeros past m
  __shared__ unsigned long long ent[kR2Group];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int z = blockIdx.y;
  const unsigned n = min(counts[z], cap);
  const unsigned long long* __restrict__ list = wl2 + (size_t)z * cap;
  const float* __restrict__ wrow = bank + (int64_t)(z_row0 + z) * P;
  for (unsigned g0 = blockIdx.x * kR2Group; g0 < n; g0 += gridDim.x * kR2Group) {
    const int m = (int)min((unsigned)kR2Group, n - g0);
    __syncthreads();
    if (tid < kR2Group) ent[tid] = tid < m ? list[g0 + tid] : 0ull;
    __syncthreads();
    for (int idx = tid; idx < kR2Group * D; idx += 256) {
      const int r = idx / D, d = idx - r * D;
      const int b = (int)((ent[r] >> 20) & 0xFFFFFF);
      xs[((size_t)(r >> 1) * D + d) * 2 + (r & 1)] = r < m ? (double)__ldg(x + (int64_t)b * D + d) : 0.0;
    }
    __syncthreads();
    for (int i0 = warp * kR2Rows; i0 < H; i0 += 8 * kR2Rows) {
      double acc[kR2Rows * kR2Group];
#pragma unroll
      for (int q = 0; q < kR2Rows * kR2Group; ++q) acc[q] = 0.0;
      const float* __restrict__ wr = wrow + w1_off + (int64_t)i0 * D;
#pragma unroll 2
      for (int d = lane; d < D; d += 32) {
        double wv[kR2Rows];
#pragma unroll
        for (int rr = 0; rr < kR2Rows; ++rr) wv[rr] = i0 + rr < H ? (double)__ldg(wr + (int64_t)rr * D + d) : 0.0;
        const double2* xp = reinterpret_cast<const double2*>(xs) + d;      // entry pair r2 of input element d: xp[r2 * D]
#pragma unroll
        for (int r2 = 0; r2 < kR2Group / 2; ++r2) {
          const double2 xv = xp[(size_t)r2 * D];
#pragma unroll
          for (int rr = 0; rr < kR2Rows; ++rr) {
            acc[rr * kR2Group + 2 * r2] = fma(xv.x, wv[rr], acc[rr * kR2Group + 2 * r2]);
            acc[rr * kR2Group + 2 * r2 + 1] = fma(xv.y, wv[rr], acc[rr * kR2Group + 2 * r2 + 1]);
          }
        }
      }
      // reduction by halving: after the round with stride s a lane keeps the half of its values selected by (lane & s);
      // lane l ends up with the warp-wide sum of value l = (row l / 8, entry l % 8)
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        const bool up = (lane & sft) != 0;
#pragma unroll
        for (int q = 0; q < sft; ++q) {
          const double keep = up ? acc[q + sft] : acc[q];
          const double send = up ? acc[q] : acc[q + sft];
          acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
        }
      }
      const int rr = lane >> 3, r = lane & 7, i = i0 + rr;
      if (i < H && r < m) {
        const double v = acc[0] + (double)__ldg(wrow + b1_off + i);
        h1[r * H + i] = v > 0.0 ? v : v * (double)kLeakySlope;
      }
    }
    __syncthreads();
    for (int r = warp; r < m; r += 8) {
      const unsigned long long e = ent[r];
      const int b = (int)((e >> 20) & 0xFFFFFF), j = (int)((e >> 4) & 0xFFFF);
      const float* __restrict__ w2r = wrow + w2_off + (int64_t)j * H;
      double sacc = 0.0;
      for (int i = lane; i < H; i += 32) sacc = fma(h1[r * H + i], (double)__ldg(w2r + i), sacc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      if (lane == 0) {
        const double pre = sacc + (double)__ldg(wrow + b2_off + j);
        const float v = (float)pre;
        H2[((int64_t)z * B + b) * H + j] = pre > 0.0 ? v : (float)(pre * (double)kLeakySlope);
      }
    }
  }
}

// xnorm[b] = ||x_b||_2 : one warp per input row
__global__ void xnorm_kernel(const float* __restrict__ x, int B, int D, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (b >= B) return;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = __ldg(x + (int64_t)b * D + d); s = fmaf(v, v, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[b] = sqrtf(s);
}

// wnorm[s] = max_j ||W_s[j, :]||_2 over the R rows of the [R, C] matrix at `off` of bank row s.  One warp per matrix
// row, 8 rows per block, grid (ceil(R / 8), samples); the per-sample maximum is an atomicMax on the float bits
// (non-negative floats order like unsigned integers), so out[s0 .. s0 + samples) must be zeroed first.
// max_bits != nullptr (F16X3): also *max_bits = max(*max_bits, max |W_s|) -- the operand range of the fp16 copies,
// from the same pass over the weights.
__global__ void __launch_bounds__(256)
wnorm_kernel(const float* __restrict__ bank, int64_t P, int64_t off, int R, int C, int s0, float* __restrict__ out,
             unsigned* __restrict__ max_bits) {
  __shared__ float red[8], redm[8];
  const int s = s0 + blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  float acc = 0.f, mabs = 0.f;
  if (r < R) {
    const float* __restrict__ w = bank + (int64_t)s * P + off + (int64_t)r * C;
    for (int c = lane; c < C; c += 32) { const float v = __ldg(w + c); acc = fmaf(v, v, acc); mabs = fmaxf(mabs, fabsf(v)); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    mabs = fmaxf(mabs, __shfl_xor_sync(0xffffffffu, mabs, o));
  }
  if (lane == 0) { red[warp] = acc; redm[warp] = mabs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f, a = 0.f;
    for (int i = 0; i < 8; ++i) { m = fmaxf(m, red[i]); a = fmaxf(a, redm[i]); }
    const bool bad = !(m == m) || !(a == a);                       // NaN weights
    atomicMax(reinterpret_cast<unsigned*>(out + s), __float_as_uint(bad ? __int_as_float(0x7f800000) : sqrtf(m)));
    if (max_bits) atomicMax(max_bits, __float_as_uint(bad ? __int_as_float(0x7f800000) : a));   // -> overflow flag
  }
}

// out[i] = (accumulate ? out[i] : 0) + sum_p partial[p][i]  (fixed order => deterministic)
__global__ void reduce_slots_kernel(const float* __restrict__ partial, int nparts, int64_t n4,
                                    float* __restrict__ out, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 s = accumulate ? reinterpret_cast<const float4*>(out)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = 0; p < nparts; ++p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(partial) + (int64_t)p * n4 + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  reinterpret_cast<float4*>(out)[i] = s;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// F16X3 operand-range helpers for the conv engine (tc_conv.cu): *bits = max(*bits, max|p|) and the call's 4 scalars
int tc_maxabs(rbnn_net* n, const float* p, int64_t count, unsigned* bits, cudaStream_t st) {
  const unsigned blocks = (unsigned)std::min<int64_t>((count / 4 + 255) / 256 + 1, (int64_t)n->sm_count * 8);
  maxabs_flat_kernel<<<blocks, 256, 0, st>>>(p, count, bits);
  n->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

int tc_call_scales(rbnn_net* n, const unsigned* xmax_bits, const unsigned* gmax_bits, float dh_factor, float* out,
                   cudaStream_t st) {
  call_scales_kernel<<<1, 1, 0, st>>>(n->tc.scales, xmax_bits, gmax_bits, out, dh_factor);
  n->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

int tc_sample_relayout_f16(rbnn_net* n, const float* d_loc, const float* d_rho, uint64_t seed, int64_t sample_index0,
                           int64_t stride, int s0, int count, cudaStream_t st, int* done, const int64_t* d_index_offset) {
  *done = 0;
  TcBank& tc = n->tc;
  const TcMat& m = tc.mat[0];
  if (n->prec != RBNN_PREC_F16X3 || n->arch != RBNN_ARCH_FC || tc.mode != n->prec || tc.capacity < s0 + count ||
      tc.capacity != n->capacity || !tc.frozen_host || tc.nmat != 1 || m.off != 0 || (m.C & 7) || (m.R & 7) || (m.ld & 7) || (n->L.P & 1) ||
      (reinterpret_cast<uintptr_t>(d_loc) & 15) || count <= 0)
    return 0;
  if (tc.overflow_host && *(volatile int*)tc.overflow_host) return 0;      // let tc_bank_refresh report / reset it
  RBNN_TRY(sample_sigma(n, d_rho, st));
  RBNN_CUDA(cudaMemsetAsync(tc.wnorm + s0, 0, (size_t)count * sizeof(float), st));
  dim3 grid((m.R + 31) / 32, count);
  sample_relayout_f16_kernel<<<grid, 256, 0, st>>>(
      d_loc, n->sigma, n->bank, n->L.P, m.off, m.R, m.C, m.ld, s0, sample_index0, stride, (uint32_t)(seed & 0xFFFFFFFFu),
      (uint32_t)(seed >> 32), tc.scales, reinterpret_cast<__half*>(m.h_hi), reinterpret_cast<__half*>(m.h_lo),
      reinterpret_cast<__half*>(m.th_hi), reinterpret_cast<__half*>(m.th_lo), tc.wnorm, d_index_offset);
  n->launches++;
  RBNN_CUDA(cudaGetLastError());
  // the rest of the rows (b1, Wo, bo) with the plain sampler, then the Wo range and the scale check
  RBNN_TRY(sample_diag_from(n, d_loc, seed, sample_index0, stride, s0, count, (int64_t)m.R * m.C, st, d_index_offset));
  maxabs_kernel<<<n->sm_count, 256, 0, st>>>(n->bank + (int64_t)s0 * n->L.P + n->L.wo, n->L.P, (int64_t)n->C * n->H, count,
                                             &tc.scales->maxwo_bits);
  freeze_scales_kernel<<<1, 1, 0, st>>>(tc.scales, tc.overflow_dev);
  n->launches += 2;
  RBNN_CUDA(cudaGetLastError());
  std::fill(tc.dirty + s0, tc.dirty + s0 + count, (uint8_t)0);
  *done = 1;
  return 0;
}

int tc_supported(const rbnn_net* n) {
  if (n->arch != RBNN_ARCH_FC && n->arch != RBNN_ARCH_FC2) return 0;
  if (n->H < 32 || n->H > 2048 || n->C > kMaxC) return 0;      // refine_kernel keeps a row of <= 2048 units in registers
  // D % 8 != 0 (half moons: D = 2): only fc2, whose bulk is the H x H layer -- the first layer and its input gradient
  // (K = D / N = D: a TMA row would be 8 bytes) run on the CUDA-core GEMM, the middle GEMMs on tcgen05 (TF32X3)
  if ((n->D & 7) && !(n->arch == RBNN_ARCH_FC2 && (n->H & 7) == 0)) return 0;
  return n->cc_major == 10;
}

// D % 8 != 0: the first layer cannot be a TMA operand (see tc_supported)
static bool small_d(const rbnn_net* n) { return (n->D & 7) != 0; }

// Row pitch (elements) of a K-major operand copy with K elements per row: rows start on 128-byte lines, so that a
// 128-byte TMA box row is one L2 line.  (784 fp16 = 1568 B rows straddle two lines per box row and halve the TMA
// rate -- measured: the first-layer forward ran its operand pipeline at half the rate of the backward one.)
static int k_pitch(const rbnn_net* n, int K) {
  const int per_line = n->prec == RBNN_PREC_TF32X3 ? 32 : 64;
  return (K + per_line - 1) / per_line * per_line;
}

// F16X3 rides on the fused forward+head kernel (dH is produced and scaled inside it): arch fc only
int tc_f16x3_supported(const rbnn_net* n) {
  return tc_supported(n) && n->arch == RBNN_ARCH_FC && n->L.w1 == 0 && tc::fused_supported(n->H, n->C);
}

void tc_bank_free(rbnn_net* n) {
  for (int i = 0; i < 2; ++i) {
    TcMat& m = n->tc.mat[i];
    cudaFree(m.hi); cudaFree(m.lo); cudaFree(m.thi); cudaFree(m.tlo); cudaFree(m.bf); cudaFree(m.tbf);
    cudaFree(m.h_hi); cudaFree(m.h_lo); cudaFree(m.th_hi); cudaFree(m.th_lo);
    m = TcMat();
  }
  cudaFree(n->tc.wnorm);
  n->tc.wnorm = nullptr;
  cudaFree(n->tc.scales);
  n->tc.scales = nullptr;
  cudaFree(n->tc.xgrid_last);
  n->tc.xgrid_last = nullptr;
  if (n->tc.overflow_host) cudaFreeHost(n->tc.overflow_host);
  n->tc.overflow_host = n->tc.overflow_dev = nullptr;
  delete[] n->tc.dirty;
  n->tc.dirty = nullptr;
  n->tc.capacity = 0;
  n->tc.mode = -1;
  n->tc.frozen_host = 0;
}

// A new posterior is being installed (other guide parameters, another uploaded bank): the F16X3 weight scale that was
// frozen for the old weights is forgotten and every derived copy is marked stale, so the next use fixes a scale from
// the rows it actually touches.
int tc_bank_invalidate(rbnn_net* n) {
  TcBank& tc = n->tc;
  n->keep.valid = 0;
  if (tc.dirty) std::fill(tc.dirty, tc.dirty + tc.capacity, (uint8_t)1);
  if (tc.scales) {
    RBNN_CUDA(cudaDeviceSynchronize());
    RBNN_CUDA(cudaMemset(tc.scales, 0, sizeof(TcScales)));
    if (tc.overflow_host) *tc.overflow_host = 0;
  }
  tc.frozen_host = 0;
  return 0;
}

static int tc_bank_refresh_once(rbnn_net* n, int s0, int s1, cudaStream_t st, bool* relaid);

// Bring the derived copies of bank rows [s0, s1) up to date (all rows after a capacity / precision change).
// F16X3: the power-of-two weight scale is frozen from the first rows it sees (2^6 of head room).  Whenever this call
// re-lays rows it waits for the range check and, should a row exceed the fp16 range under the frozen scale, forgets
// the scale and re-lays the rows of this call under a new one -- inside the same call, transparently.  (Rows drawn by
// the fused sampler path between refreshes are covered by the sticky flag read at the start of the next call.)
int tc_bank_refresh(rbnn_net* n, int s0, int s1, cudaStream_t st) {
  TcBank& tc = n->tc;
  const bool f16 = n->prec == RBNN_PREC_F16X3;
  for (int attempt = 0; attempt < 3; ++attempt) {
    bool relaid = false;
    RBNN_TRY(tc_bank_refresh_once(n, s0, s1, st, &relaid));
    if (!f16 || !relaid || !tc.overflow_host) return 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap != cudaStreamCaptureStatusNone) return 0;          // no host reads inside a graph capture
    RBNN_CUDA(cudaStreamSynchronize(st));
    if (!*(volatile int*)tc.overflow_host) return 0;
    RBNN_TRY(tc_bank_invalidate(n));                            // new scale from the rows of this call
  }
  set_error("F16X3: bank rows do not fit the fp16 operand range (non-finite weights?)");
  return 1;
}

static int tc_bank_refresh_once(rbnn_net* n, int s0, int s1, cudaStream_t st, bool* relaid) {
  TcBank& tc = n->tc;
  const bool bf = n->prec == RBNN_PREC_BF16, f16 = n->prec == RBNN_PREC_F16X3;
  *relaid = false;
  if (f16 && tc.overflow_host && *(volatile int*)tc.overflow_host) {
    // rows drawn by the fused sampler path after s_w1 was frozen did not fit the fp16 range (the guide changed without
    // rbnn_bank_invalidate): results since then are invalid.  Start over (new scale) and tell the caller.
    RBNN_TRY(tc_bank_invalidate(n));
    set_error("F16X3: posterior samples drawn after the operand scale was fixed exceed the fp16 range (> 64x the earlier "
              "maximum |w|); the results of calls since then are invalid.  Call rbnn_bank_invalidate when the guide "
              "parameters change.  The scale has been reset: repeat the call");
    return 1;
  }
  if (tc.capacity < n->capacity || tc.mode != n->prec) {
    RBNN_CUDA(cudaDeviceSynchronize());
    tc_bank_free(n);
    tc.nmat = n->arch == RBNN_ARCH_FC2 ? 2 : 1;
    tc.mat[0].off = n->L.w1; tc.mat[0].R = n->H; tc.mat[0].C = n->D; tc.mat[0].ld = k_pitch(n, n->D);
    if (n->arch == RBNN_ARCH_CONV) {      // the conv2 filters [H][32*5*5]; conv1 and the output layer stay on CUDA cores
      tc.mat[0].off = n->L.cw2; tc.mat[0].C = 800; tc.mat[0].ld = k_pitch(n, 800); tc.mat[0].perm25 = 1;
      if (f16) { tc.mat[0].ld = 960; tc.mat[0].perm25 = 2; }       // two taps per 128-byte K-block, 5 x 6 taps per filter
    }
    if (tc.nmat == 2) { tc.mat[1].off = n->L.w2; tc.mat[1].R = n->H; tc.mat[1].C = n->H; tc.mat[1].ld = n->H; }
    for (int i = 0; i < tc.nmat; ++i) {
      TcMat& m = tc.mat[i];
      const size_t elems = (size_t)n->capacity * m.R * m.ld, telems = (size_t)n->capacity * m.R * m.C;
      if (bf) {
        RBNN_CUDA(cudaMalloc(&m.bf, elems * 2));
        RBNN_CUDA(cudaMalloc(&m.tbf, telems * 2));
      } else if (f16) {
        RBNN_CUDA(cudaMalloc(&m.h_hi, elems * 2));
        RBNN_CUDA(cudaMalloc(&m.h_lo, elems * 2));
        if (m.perm25 == 2) {                                          // zero padding taps
          RBNN_CUDA(cudaMemset(m.h_hi, 0, elems * 2));
          RBNN_CUDA(cudaMemset(m.h_lo, 0, elems * 2));
        }
        RBNN_CUDA(cudaMalloc(&m.th_hi, telems * 2));
        RBNN_CUDA(cudaMalloc(&m.th_lo, telems * 2));
      } else {
        RBNN_CUDA(cudaMalloc(&m.hi, elems * 4));
        RBNN_CUDA(cudaMalloc(&m.lo, elems * 4));
        RBNN_CUDA(cudaMalloc(&m.thi, telems * 4));
        RBNN_CUDA(cudaMalloc(&m.tlo, telems * 4));
      }
    }
    RBNN_CUDA(cudaMalloc(&tc.wnorm, (size_t)n->capacity * sizeof(float)));
    if (f16) {
      RBNN_CUDA(cudaMalloc(&tc.scales, sizeof(TcScales)));
      RBNN_CUDA(cudaMemset(tc.scales, 0, sizeof(TcScales)));
      RBNN_CUDA(cudaMalloc(&tc.xgrid_last, sizeof(unsigned)));
      RBNN_CUDA(cudaMemset(tc.xgrid_last, 0, sizeof(unsigned)));
      RBNN_CUDA(cudaHostAlloc(&tc.overflow_host, sizeof(int), cudaHostAllocMapped));
      *tc.overflow_host = 0;
      RBNN_CUDA(cudaHostGetDevicePointer(&tc.overflow_dev, tc.overflow_host, 0));
    }
    tc.dirty = new uint8_t[n->capacity];
    std::fill(tc.dirty, tc.dirty + n->capacity, (uint8_t)1);
    tc.capacity = n->capacity;
    tc.mode = n->prec;
    n->alloc_epoch++;
  }
  if (f16) {
    // operand ranges of the dirty rows first (W1 fixes / checks the frozen scale, Wo bounds dH)
    bool any = false;
    for (int s = s0; s < s1;) {
      if (!tc.dirty[s]) { ++s; continue; }
      int e = s;
      while (e < s1 && tc.dirty[e]) ++e;
      const int64_t P = n->L.P;
      RBNN_CUDA(cudaMemsetAsync(tc.wnorm + s, 0, (size_t)(e - s) * sizeof(float), st));
      wnorm_kernel<<<dim3((tc.mat[0].R + 7) / 8, e - s), 256, 0, st>>>(n->bank, P, tc.mat[0].off, tc.mat[0].R, tc.mat[0].C,
                                                                       s, tc.wnorm, &tc.scales->maxw1_bits);
      const int64_t wo_len = (int64_t)n->C * n->H * (n->arch == RBNN_ARCH_CONV ? 49 : 1);
      maxabs_kernel<<<n->sm_count, 256, 0, st>>>(n->bank + (int64_t)s * P + n->L.wo, P, wo_len, e - s,
                                                 &tc.scales->maxwo_bits);
      n->launches += 2;
      RBNN_CUDA(cudaGetLastError());
      any = true;
      s = e;
    }
    if (any) {
      freeze_scales_kernel<<<1, 1, 0, st>>>(tc.scales, tc.overflow_dev);
      n->launches++;
      RBNN_CUDA(cudaGetLastError());
      tc.frozen_host = 1;
    }
  }
  int s = s0;
  while (s < s1) {
    if (!tc.dirty[s]) { ++s; continue; }
    int e = s;
    while (e < s1 && tc.dirty[e]) ++e;
    for (int i = 0; i < tc.nmat; ++i) {
      TcMat& m = tc.mat[i];
      dim3 grid((m.C + 31) / 32, (m.R + 31) / 32, e - s);
      if (f16)
        relayout_f16_kernel<<<grid, dim3(32, 8), 0, st>>>(n->bank, n->L.P, m.off, m.R, m.C, m.ld, s, tc.scales,
                                                          reinterpret_cast<__half*>(m.h_hi), reinterpret_cast<__half*>(m.h_lo),
                                                          reinterpret_cast<__half*>(m.th_hi), reinterpret_cast<__half*>(m.th_lo), m.perm25);
      else
        relayout_kernel<<<grid, dim3(32, 8), 0, st>>>(n->bank, n->L.P, m.off, m.R, m.C, m.ld, s, m.hi, m.lo, m.thi, m.tlo,
                                                      reinterpret_cast<__nv_bfloat16*>(m.bf),
                                                      reinterpret_cast<__nv_bfloat16*>(m.tbf), m.perm25);
      n->launches++;
      RBNN_CUDA(cudaGetLastError());
    }
    if (!f16) {        // F16X3: the row norms came with the operand-range pass above
      RBNN_CUDA(cudaMemsetAsync(tc.wnorm + s, 0, (size_t)(e - s) * sizeof(float), st));
      wnorm_kernel<<<dim3((tc.mat[0].R + 7) / 8, e - s), 256, 0, st>>>(n->bank, n->L.P, tc.mat[0].off, tc.mat[0].R,
                                                                       tc.mat[0].C, s, tc.wnorm, nullptr);
      n->launches++;
      RBNN_CUDA(cudaGetLastError());
    }
    std::fill(tc.dirty + s, tc.dirty + e, (uint8_t)0);
    *relaid = true;
    s = e;
  }
  return 0;
}

static int run_gemm(rbnn_net* n, tc::GemmDesc& d, int tag, cudaStream_t st) {
  d.mode = n->prec == RBNN_PREC_BF16 ? tc::MODE_BF16 : (n->prec == RBNN_PREC_F16X3 ? tc::MODE_F16X3 : tc::MODE_TF32X3);
  d.sm_count = n->sm_count;
  // RBNN_FC_PAIR (experiments): 1 = per-sample GEMMs on CTA pairs, 2 = the sample-reduced input-gradient GEMM too
  static const int fc_pair = getenv("RBNN_FC_PAIR") ? atoi(getenv("RBNN_FC_PAIR")) : 0;
  if (fc_pair && !d.pair && (!d.reduce_z || fc_pair >= 2) && d.BN % 32 == 0) d.pair = 1;
  std::string err;
  if (tag) RBNN_TRY(timing_begin(n, tag, st));
  if (tc::gemm(d, st, &err)) {
    set_error("%s", err.c_str());
    return 1;
  }
  n->launches++;
  if (tag) RBNN_TRY(timing_end(n, tag, st));
  return 0;
}

static int pick_bn(int N) {
  // widest tile that wastes the least of the last column block (784 -> 4 x 208, 512 -> 2 x 256)
  int best = 16;
  double best_eff = 0.0;
  for (int bn = 256; bn >= 16; bn -= 16) {
    const int tiles = (N + bn - 1) / bn;
    const double eff = (double)N / ((double)tiles * bn) * (bn >= 128 ? 1.0 : 0.5 + bn / 256.0);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = bn; }
  }
  return best;
}

namespace {
struct FcWs {
  // per call
  float *x_hi = nullptr, *x_lo = nullptr;
  __nv_bfloat16* x_bf = nullptr;
  __half *x_h16 = nullptr, *x_l16 = nullptr;                  // F16X3: fp16 split of s_x * x
  float* call_sc = nullptr;                                   // F16X3: the call's 4 scale scalars (call_scales_kernel)
  unsigned* max_bits = nullptr;                               // F16X3: [0] max|x|, [1] max|d_pbar| (UPSTREAM head)
  __half *dtop_h16 = nullptr, *dtop_l16 = nullptr;            // F16X3: dH as scaled fp16 hi/lo
  // per chunk of Z samples ([Z, B, H] each)
  float *h1 = nullptr, *h1_lo = nullptr, *h2 = nullptr;       // tf32x3/fc2: h1 holds the hi part (sign == sign of H1)
  __nv_bfloat16 *h1_bf = nullptr;
  float *dtop_hi = nullptr, *dtop_lo = nullptr, *d1_hi = nullptr, *d1_lo = nullptr;
  __nv_bfloat16 *dtop_bf = nullptr, *d1_bf = nullptr;
  float *logits = nullptr, *partial = nullptr, *xnorm = nullptr;
  unsigned long long* worklist = nullptr;
  unsigned long long* wl2 = nullptr;       // fc2: second-layer units to settle from exact first-layer values
  unsigned* wl2_count = nullptr;
  float* dx_part = nullptr;                                   // D % 8 != 0: [Z][B][D] per-sample dX partials (dx_small_rows_kernel)
  unsigned wl2_cap = 0;
};
}  // namespace

// arch fc with a hidden layer the fused forward+head kernel covers: H never goes to HBM
static bool use_fused(const rbnn_net* n) {
  return n->arch == RBNN_ARCH_FC && n->L.w1 == 0 && tc::fused_supported(n->H, n->C) &&
         (!n->tc_unfused || n->prec == RBNN_PREC_F16X3);
}

static size_t fc_per_z_bytes(const rbnn_net* n, int B, bool grad) {
  const bool two = n->arch == RBNN_ARCH_FC2, bf = n->prec == RBNN_PREC_BF16, f16 = n->prec == RBNN_PREC_F16X3;
  const size_t bh = pad256((size_t)B * n->H * 4), bh2 = pad256((size_t)B * n->H * 2);
  if (use_fused(n)) {
    if (!grad) return pad256((size_t)B * n->C * 4);
    return (bf ? bh2 : (f16 ? 2 * bh2 : 2 * bh)) + pad256(tc::fused_worklist_slots(B, 1) * 8);
  }
  size_t per = bh;                                   // h1 (fp32 or hi)
  if (two) per += (bf ? bh2 : bh) + bh;              // h1 lo / bf16 + h2
  if (two && !bf) per += pad256(std::max<size_t>((size_t)B * n->H / 8, 2048) + 64);   // second worklist: one 8-byte slot per 64 hidden units (>= 256 slots) + its counter
  if (grad) {
    per += bf ? bh2 : 2 * bh;                        // dtop
    if (two) per += bf ? bh2 : 2 * bh;               // d1
  } else {
    per += pad256((size_t)B * n->C * 4);             // logits
  }
  return per;
}

// forward GEMMs of one chunk; returns the top hidden activations (fp32) in *top
// Guard band of the fused kernel: |pre-activation| < eps * ||x_b|| * max_j ||w_zj||.  Error model of the TF32x3 GEMM:
// 294 accumulation steps, each truncating <= 2^-23 of the running sum (|partial sums| <~ 0.1 ||x|| ||w||), plus
// 3 * 2^-22 * sum|x_d w_d| from the dropped lo.lo term and the truncated lo operands  =>  <~ 4e-6 ||x|| ||w||.
// Measured on B200: max error 5e-6 of the output max (~7e-7 ||x|| ||w||).  eps = 2^-16 = 1.5e-5 keeps a 4x margin
// over the model and ~20x over the measurement; ~4e-4 of the units qualify.
constexpr float kGuardEpsFused = 1.0f / 65536.0f;
// F16X3: half the accumulation steps of TF32x3 (K = 16 per MMA) and a smaller measured error (2.9e-6 vs 5e-6 of the
// output max on the same operands), hence half the band: 2^-17 keeps the same ~20x margin and halves the fix-up work.
constexpr float kGuardEpsFusedF16 = 1.0f / 131072.0f;
constexpr float kGuardEps = 1.0f / 4096.0f;   // ~50x the measured TF32x3 error bound (5e-6 of the output max)

// Narrow band of fc2's second layer (fraction of the row maximum): bounds the first-layer tensor-core error (measured
// <= 5e-6 of the layer maximum) propagated through one row of W2 -- ~6e-6 of the second layer's row maximum at 4 sigma for
// well-scaled weights; 2^-14 = 6.1e-5 keeps 10x over that.  ~1.5e-4 of the units qualify.
constexpr float kGuardEps2 = 1.0f / 16384.0f;

static int refine(rbnn_net* n, float* h_hi, float* h_lo, int Z, int B, const float* a_hi, const float* a_lo,
                  int64_t a_zstride, int K, int64_t w_off, int64_t b_off, int z0, cudaStream_t st,
                  unsigned long long* wl2 = nullptr, unsigned* wl2_count = nullptr, unsigned wl2_cap = 0,
                  const float* x0 = nullptr) {
  const int64_t rows = (int64_t)Z * B;
  const unsigned blocks = (unsigned)std::min<int64_t>((rows + 7) / 8, (int64_t)n->sm_count * 8);
#define RBNN_REFINE(NCH) refine_kernel<NCH><<<blocks, 256, 0, st>>>(h_hi, h_lo, Z, B, n->H, a_hi, a_lo, a_zstride, K, n->bank, \
                                                                  n->L.P, w_off, b_off, z0, kGuardEps, wl2, wl2_count,     \
                                                                  wl2_cap, kGuardEps2, x0, n->D, n->L.w1, n->L.b1)
  const int nch = (n->H + 127) / 128;
  RBNN_CHECK(nch <= 16, "tcgen05 engine (unfused route): hidden sizes up to 2048");
  if (nch <= 1) RBNN_REFINE(1); else if (nch <= 2) RBNN_REFINE(2); else if (nch <= 4) RBNN_REFINE(4);
  else if (nch <= 8) RBNN_REFINE(8); else RBNN_REFINE(16);
#undef RBNN_REFINE
  n->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

static int fc_forward_chunk_tc(rbnn_net* n, const FcWs& w, const float* x, int B, int z0, int Z, const float** top,
                               cudaStream_t st) {
  const int H = n->H, D = n->D;
  const int64_t P = n->L.P;
  const bool two = n->arch == RBNN_ARCH_FC2, bf = n->prec == RBNN_PREC_BF16;
  const TcMat& m1 = n->tc.mat[0];
  if (small_d(n)) {
    // first layer in fp32 on the CUDA cores (exactly the FP32 engine's), then H1 -> tf32 hi (in place) + residual for the
    // tcgen05 GEMM of the second layer; no guard band needed for it
    GemmArgs a{};
    a.A = x; a.lda = D; a.sAz = 0;
    a.B = n->bank + (int64_t)z0 * P + n->L.w1; a.ldb = D; a.sBz = P;
    a.bias = n->bank + (int64_t)z0 * P + n->L.b1; a.sbz = P;
    a.C = w.h1; a.ldc = H; a.sCz = (int64_t)B * H;
    a.M = B; a.N = H; a.K = D; a.Z = Z; a.epi = EPI_BIAS_LEAKY; a.act = RBNN_ACT_LEAKY;
    a.tag = 1;
    RBNN_TRY(gemm_simt(n, a, st));
    const int64_t n4 = (int64_t)Z * B * H / 4;
    split_kernel<<<(unsigned)std::min<int64_t>((n4 + 255) / 256, 148 * 16), 256, 0, st>>>(w.h1, w.h1, w.h1_lo, nullptr, n4, H / 4, H / 4);
    n->launches++;
    RBNN_CUDA(cudaGetLastError());
  }
  tc::GemmDesc g;
  g.M = B; g.N = H; g.K = D; g.Z = Z; g.BN = pick_bn(H);
  g.A.hi = bf ? (const void*)w.x_bf : (const void*)w.x_hi; g.A.lo = w.x_lo; g.A.rows = B; g.A.ld = m1.ld; g.A.zstride = 0;
  if (bf) g.B.hi = reinterpret_cast<const __nv_bfloat16*>(m1.bf) + (int64_t)z0 * H * m1.ld;
  else { g.B.hi = m1.hi + (int64_t)z0 * H * m1.ld; g.B.lo = m1.lo + (int64_t)z0 * H * m1.ld; }
  g.B.rows = H; g.B.ld = m1.ld; g.B.zstride = (int64_t)H * m1.ld;
  g.epi = tc::EPI_BIAS_LEAKY;
  g.bias = n->bank + (int64_t)z0 * P + n->L.b1; g.bias_zstride = P;
  g.out = w.h1; g.out_ld = H; g.out_zstride = (int64_t)B * H;
  if (two) {
    if (bf) g.out_bf = w.h1_bf; else g.out_lo = w.h1_lo;
  }
  if (!small_d(n)) {
    RBNN_TRY(run_gemm(n, g, 1, st));
    if (!bf) RBNN_TRY(refine(n, w.h1, two ? w.h1_lo : nullptr, Z, B, x, nullptr, 0, D, n->L.w1, n->L.b1, z0, st));
  }
  *top = w.h1;
  if (two) {
    const TcMat& m2 = n->tc.mat[1];
    tc::GemmDesc q;
    q.M = B; q.N = H; q.K = H; q.Z = Z; q.BN = pick_bn(H);
    q.A.hi = bf ? (const void*)w.h1_bf : (const void*)w.h1; q.A.lo = w.h1_lo; q.A.rows = B; q.A.ld = H;
    q.A.zstride = (int64_t)B * H;
    if (bf) q.B.hi = reinterpret_cast<const __nv_bfloat16*>(m2.bf) + (int64_t)z0 * H * H;
    else { q.B.hi = m2.hi + (int64_t)z0 * H * H; q.B.lo = m2.lo + (int64_t)z0 * H * H; }
    q.B.rows = H; q.B.ld = H; q.B.zstride = (int64_t)H * H;
    q.epi = tc::EPI_BIAS_LEAKY;
    q.bias = n->bank + (int64_t)z0 * P + n->L.b2; q.bias_zstride = P;
    q.out = w.h2; q.out_ld = H; q.out_zstride = (int64_t)B * H;
    RBNN_TRY(run_gemm(n, q, 0, st));
    if (!bf) {
      // guard band of the second layer: wide band from the stored first-layer activations, narrow band (their propagated
      // tensor-core error) from exact first-layer values (refine2_kernel)
      RBNN_CUDA(cudaMemsetAsync(w.wl2_count, 0, (size_t)Z * sizeof(unsigned), st));
      RBNN_TRY(refine(n, w.h2, nullptr, Z, B, w.h1, w.h1_lo, (int64_t)B * H, H, n->L.w2, n->L.b2, z0, st, w.wl2,
                      w.wl2_count, w.wl2_cap, x));
      const int smem2 = kR2Group * (H + D) * (int)sizeof(double);
      static int smem2_set = 0;
      if (smem2 > 48 * 1024 && smem2_set < smem2) {
        RBNN_CUDA(cudaFuncSetAttribute(refine2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        smem2_set = smem2;
      }
      const int gx = std::max(1, std::min(16, (2 * n->sm_count + Z - 1) / Z));      // blocks per sample
      refine2_kernel<<<dim3(gx, Z), 256, smem2, st>>>(w.wl2, w.wl2_count, w.wl2_cap, x, B, D, H, n->bank, P, n->L.w1,
                                                       n->L.b1, n->L.w2, n->L.b2, z0, w.h2);
      n->launches++;
      RBNN_CUDA(cudaGetLastError());
    }
    *top = w.h2;
  }
  return 0;
}

// fused forward + head of one chunk (arch fc): dH (head >= 0) or logits (head < 0) straight from TMEM
static int fused_chunk(rbnn_net* n, const FcWs& w, int head, const float* x, const int32_t* labels, const float* pbar,
                       int B, int z0, int Z, float* logits, cudaStream_t st, uint32_t* maskbuf = nullptr) {
  const int H = n->H, D = n->D;
  const bool bf = n->prec == RBNN_PREC_BF16, f16 = n->prec == RBNN_PREC_F16X3;
  const TcMat& m1 = n->tc.mat[0];
  tc::FusedDesc f;
  f.mode = bf ? tc::MODE_BF16 : (f16 ? tc::MODE_F16X3 : tc::MODE_TF32X3);
  f.B = B; f.D = D; f.H = H; f.C = n->C; f.Z = Z;
  f.X.rows = B; f.X.ld = m1.ld;
  if (bf) {
    f.X.hi = w.x_bf;
    f.W1.hi = reinterpret_cast<const __nv_bfloat16*>(m1.bf) + (int64_t)z0 * H * m1.ld;
  } else if (f16) {
    f.X.hi = w.x_h16; f.X.lo = w.x_l16;
    f.W1.hi = reinterpret_cast<const __half*>(m1.h_hi) + (int64_t)z0 * H * m1.ld;
    f.W1.lo = reinterpret_cast<const __half*>(m1.h_lo) + (int64_t)z0 * H * m1.ld;
    f.unscale = w.call_sc + 1; f.dh_scale = w.call_sc + 2;
    f.xlo_zero = w.max_bits + 3;
  } else {
    f.X.hi = w.x_hi; f.X.lo = w.x_lo;
    f.W1.hi = m1.hi + (int64_t)z0 * H * m1.ld; f.W1.lo = m1.lo + (int64_t)z0 * H * m1.ld;
  }
  f.W1.rows = H; f.W1.ld = m1.ld; f.W1.zstride = (int64_t)H * m1.ld;
  f.head = head;
  f.bank = n->bank; f.P = n->L.P; f.b1_off = n->L.b1; f.wo_off = n->L.wo; f.bo_off = n->L.bo; f.z_row0 = z0;
  f.labels = labels; f.pbar = pbar;
  f.x = x; f.xnorm = w.xnorm; f.wnorm = n->tc.wnorm; f.eps = f16 ? kGuardEpsFusedF16 : kGuardEpsFused;
  f.dh_hi = f16 ? (void*)w.dtop_h16 : (void*)w.dtop_hi; f.dh_lo = f16 ? (void*)w.dtop_l16 : (void*)w.dtop_lo;
  f.dh_bf = w.dtop_bf; f.logits = logits;
  f.worklist = w.worklist;
  f.maskbuf = maskbuf;
  f.sm_count = n->sm_count;
  std::string err;
  if (head >= 0) RBNN_TRY(timing_begin(n, 1, st));
  if (tc::fused_forward_head(f, st, &err)) { set_error("%s", err.c_str()); return 1; }
  n->launches++;
  if (head >= 0) RBNN_TRY(timing_end(n, 1, st));
  if (head >= 0 && !bf) {
    if (tc::fused_fixup(f, st, &err)) { set_error("%s", err.c_str()); return 1; }
    n->launches++;
  }
  if (head == -2 && !bf && w.worklist) {
    if (tc::fused_keep_fixup(f, maskbuf, st, &err)) { set_error("%s", err.c_str()); return 1; }
    n->launches++;
  }
  return 0;
}

static int launch_head(rbnn_net* n, bool grad, int head, const float* top, int z0, int Z, const int32_t* labels,
                       const float* pbar, int B, float* logits, float* dh_hi, float* dh_lo, __nv_bfloat16* dh_bf,
                       cudaStream_t st) {
  const int H = n->H, C = n->C;
  const size_t smem = (size_t)(C * H + C) * sizeof(float);
  dim3 grid((B + kHeadWarps * kHeadRowsPerWarp - 1) / (kHeadWarps * kHeadRowsPerWarp), Z);
#define RBNN_HEAD_LAUNCH(CM, G)                                                                                   \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      RBNN_CUDA(cudaFuncSetAttribute(fc_head_kernel<CM, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    fc_head_kernel<CM, G><<<grid, kHeadWarps * 32, smem, st>>>(head, top, n->bank, n->L.P, n->L.wo, n->L.bo, z0,   \
                                                                labels, pbar, B, H, C, logits, dh_hi, dh_lo, dh_bf); \
  } while (0)
  RBNN_CHECK(smem <= 200 * 1024, "tcgen05 engine: n_classes * hidden too large for the head kernel");
  if (C <= 10) {          // MNIST / Fashion-MNIST / half-moons: no predicated padding classes
    if (grad) RBNN_HEAD_LAUNCH(10, true); else RBNN_HEAD_LAUNCH(10, false);
  } else if (C <= 16) {
    if (grad) RBNN_HEAD_LAUNCH(16, true); else RBNN_HEAD_LAUNCH(16, false);
  } else {
    if (grad) RBNN_HEAD_LAUNCH(32, true); else RBNN_HEAD_LAUNCH(32, false);
  }
#undef RBNN_HEAD_LAUNCH
  n->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// pbar_for_scale: the UPSTREAM head's d_pbar [B*C] (its magnitude bounds dH), else nullptr
static int split_x(rbnn_net* n, const float* x, int64_t count, FcWs& w, const float* pbar_for_scale, int64_t pbar_count,
                   float dh_factor, cudaStream_t st) {
  const int64_t n4 = count / 4;
  const int d4 = n->D / 4, ld4 = n->tc.mat[0].ld / 4;      // the copies take the row pitch of the W1 copies
  const unsigned blocks = (unsigned)std::min<int64_t>((n4 + 255) / 256, 148 * 16);
  if (n->prec == RBNN_PREC_F16X3) {
    // RBNN_XGRID=0: never take the two-pass route for inputs on the pixel grid (A/B measurements)
    static const bool xgrid_on = !(getenv("RBNN_XGRID") && atoi(getenv("RBNN_XGRID")) == 0);
    RBNN_CUDA(cudaMemsetAsync(w.max_bits, 0, 4 * sizeof(unsigned), st));
    if (xgrid_on && use_fused(n)) maxabs_grid_kernel<<<n->sm_count * 4, 256, 0, st>>>(x, count, w.max_bits);
    else maxabs_flat_kernel<<<n->sm_count * 4, 256, 0, st>>>(x, count, w.max_bits);
    if (pbar_for_scale) {
      maxabs_kernel<<<n->sm_count, 256, 0, st>>>(pbar_for_scale, 0, pbar_count, 1, w.max_bits + 1);
      n->launches++;
    }
    call_scales_kernel<<<1, 1, 0, st>>>(n->tc.scales, w.max_bits, pbar_for_scale ? w.max_bits + 1 : nullptr, w.call_sc,
                                        dh_factor, (xgrid_on && use_fused(n)) ? w.max_bits : nullptr);
    split_f16_kernel<<<blocks, 256, 0, st>>>(x, w.call_sc, w.x_h16, w.x_l16, n4, d4, ld4, w.max_bits);
    if (n->tc.xgrid_last)      // for rbnn_net_input_grid()
      RBNN_CUDA(cudaMemcpyAsync(n->tc.xgrid_last, w.max_bits + 3, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    n->launches += 3;
    RBNN_CUDA(cudaGetLastError());
    return 0;
  }
  split_kernel<<<blocks, 256, 0, st>>>(x, w.x_hi, w.x_lo, w.x_bf, n4, d4, ld4);
  n->launches++;
  RBNN_CUDA(cudaGetLastError());
  return 0;
}

// dX of a net with D < 8 inputs (half moons): part[z][b][d] = sum_j dH1[z][b][j] W1_z[j][d], one warp per (z, b) row, then
// a fixed-order sum over the samples.  (The CUDA-core GEMM pads N = D to a 64-column tile: 0.46 ms of a 1.1 ms evaluation
// at H = 512; this pass reads dH1 once.)
__global__ void __launch_bounds__(256)
dx_small_rows_kernel(const float* __restrict__ dh, const float* __restrict__ bank, int64_t P, int64_t w1_off, int z_row0,
                     int Z, int B, int H, int D, float* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)Z * B, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += nwarps) {
    const int z = (int)(row / B);
    const float* __restrict__ d = dh + row * H;
    const float* __restrict__ w = bank + (int64_t)(z_row0 + z) * P + w1_off;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int j = lane; j < H; j += 32) {
      const float v = __ldg(d + j);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < D) acc[k] = fmaf(v, __ldg(w + (int64_t)j * D + k), acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k >= D) break;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      if (lane == 0) part[row * D + k] = acc[k];
    }
  }
}
// out[b][d] (+)= sum_z part[z][b][d], fixed order
__global__ void dx_small_reduce_kernel(const float* __restrict__ part, int Z, int n, float* __restrict__ out, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = accumulate ? out[i] : 0.f;
  for (int z = 0; z < Z; ++z) t += __ldg(part + (int64_t)z * n + i);
  out[i] = t;
}

// rows of the batch per pass such that at least one sample fits the workspace budget
static int tc_batch_rows(const rbnn_net* n, int B, bool grad) {
  const size_t per_row = fc_per_z_bytes(n, 1024, grad) / 1024 + (size_t)n->D * 12 + 64;
  const size_t rows = std::max<size_t>(128, (n->ws_budget / 2) / per_row);
  return (int)std::min<size_t>(rows, (size_t)B);
}

static int tc_grad_pass(rbnn_net* n, int head, const float* x, const int32_t* labels, int B, int s0, int s1,
                        const float* pbar, float* out_sum, cudaStream_t st) {
  const int H = n->H, D = n->D;
  const bool two = n->arch == RBNN_ARCH_FC2, bf = n->prec == RBNN_PREC_BF16, f16 = n->prec == RBNN_PREC_F16X3;
  const int S = s1 - s0;
  const size_t per = fc_per_z_bytes(n, B, true);
  const bool kept = x == nullptr;       // gradient pass of a kept forward (n->keep): no forward GEMM, dH from logits + masks
  const size_t ldx = (size_t)k_pitch(n, D);
  const size_t x_bytes = kept ? 0 : (bf ? pad256((size_t)B * ldx * 2) : (f16 ? 2 * pad256((size_t)B * ldx * 2) + 512 : 2 * pad256((size_t)B * ldx * 4)));
  const size_t out_bytes = pad256((size_t)B * D * 4);
  // samples per chunk: as many as the budget allows; the backward reduce cuts a chunk into `slots`
  // K-concatenated ranges of ~4 samples (bounds the tensor-core accumulation chain, fills the SMs)
  size_t avail = n->ws_budget > x_bytes ? n->ws_budget - x_bytes : 0;
  int zc = (int)std::max<size_t>(1, std::min<size_t>(avail / (per + out_bytes / 4 + 1), (size_t)S));
  zc = (S + (S + zc - 1) / zc - 1) / ((S + zc - 1) / zc);      // equal chunks (1000 -> 3 x 334, not 419 + 419 + 162)
  // experiment knobs for the input-gradient GEMM: tile width and CTA pairs (cta_group::2)
  static const int env_bn = getenv("RBNN_BWD_BN") ? atoi(getenv("RBNN_BWD_BN")) : 0;
  static const int env_pair = getenv("RBNN_BWD_PAIR") ? atoi(getenv("RBNN_BWD_PAIR")) : 0;
  const int bn_b = env_bn ? env_bn : pick_bn(D);
  const int tiles_mn = ((B + tc::kBM - 1) / tc::kBM) * ((D + bn_b - 1) / bn_b);
  auto slots_for = [&](int Z) {
    int best = 1;
    double best_score = -1.0;
    for (int s = 1; s <= std::min(Z, 64); ++s) {
      const int tiles = tiles_mn * s;
      const int waves = (tiles + n->sm_count - 1) / n->sm_count;
      double score = (double)tiles / ((double)waves * n->sm_count);          // SM fill
      if ((Z + s - 1) / s > 4) score -= 0.02 * ((Z + s - 1) / s - 4);        // prefer chains of <= 4 samples
      if (score > best_score + 1e-9) { best_score = score; best = s; }
    }
    return best;
  };
  int slots = slots_for(zc);
  RBNN_TRY(ws_reserve(n, x_bytes + per * zc + out_bytes * slots + pad256((size_t)B * 4) + 4096 + 65536 +
                         (small_d(n) ? pad256((size_t)zc * B * D * 4) : 0)));
  Arena ar(n);
  FcWs w;
  if (kept) { w.call_sc = n->keep.call_sc; w.max_bits = n->keep.max_bits; }
  else if (bf) w.x_bf = ar.take<__nv_bfloat16>((size_t)B * ldx);
  else if (f16) {
    w.x_h16 = ar.take<__half>((size_t)B * ldx); w.x_l16 = ar.take<__half>((size_t)B * ldx);
    w.call_sc = ar.take<float>(4); w.max_bits = ar.take<unsigned>(4);
  } else { w.x_hi = ar.take<float>((size_t)B * ldx); w.x_lo = ar.take<float>((size_t)B * ldx); }
  const size_t zbh = (size_t)zc * B * H;
  const bool fused = use_fused(n);
  RBNN_CHECK(!f16 || fused, "F16X3 covers arch fc with a hidden layer the fused kernel supports");
  RBNN_CHECK(!kept || fused || n->keep.fc_h, "no kept hidden activations for the unfused route");
  const bool kept_unfused = kept && !fused;       // H1 [, H2] of every kept unit live in n->keep.fc_h
  const size_t keep_units = (size_t)(s1 - s0) * B * H;
  if (!fused && !kept) w.h1 = ar.take<float>(zbh);
  if (two && !kept) {
    if (bf) w.h1_bf = ar.take<__nv_bfloat16>(zbh); else w.h1_lo = ar.take<float>(zbh);
    w.h2 = ar.take<float>(zbh);
  }
  if (bf) w.dtop_bf = ar.take<__nv_bfloat16>(zbh);
  else if (f16) { w.dtop_h16 = ar.take<__half>(zbh); w.dtop_l16 = ar.take<__half>(zbh); }
  else { w.dtop_hi = ar.take<float>(zbh); w.dtop_lo = ar.take<float>(zbh); }
  if (two) {
    if (bf) w.d1_bf = ar.take<__nv_bfloat16>(zbh);
    else { w.d1_hi = ar.take<float>(zbh); w.d1_lo = ar.take<float>(zbh); }
  }
  if (fused && !bf && !kept) w.worklist = ar.take<unsigned long long>(tc::fused_worklist_slots(B, zc));
  if (two && !bf && !kept) {
    w.wl2_cap = (unsigned)std::max<size_t>(256, (size_t)B * H / 64);      // per sample of the chunk
    w.wl2 = ar.take<unsigned long long>((size_t)w.wl2_cap * zc);
    w.wl2_count = ar.take<unsigned>((size_t)zc);
  }
  w.xnorm = ar.take<float>((size_t)B);
  w.partial = ar.take<float>((size_t)slots * B * D);
  if (small_d(n)) w.dx_part = ar.take<float>((size_t)zc * B * D);
  if (kept) {
    if (f16) {       // the dH range of THIS head (max|x| is still in keep.max_bits[0] from the forward phase)
      const bool up = head == RBNN_HEAD_UPSTREAM || head == RBNN_HEAD_LOGITS_UPSTREAM;
      if (up) {
        RBNN_CUDA(cudaMemsetAsync(w.max_bits + 1, 0, sizeof(unsigned), st));
        maxabs_kernel<<<n->sm_count, 256, 0, st>>>(pbar, 0, (int64_t)B * n->C, 1, w.max_bits + 1);
        n->launches++;
      }
      call_scales_kernel<<<1, 1, 0, st>>>(n->tc.scales, w.max_bits, up ? w.max_bits + 1 : nullptr, w.call_sc,
                                          head == RBNN_HEAD_LOGITS_UPSTREAM ? (float)n->C : 2.f, w.max_bits);
      n->launches++;
      RBNN_CUDA(cudaGetLastError());
    }
  } else {
    // |dH| <= sum_c |dlogits_c| max|Wo|: <= 2 max|g| for the softmax heads, <= C max|g| when g goes to the logits as is
    if (!small_d(n))
      RBNN_TRY(split_x(n, x, (int64_t)B * D, w,
                       (head == RBNN_HEAD_UPSTREAM || head == RBNN_HEAD_LOGITS_UPSTREAM) ? pbar : nullptr, (int64_t)B * n->C,
                       head == RBNN_HEAD_LOGITS_UPSTREAM ? (float)n->C : 2.f, st));
  }
  if (fused && !kept) {       // row norms: guard band (parity modes) and the activation range of the fused head
    xnorm_kernel<<<(B + 7) / 8, 256, 0, st>>>(x, B, D, w.xnorm);
    n->launches++;
    RBNN_CUDA(cudaGetLastError());
  }

  bool first = true;
  for (int z0 = s0; z0 < s1; z0 += zc) {
    const int Z = std::min(zc, s1 - z0);
    const int sl = Z == zc ? slots : std::min(slots, slots_for(Z));
    if (kept_unfused) {
      w.h1 = n->keep.fc_h + (size_t)(z0 - s0) * B * H;
      if (two) w.h2 = n->keep.fc_h + keep_units + (size_t)(z0 - s0) * B * H;
      RBNN_TRY(launch_head(n, true, head, two ? w.h2 : w.h1, z0, Z, labels, pbar, B, nullptr, w.dtop_hi, w.dtop_lo,
                           w.dtop_bf, st));
    } else if (kept) {
      tc::KeptDesc k;
      k.mode = bf ? tc::MODE_BF16 : (f16 ? tc::MODE_F16X3 : tc::MODE_TF32X3);
      k.B = B; k.H = H; k.C = n->C; k.Z = Z;
      k.bank = n->bank; k.P = n->L.P; k.wo_off = n->L.wo; k.z_row0 = z0;
      k.head = head; k.labels = labels; k.pbar = pbar;
      k.logits = n->keep.logits + (size_t)(z0 - s0) * B * n->C;
      k.masks = n->keep.masks + tc::keep_mask_words(B, z0 - s0);
      k.dh_hi = f16 ? (void*)w.dtop_h16 : (void*)w.dtop_hi; k.dh_lo = f16 ? (void*)w.dtop_l16 : (void*)w.dtop_lo;
      k.dh_bf = w.dtop_bf;
      k.dh_scale = f16 ? w.call_sc + 2 : nullptr;
      k.sm_count = n->sm_count;
      std::string err;
      RBNN_TRY(timing_begin(n, 1, st));
      if (tc::dh_from_kept(k, st, &err)) { set_error("%s", err.c_str()); return 1; }
      n->launches++;
      RBNN_TRY(timing_end(n, 1, st));
    } else if (fused) {
      RBNN_TRY(fused_chunk(n, w, head, x, labels, pbar, B, z0, Z, nullptr, st));
    } else {
      const float* top = nullptr;
      RBNN_TRY(fc_forward_chunk_tc(n, w, x, B, z0, Z, &top, st));
      RBNN_TRY(launch_head(n, true, head, top, z0, Z, labels, pbar, B, nullptr, w.dtop_hi, w.dtop_lo, w.dtop_bf, st));
    }
    const void* dfirst_hi = bf ? (const void*)w.dtop_bf : (f16 ? (const void*)w.dtop_h16 : (const void*)w.dtop_hi);
    const void* dfirst_lo = f16 ? (const void*)w.dtop_l16 : (const void*)w.dtop_lo;
    if (two) {
      // dH1 = (dH2 . W2) (.) leaky'(H1): per-z GEMM with K = H over the transposed copy of W2
      const TcMat& m2 = n->tc.mat[1];
      tc::GemmDesc q;
      q.M = B; q.N = H; q.K = H; q.Z = Z; q.BN = pick_bn(H);
      q.A.hi = dfirst_hi; q.A.lo = dfirst_lo; q.A.rows = B; q.A.ld = H; q.A.zstride = (int64_t)B * H;
      if (bf) q.B.hi = reinterpret_cast<const __nv_bfloat16*>(m2.tbf) + (int64_t)z0 * H * H;
      else { q.B.hi = m2.thi + (int64_t)z0 * H * H; q.B.lo = m2.tlo + (int64_t)z0 * H * H; }
      q.B.rows = H; q.B.ld = H; q.B.zstride = (int64_t)H * H;
      q.epi = tc::EPI_MASK;
      q.act = w.h1; q.act_zstride = (int64_t)B * H; q.act_ld = H;
      if (bf) q.out_bf = w.d1_bf; else { q.out = w.d1_hi; q.out_lo = small_d(n) ? nullptr : w.d1_lo; }   // small D: plain fp32 for the CUDA-core GEMM below
      q.out_ld = H; q.out_zstride = (int64_t)B * H;
      RBNN_TRY(run_gemm(n, q, 0, st));
      dfirst_hi = bf ? (const void*)w.d1_bf : (const void*)w.d1_hi;
      dfirst_lo = w.d1_lo;
    }
    if (small_d(n)) {
      // dX (+)= sum_z dH1_z . W1_z with N = D < 8: the FP32 engine's CUDA-core GEMM, summed over the chunk's samples
      // dX (+)= sum_z dH1_z . W1_z with N = D < 8: per-sample partials (w.dx_part), then a fixed-order sum over the samples
      RBNN_TRY(timing_begin(n, 2, st));
      const int64_t rows = (int64_t)Z * B;
      dx_small_rows_kernel<<<(unsigned)std::min<int64_t>((rows + 7) / 8, (int64_t)n->sm_count * 8), 256, 0, st>>>(
          w.d1_hi, n->bank, n->L.P, n->L.w1, z0, Z, B, H, D, w.dx_part);
      dx_small_reduce_kernel<<<(B * D + 255) / 256, 256, 0, st>>>(w.dx_part, Z, B * D, out_sum, first ? 0 : 1);
      n->launches += 2;
      RBNN_CUDA(cudaGetLastError());
      RBNN_TRY(timing_end(n, 2, st));
      first = false;
      continue;
    }
    // dX partial sums: K-concatenated over the samples of each slot
    const TcMat& m1 = n->tc.mat[0];
    tc::GemmDesc r;
    r.M = B; r.N = D; r.K = H; r.Z = Z; r.BN = bn_b;
    r.A.hi = dfirst_hi; r.A.lo = dfirst_lo; r.A.rows = B; r.A.ld = H; r.A.zstride = (int64_t)B * H;
    if (bf) r.B.hi = reinterpret_cast<const __nv_bfloat16*>(m1.tbf) + (int64_t)z0 * D * H;
    else if (f16) {
      r.B.hi = reinterpret_cast<const __half*>(m1.th_hi) + (int64_t)z0 * D * H;
      r.B.lo = reinterpret_cast<const __half*>(m1.th_lo) + (int64_t)z0 * D * H;
      r.unscale = w.call_sc + 3;
    } else { r.B.hi = m1.thi + (int64_t)z0 * D * H; r.B.lo = m1.tlo + (int64_t)z0 * D * H; }
    r.B.rows = D; r.B.ld = H; r.B.zstride = (int64_t)D * H;
    r.reduce_z = 1; r.slots = sl;
    r.pair = env_pair ? 1 : 0; r.pair_relay = env_pair == 2 ? 0 : 1;
    r.out = w.partial; r.out_ld = D; r.out_zstride = (int64_t)B * D;
    RBNN_TRY(run_gemm(n, r, 2, st));
    const int64_t n4 = (int64_t)B * D / 4;
    reduce_slots_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(w.partial, sl, n4, out_sum, first ? 0 : 1);
    n->launches++;
    RBNN_CUDA(cudaGetLastError());
    first = false;
  }
  return 0;
}

int tc_fc_input_grad_sum(rbnn_net* n, int head, const float* x, const int32_t* labels, int B, int s0, int s1,
                         const float* pbar, float* out_sum, cudaStream_t st) {
  RBNN_CHECK(tc_supported(n), "tcgen05 engine does not cover this network");
  RBNN_TRY(tc_bank_refresh(n, s0, s1, st));
  const int bc = tc_batch_rows(n, B, true);
  for (int b0 = 0; b0 < B; b0 += bc) {
    const int nb = std::min(bc, B - b0);
    RBNN_TRY(tc_grad_pass(n, head, x + (int64_t)b0 * n->D, labels + b0, nb, s0, s1,
                          pbar ? pbar + (int64_t)b0 * n->C : nullptr, out_sum + (int64_t)b0 * n->D, st));
  }
  return 0;
}

static int tc_forward_pass(rbnn_net* n, const float* x, int B, int s0, int s1, float* out_sum, float* out_logits,
                           cudaStream_t st, bool keep = false) {
  const int H = n->H, D = n->D, C = n->C;
  const bool two = n->arch == RBNN_ARCH_FC2, bf = n->prec == RBNN_PREC_BF16, f16 = n->prec == RBNN_PREC_F16X3;
  const int S = s1 - s0;
  // keep mode: logits go to n->keep (all samples), the arena holds the guard-band worklist of a chunk
  const size_t per = (keep && use_fused(n)) ? pad256(tc::fused_worklist_slots(B, 1) * 8) + 256 : fc_per_z_bytes(n, B, false);
  const size_t ldx = (size_t)k_pitch(n, D);
  const size_t x_bytes = bf ? pad256((size_t)B * ldx * 2) : (f16 ? 2 * pad256((size_t)B * ldx * 2) + 512 : 2 * pad256((size_t)B * ldx * 4));
  size_t avail = n->ws_budget > x_bytes ? n->ws_budget - x_bytes : 0;
  int zc = (int)std::max<size_t>(1, std::min<size_t>(avail / per, (size_t)S));
  zc = (S + (S + zc - 1) / zc - 1) / ((S + zc - 1) / zc);      // equal chunks
  RBNN_TRY(ws_reserve(n, x_bytes + per * zc + pad256((size_t)B * 4) + 65536));
  Arena ar(n);
  FcWs w;
  if (bf) w.x_bf = ar.take<__nv_bfloat16>((size_t)B * ldx);
  else if (f16) {
    w.x_h16 = ar.take<__half>((size_t)B * ldx); w.x_l16 = ar.take<__half>((size_t)B * ldx);
    if (keep) { w.call_sc = n->keep.call_sc; w.max_bits = n->keep.max_bits; }
    else { w.call_sc = ar.take<float>(4); w.max_bits = ar.take<unsigned>(4); }
  } else { w.x_hi = ar.take<float>((size_t)B * ldx); w.x_lo = ar.take<float>((size_t)B * ldx); }
  const size_t zbh = (size_t)zc * B * H;
  const bool fused = use_fused(n);
  RBNN_CHECK(!f16 || fused, "F16X3 covers arch fc with a hidden layer the fused kernel supports");
  const bool keep_unfused = keep && !fused;       // H1 [, H2] go to n->keep.fc_h (all samples) instead of the arena
  const size_t keep_units = (size_t)S * B * H;
  if (!fused && !keep) w.h1 = ar.take<float>(zbh);
  if (two) {
    if (bf) w.h1_bf = ar.take<__nv_bfloat16>(zbh); else w.h1_lo = ar.take<float>(zbh);
    if (!keep) w.h2 = ar.take<float>(zbh);
  }
  if (keep && fused) { if (!bf) w.worklist = ar.take<unsigned long long>(tc::fused_worklist_slots(B, zc)); }
  else w.logits = ar.take<float>((size_t)zc * B * C);
  if (two && !bf) {
    w.wl2_cap = (unsigned)std::max<size_t>(256, (size_t)B * H / 64);      // per sample of the chunk
    w.wl2 = ar.take<unsigned long long>((size_t)w.wl2_cap * zc);
    w.wl2_count = ar.take<unsigned>((size_t)zc);
  }
  w.xnorm = ar.take<float>((size_t)B);
  if (!small_d(n)) RBNN_TRY(split_x(n, x, (int64_t)B * D, w, nullptr, 0, 2.f, st));
  if (fused) {
    xnorm_kernel<<<(B + 7) / 8, 256, 0, st>>>(x, B, D, w.xnorm);
    n->launches++;
    RBNN_CUDA(cudaGetLastError());
  }
  for (int z0 = s0; z0 < s1; z0 += zc) {
    const int Z = std::min(zc, s1 - z0);
    float* lg = (keep && fused) ? n->keep.logits + (size_t)(z0 - s0) * B * C : (out_logits ? out_logits : w.logits);
    if (keep_unfused) {
      w.h1 = n->keep.fc_h + (size_t)(z0 - s0) * B * H;
      if (two) w.h2 = n->keep.fc_h + keep_units + (size_t)(z0 - s0) * B * H;
    }
    if (keep && fused) {
      RBNN_TRY(fused_chunk(n, w, -2, x, nullptr, nullptr, B, z0, Z, lg, st,
                           n->keep.masks + tc::keep_mask_words(B, z0 - s0)));
    } else if (fused) {
      RBNN_TRY(fused_chunk(n, w, -1, x, nullptr, nullptr, B, z0, Z, lg, st));
    } else {
      const float* top = nullptr;
      RBNN_TRY(fc_forward_chunk_tc(n, w, x, B, z0, Z, &top, st));
      RBNN_TRY(launch_head(n, false, 0, top, z0, Z, nullptr, nullptr, B, lg, nullptr, nullptr, nullptr, st));
    }
    if (out_sum) RBNN_TRY(head_probs_accumulate(n, lg, Z, B, C, out_sum, st));
  }
  return 0;
}

int tc_fc_forward(rbnn_net* n, const float* x, int B, int s0, int s1, float* out_sum, float* out_logits,
                  cudaStream_t st) {
  RBNN_CHECK(tc_supported(n), "tcgen05 engine does not cover this network");
  RBNN_TRY(tc_bank_refresh(n, s0, s1, st));
  const int bc = tc_batch_rows(n, B, false);
  for (int b0 = 0; b0 < B; b0 += bc) {
    const int nb = std::min(bc, B - b0);
    RBNN_TRY(tc_forward_pass(n, x + (int64_t)b0 * n->D, nb, s0, s1, out_sum ? out_sum + (int64_t)b0 * n->C : nullptr,
                             out_logits ? out_logits + (int64_t)b0 * n->C : nullptr, st));
  }
  return 0;
}

void tc_keep_free(rbnn_net* n) {
  cudaFree(n->keep.logits); cudaFree(n->keep.masks); cudaFree(n->keep.call_sc); cudaFree(n->keep.max_bits);
  cudaFree(n->keep.conv_buf); cudaFree(n->keep.fc_h);
  n->keep = KeepCache();
}

// Phase 1 of the two-phase attack gradient: out_sum[B, C] += sum_s softmax(logits_s) like tc_fc_forward, and the
// per-sample logits + LeakyReLU masks stay in n->keep.  Falls back to the plain forward (keep.valid = 0) when the
// engine / shape has no fused route or the batch does not fit one pass.
int tc_fc_forward_keep(rbnn_net* n, const float* x, int B, int s0, int s1, float* out_sum, cudaStream_t st) {
  RBNN_CHECK(tc_supported(n), "tcgen05 engine does not cover this network");
  n->keep.valid = 0;
  if (tc_batch_rows(n, B, false) < B || tc_batch_rows(n, B, true) < B)
    return tc_fc_forward(n, x, B, s0, s1, out_sum, nullptr, st);
  KeepCache& k = n->keep;
  if (!use_fused(n)) {
    // unfused route (fc2, RBNN_TC_UNFUSED): keep the refined hidden activations themselves (fp32, 4 B per hidden unit)
    const size_t need = (size_t)(s1 - s0) * B * n->H * (n->arch == RBNN_ARCH_FC2 ? 2 : 1);
    if (need * sizeof(float) > ((size_t)16 << 30)) return tc_fc_forward(n, x, B, s0, s1, out_sum, nullptr, st);
    RBNN_TRY(tc_bank_refresh(n, s0, s1, st));
    if (k.fc_cap < need) {
      RBNN_CUDA(cudaDeviceSynchronize());
      cudaFree(k.fc_h);
      k.fc_h = nullptr; k.fc_cap = 0;
      RBNN_CUDA(cudaMalloc(&k.fc_h, need * sizeof(float)));
      k.fc_cap = need;
      n->alloc_epoch++;
    }
    RBNN_TRY(tc_forward_pass(n, x, B, s0, s1, out_sum, nullptr, st, true));
    k.valid = 1; k.B = B; k.s0 = s0; k.s1 = s1;
    return 0;
  }
  RBNN_TRY(tc_bank_refresh(n, s0, s1, st));
  const size_t need_l = (size_t)(s1 - s0) * B * n->C, need_m = tc::keep_mask_words(B, s1 - s0);
  if (k.logits_cap < need_l || k.masks_cap < need_m) {
    RBNN_CUDA(cudaDeviceSynchronize());
    cudaFree(k.logits); cudaFree(k.masks);
    k.logits = nullptr; k.masks = nullptr; k.logits_cap = k.masks_cap = 0;
    RBNN_CUDA(cudaMalloc(&k.logits, need_l * sizeof(float)));
    RBNN_CUDA(cudaMalloc(&k.masks, need_m * sizeof(uint32_t)));
    k.logits_cap = need_l; k.masks_cap = need_m;
    n->alloc_epoch++;
  }
  if (!k.call_sc) {
    RBNN_CUDA(cudaMalloc(&k.call_sc, 4 * sizeof(float)));
    RBNN_CUDA(cudaMalloc(&k.max_bits, 4 * sizeof(unsigned)));
    n->alloc_epoch++;
  }
  RBNN_TRY(tc_forward_pass(n, x, B, s0, s1, out_sum, nullptr, st, true));
  k.valid = 1; k.B = B; k.s0 = s0; k.s1 = s1;
  return 0;
}

// Phase 2: out_sum[B, D] = sum_s dL/dx from the kept forward (head: GRAD_OF_MEAN / UPSTREAM with d_pbar, or MEAN_OF_GRADS)
int tc_fc_grad_kept(rbnn_net* n, int head, const int32_t* labels, const float* pbar, float* out_sum, cudaStream_t st) {
  RBNN_CHECK(n->keep.valid, "no kept forward: call rbnn_forward_probs_sum_keep first (and check rbnn_keep_valid)");
  RBNN_CHECK(head != RBNN_HEAD_LOGITS_CE, "LOGITS_CE has no kept route");
  return tc_grad_pass(n, head, nullptr, labels, n->keep.B, n->keep.s0, n->keep.s1, pbar, out_sum, st);
}

}  // namespace rbnn
