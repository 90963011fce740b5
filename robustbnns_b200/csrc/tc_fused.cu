// Fused forward + loss head for arch fc on tcgen05 (see FusedDesc in tc_gemm.cuh).
//
// Replaces, for every posterior sample z and every input b at once, the reference's
//   output = net.forward(x, n_samples=1, seeds=[z]); loss = CE(output, y); loss.backward()
// up to the gradient w.r.t. the hidden layer (lossGradients.py:33-36 / adversarialAttacks.py:74-78;
// model_nn.py:77-82 for the network).  The hidden activations never leave the SM:
//
//   work item = (sample z, tile of 128 inputs); per item the hidden dimension is cut into n-tiles of
//   BN <= 256 columns that alternate between the two TMEM accumulator stages.
//   warp 0      TMA producer (X tile + W1_z tile K-blocks, SWIZZLE_128B ring)
//   warp 1      tcgen05.mma issuer (kind::tf32 x3 passes, kind::f16 x3 passes on scaled fp16 hi/lo, or kind::f16 on bf16)
//   warps 2..9  epilogue: warp = (TMEM lane quadrant, column half).  Pass 1 (per n-tile, overlapping the next
//               n-tile's MMAs): tcgen05.ld -> +b1 -> LeakyReLU -> mask bit, partial logits += h * Wo (Wo_z in
//               shared memory), guard-band check.  After the last n-tile the two column halves exchange their
//               partial logits through shared memory, every thread evaluates the loss head of its input
//               row, and pass 2 rebuilds dH = (dlogits . Wo) * mask from the mask bits and stores it pre-split.
//
// Guard band: a pre-activation closer to zero than eps * ||x_b|| * max_j ||w_zj|| (a bound on what the tensor
// core rounding can move) is queued as (z, b, j, assumed sign); fused_fixup re-evaluates those dot products
// exactly (fp64 accumulation) and rescales dH[z, b, j] when the sign was wrong.  Each item owns kWorkPerItem
// slots of the worklist (no global atomics, deterministic); an item that overflows evaluates inline.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "../../include/rbnn.h"
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace rbnn {
namespace tc {

namespace {

constexpr float kSlopeF = 0.01f;
constexpr int kRingF = 196608;
constexpr int kEpiWarps = 16;                     // 4 TMEM lane quadrants x 2 lane halves x 2 column halves
constexpr int kThreadsF = 64 + kEpiWarps * 32;     // 576
constexpr int kCMax = 10;                          // classes the fused epilogue covers (two mma.sync class groups: 0-7, 8-15)
constexpr int kHMax = 512;                         // hidden units the staged parameters cover
constexpr int kMaskWords = 8;                      // 32x32 blocks per thread: n_tiles * (BN/2)/32 <= 8
constexpr int kWoRows = kCMax + 1;                 // + the zero row the padding classes read
constexpr int kWoHalfs = kWoRows * (kHMax + 8);    // one fp16 copy (hi or lo) of s_wo * Wo_z, rows padded by 8 halfs
constexpr int kWoPerThread = (kCMax * kHMax + kEpiWarps * 32 - 1) / (kEpiWarps * 32);
constexpr int kXchgFloats = kBM * 16;              // partial logits / dlogits exchange between column halves
// shared memory after the ring and the barriers: Wo16 hi | Wo16 lo | b1 [kHMax] | bo [16] | red [2 * kEpiWarps] | xchg
constexpr int kFusedSmem = kRingF + 1024 + 256 + 2 * kWoHalfs * 2 + (kHMax + 16 + 2 * kEpiWarps + kXchgFloats) * 4 + 16;
static_assert(kFusedSmem <= 232448, "fused kernel exceeds the 227 KB shared memory of an sm_100 CTA");
constexpr unsigned long long kSentinel = ~0ull;

struct FParams {
  int B, D, H, C, Z, BN, n_tiles, m_tiles, num_items, num_kb;
  int head;
  const float* bank; long long P, b1_off, wo_off, bo_off; int z_row0;
  const int32_t* labels; const float* pbar;
  const float* x; const float* xnorm; const float* wnorm; float eps;
  void* dh_hi; void* dh_lo; __nv_bfloat16* dh_bf; float* logits;
  const float* unscale; const float* dh_scale;      // F16X3 device scalars (see FusedDesc)
  uint32_t* maskbuf;                                 // head == -2 (keep mode): [item][kMaskWords][512] LeakyReLU mask words
  int debug;                                         // timing experiments (RBNN_FUSED_DEBUG): 1 no dH stores, 2 no pass 2, 4 no pass-1 math
  unsigned long long* worklist;
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

__device__ __forceinline__ unsigned long long pack_entry(int z, int b, int j, bool pos) {
  return ((unsigned long long)z << 44) | ((unsigned long long)b << 20) | ((unsigned long long)j << 4) |
         (pos ? 1ull : 0ull);
}

// exact sign of b1[j] + <x_b, w_j> (fp64 accumulation of the exact fp32 products), one thread
__device__ bool exact_positive_serial(const float* __restrict__ x, const float* __restrict__ w, float bias, int D) {
  double s = 0.0;
  for (int d = 0; d < D; ++d) s = fma((double)__ldg(x + d), (double)__ldg(w + d), s);
  return (float)(s + (double)bias) > 0.f;
}

// s = 2^k with s * maxabs in [2^8, 2^9) (1 for zero / denormal-range / non-finite maxima): exact scaling into the
// range where an fp16 hi/lo pair keeps 22 significant bits
__device__ __forceinline__ float pow2_scale(float maxabs) {
  const uint32_t eb = (__float_as_uint(maxabs) >> 23) & 0xFFu;
  return (eb < 8u || eb == 255u) ? 1.f : __uint_as_float((262u - eb) << 23);
}
// (x, y) -> packed fp16 pairs hi = rn(x, y), lo = rn((x, y) - hi); the lower half holds x
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - f.x, y - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// c[16x8] += a[16x16] . b[16x8], fp16 operands from registers, fp32 accumulation
__device__ __forceinline__ void hmma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 4 x 4 transpose of 32-bit words across the 4 lanes (q = lane % 4) that share a fragment row: on return a[k] of lane q is
// what lane k held in a[q].  Turns "2 columns of each of 4 column groups" into "8 consecutive columns of group q".
__device__ __forceinline__ void quad_transpose(uint32_t (&a)[4], int q) {
  const bool odd = q & 1, up = q & 2;
  uint32_t s0 = odd ? a[0] : a[1], s1 = odd ? a[2] : a[3];
  s0 = __shfl_xor_sync(0xffffffffu, s0, 1);
  s1 = __shfl_xor_sync(0xffffffffu, s1, 1);
  if (odd) { a[0] = s0; a[2] = s1; } else { a[1] = s0; a[3] = s1; }
  uint32_t t0 = up ? a[0] : a[2], t1 = up ? a[1] : a[3];
  t0 = __shfl_xor_sync(0xffffffffu, t0, 2);
  t1 = __shfl_xor_sync(0xffffffffu, t1, 2);
  if (up) { a[0] = t0; a[1] = t1; } else { a[2] = t0; a[3] = t1; }
}
__device__ __forceinline__ void st_cs_u4(void* ptr, const uint32_t (&a)[4]) {
  asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]) : "memory");
}

// softmax over the classes of one row spread over the 4 lanes that share it (4 class slots per lane)
__device__ __forceinline__ void softmax_quad(float (&v)[4], const bool (&valid)[4]) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (valid[i]) mx = fmaxf(mx, v[i]);
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[i] = valid[i] ? expf(v[i] - mx) : 0.f; sum += v[i]; }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] *= inv;
}

template <int C_MAX>
__device__ __forceinline__ void softmax_r(float (&v)[C_MAX], int C) {
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


// Loss head on one input row whose class values are spread over the 4 lanes of a fragment row (slot i of this lane
// holds class cls[i]): l = logits (bias included) -> l = dL/dlogits.  Same algebra as head.cu::dlogits_kernel.
__device__ __forceinline__ void head_quad(int head, float (&l)[4], const bool (&valid)[4], const int (&cls)[4], int y,
                                          const float* __restrict__ pbar_row) {
  float g[4];
  if (head == RBNN_HEAD_LOGITS_UPSTREAM) {     // loss of the mean LOGITS (ensembles, deterministic nets): dlogits = d_pbar
#pragma unroll
    for (int i = 0; i < 4; ++i) l[i] = (valid[i] && pbar_row) ? __ldg(pbar_row + cls[i]) : 0.f;
    return;
  }
  softmax_quad(l, valid);
  if (head == RBNN_HEAD_LOGITS_CE) {
#pragma unroll
    for (int i = 0; i < 4; ++i) l[i] = valid[i] ? l[i] - (cls[i] == y ? 1.f : 0.f) : 0.f;
    return;
  }
  if (head == RBNN_HEAD_MEAN_OF_GRADS) {
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = l[i];
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = (valid[i] && pbar_row) ? __ldg(pbar_row + cls[i]) : 0.f;
  }
  if (head != RBNN_HEAD_UPSTREAM) {
    softmax_quad(g, valid);
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] -= (cls[i] == y ? 1.f : 0.f);
  }
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (valid[i]) dot = fmaf(l[i], g[i], dot);
  dot += __shfl_xor_sync(0xffffffffu, dot, 1);
  dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
  for (int i = 0; i < 4; ++i) l[i] = valid[i] ? l[i] * (g[i] - dot) : 0.f;
}

// A fragments of pass 2: dlogits of rows (g, g+8) x classes, scaled per row into the fp16 range and split;
// mul[r] = dh_scale / (row scale * s_wo) turns the accumulator of row r into the stored dH
__device__ __forceinline__ void dlogits_frags(const float (&dl)[2][4], float s_wo, float dh_scale, uint32_t (&Dh)[4],
                                              uint32_t (&Dl)[4], float (&mul)[2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float mx = fmaxf(fmaxf(fabsf(dl[r][0]), fabsf(dl[r][1])), fmaxf(fabsf(dl[r][2]), fabsf(dl[r][3])));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float s_d = pow2_scale(mx);
    split_pair(dl[r][0] * s_d, dl[r][1] * s_d, Dh[r], Dl[r]);            // classes 2q, 2q+1
    split_pair(dl[r][2] * s_d, dl[r][3] * s_d, Dh[2 + r], Dl[2 + r]);    // classes 8+2q, 9+2q
    mul[r] = dh_scale / (s_d * s_wo);
  }
}

// Pass 2 on one 16 x 32 block (rows g, g+8 of this lane; hidden columns j0 .. j0+31): dH = (dlogits . Wo) * leaky'(H),
// written as fp16 hi/lo (F16X3), tf32 hi/lo (TF32X3) or bf16.  `bits`: the block's mask word, bit (r*8 + k*2 + e).
template <int MODE>
__device__ __forceinline__ void pass2_block(const uint32_t (&Dh)[4], const uint32_t (&Dl)[4], const float (&mul)[2],
                                            uint32_t bits, uint32_t wo_hi_a, uint32_t wo_lo_a, uint32_t off2, int j0, int q,
                                            const bool (&rok)[2], const long long (&orow)[2], void* dh_hi, void* dh_lo,
                                            __nv_bfloat16* dh_bf, bool no_store) {
  constexpr bool BF16 = MODE == MODE_BF16;
  uint32_t whi[2][4], wlo[2][4];                        // [row][column group]: packed 16-bit pairs (hi, lo / bf16)
#pragma unroll
  for (int gp = 0; gp < 2; ++gp) {                      // 16 columns: two 8-column groups
    uint32_t bh[4], bl[4];
    ldsm_x4_trans(wo_hi_a + off2 + (uint32_t)(j0 + 16 * gp) * 2u, bh);
    ldsm_x4_trans(wo_lo_a + off2 + (uint32_t)(j0 + 16 * gp) * 2u, bl);
#pragma unroll
    for (int gs = 0; gs < 2; ++gs) {
      const int k = 2 * gp + gs;
      float d[4] = {0.f, 0.f, 0.f, 0.f};                // row 0: (j, j+1), row 1: (j, j+1)
      hmma_16816(d, Dl, bh[2 * gs], bh[2 * gs + 1]);
      hmma_16816(d, Dh, bl[2 * gs], bl[2 * gs + 1]);
      hmma_16816(d, Dh, bh[2 * gs], bh[2 * gs + 1]);
      const int j = j0 + 8 * k + 2 * q;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float d0 = d[2 * r] * mul[r], d1 = d[2 * r + 1] * mul[r];
        if (!((bits >> (r * 8 + k * 2)) & 1u)) d0 *= kSlopeF;
        if (!((bits >> (r * 8 + k * 2 + 1)) & 1u)) d1 *= kSlopeF;
        if (BF16) {
          const __nv_bfloat162 a = __floats2bfloat162_rn(d0, d1);
          whi[r][k] = *reinterpret_cast<const uint32_t*>(&a);
        } else if (MODE == MODE_F16X3) {
          split_pair(d0, d1, whi[r][k], wlo[r][k]);
        } else if (rok[r] && !no_store) {
          float2 hi2, lo2;
          hi2.x = to_tf32_rn(d0); hi2.y = to_tf32_rn(d1);
          lo2.x = d0 - hi2.x; lo2.y = d1 - hi2.y;
          __stcs(reinterpret_cast<float2*>(reinterpret_cast<float*>(dh_hi) + orow[r] + j), hi2);   // streaming: do not displace X / W1 in L2
          __stcs(reinterpret_cast<float2*>(reinterpret_cast<float*>(dh_lo) + orow[r] + j), lo2);
        }
      }
    }
  }
  if (MODE != MODE_TF32X3) {
    // 16-bit outputs: gather 8 consecutive columns per lane (4 x 4 word transpose over the lanes of a row), then
    // one 16-byte streaming store per row and array -- whole 32-byte sectors instead of 4-byte pieces
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      quad_transpose(whi[r], q);
      if (!BF16) quad_transpose(wlo[r], q);
      if (!rok[r] || no_store) continue;
      const long long o = orow[r] + j0 + 8 * q;
      if (BF16) {
        st_cs_u4(dh_bf + o, whi[r]);
      } else {
        st_cs_u4(reinterpret_cast<__half*>(dh_hi) + o, whi[r]);
        st_cs_u4(reinterpret_cast<__half*>(dh_lo) + o, wlo[r]);
      }
    }
  }
}

template <int MODE, int KBB>
__global__ void __launch_bounds__(kThreadsF, 1)
fc_fused_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                const FParams p) {
  constexpr bool BF16 = MODE == MODE_BF16;
  constexpr bool F16K = MODE != MODE_TF32X3;            // kind::f16 MMAs on 2-byte operands
  constexpr int NARR = BF16 ? 1 : 2;
  constexpr int kATileF = kBM * KBB, kBTileF = kBNMax * KBB;
  constexpr int STAGE = NARR * (kATileF + kBTileF);
  constexpr int NSTAGE = kRingF / STAGE;
  constexpr int KBE = KBB / (F16K ? 2 : 4);
  constexpr int KSTEPS = KBB / 32;
  constexpr uint32_t FMT = MODE == MODE_BF16 ? 1u : (MODE == MODE_F16X3 ? 0u : 2u);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  const uint32_t bars = ring + kRingF;
  const uint32_t full0 = bars, empty0 = bars + 8 * NSTAGE;
  const uint32_t tfull0 = bars + 16 * NSTAGE, tempty0 = tfull0 + 16;
  uint8_t* gen = smem_raw + (bars - raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 16 * NSTAGE + 32);
  uint32_t* wl_count = reinterpret_cast<uint32_t*>(gen + 16 * NSTAGE + 40);
  __half* wo_hi = reinterpret_cast<__half*>(gen + 256);   // [C + 1][H + 8]: fp16 hi part of s_wo * Wo_z (row C = zeros)
  __half* wo_lo = wo_hi + kWoHalfs;                        // the residual part
  float* b1_s = reinterpret_cast<float*>(wo_lo + kWoHalfs);   // [H]
  float* bo_s = b1_s + kHMax;                              // [16]
  float* red = bo_s + 16;                                  // [2][kEpiWarps] block reductions
  float* xchg = red + 2 * kEpiWarps;                       // [128][16] partial logits / dlogits

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
    if (!BF16) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t stage_tx = (uint32_t)NARR * (uint32_t)(kATileF + p.BN * KBB);

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0;
      const long long t_begin = clock64();
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int z = item / p.m_tiles, m_idx = item % p.m_tiles;
        for (int n = 0; n < p.n_tiles; ++n) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            const long long t0 = clock64();
            mbar_wait(empty0 + 8 * stage, phase ^ 1u);
            w_empty += clock64() - t0;
            const uint32_t fb = full0 + 8 * stage;
            mbar_expect_tx(fb, stage_tx);
            const uint32_t sa = ring + stage * STAGE;
            const uint32_t sb = sa + NARR * kATileF;
            tma_load_3d(sa, &tmAh, fb, kb * KBE, m_idx * kBM, 0);
            tma_load_3d(sb, &tmBh, fb, kb * KBE, n * p.BN, z);
            if (!BF16) {
              tma_load_3d(sa + kATileF, &tmAl, fb, kb * KBE, m_idx * kBM, 0);
              tma_load_3d(sb + kBTileF, &tmBl, fb, kb * KBE, n * p.BN, z);
            }
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if ((p.debug & 8) && blockIdx.x == 0)
        printf("[fused cta0] TMA producer: total %lld cyc, waiting for an empty stage %lld\n", clock64() - t_begin, w_empty);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(kBM >> 4) << 24);
      uint32_t stage = 0, phase = 0, it = 0;
      long long w_tempty = 0, w_full = 0;
      const long long t_begin = clock64();
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        for (int n = 0; n < p.n_tiles; ++n, ++it) {
          const uint32_t as = it & 1u;
          long long t0 = clock64();
          mbar_wait(tempty0 + 8 * as, ((it >> 1) & 1u) ^ 1u);
          w_tempty += clock64() - t0;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * (uint32_t)kBNMax;
          uint32_t accumulate = 0;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            t0 = clock64();
            mbar_wait(full0 + 8 * stage, phase);
            w_full += clock64() - t0;
            tc_fence_after();
            const uint32_t sa = ring + stage * STAGE;
            const uint32_t sb = sa + NARR * kATileF;
            const uint64_t a_hi = smem_desc<KBB>(sa), b_hi = smem_desc<KBB>(sb);
            if (BF16) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                tc_mma<true>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, accumulate);
                accumulate = 1;
              }
            } else {
              const uint64_t a_lo = smem_desc<KBB>(sa + kATileF), b_lo = smem_desc<KBB>(sb + kBTileF);
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                tc_mma<F16K>(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, accumulate);
                tc_mma<F16K>(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                tc_mma<F16K>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
                accumulate = 1;
              }
            }
            tc_commit(empty0 + 8 * stage);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
          }
          tc_commit(tfull0 + 8 * as);
        }
      }
      if ((p.debug & 8) && blockIdx.x == 0)
        printf("[fused cta0] MMA issuer: total %lld cyc, waiting for a free accumulator %lld, for operands %lld\n",
               clock64() - t_begin, w_tempty, w_full);
    }
  } else {
    // ===================== epilogue: warps 2..17 =====================
    // TMEM is read with the 16x256b shape: lane t of the warp receives, for every group of 8 columns, rows
    // {t/4, t/4+8} x columns {2(t%4), 2(t%4)+1} -- the mma.sync accumulator fragment layout.  That is also the
    // A-fragment layout of mma.sync.m16n8k16 (two adjacent column groups = one 16x16 A tile), so the two small
    // GEMMs of the head run on the tensor cores straight from registers:
    //   pass 1  logits[16 rows, classes] += act[16, 16 hidden] . Wo^T      (B fragments: ldmatrix of Wo16[c][j])
    //   pass 2  dH[16 rows, 8 hidden]     = dlogits[16, classes] . Wo      (B fragments: ldmatrix.trans of Wo16[c][j])
    // with every fp32 operand split into fp16 hi/lo (power-of-two scaled per row / per sample) and the usual three
    // products, fp32 accumulation.  This replaces 20 FFMA + 5 LDS per hidden unit by 1.5 HMMA + 0.5 LDSM per 16 of them.
    // warp = (TMEM lane quadrant, 16-lane half of the quadrant, column half of the n-tile).
    const int ew = warp - 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int lhalf = (ew >> 2) & 1;           // which 16 lanes of the quadrant
    const int half = ew >> 3;                  // which half of an n-tile's columns
    const int et = threadIdx.x - 64;           // 0..511
    const int cols_half = p.BN >= 64 ? p.BN / 2 : p.BN;   // columns per (n-tile, half); BN < 64: half 1 idles
    const bool active = p.BN >= 64 || half == 0;
    const int nblocks = cols_half / 32;        // 32-column blocks per (n-tile, half)
    const int q = lane & 3, rsub = lane >> 2;  // fragment coordinates
    const int C = p.C, H = p.H;
    const int ldw = H + 8;                     // halfs per row of the staged Wo copies (+8: conflict-free ldmatrix)
    const float unscale = (MODE == MODE_F16X3) ? __ldg(p.unscale) : 1.f;     // 1 / (s_X s_W1)
    const float dh_scale = (MODE == MODE_F16X3) ? __ldg(p.dh_scale) : 1.f;   // dH -> fp16 range
    // ldmatrix row addresses of this lane (matrix = lane / 8, row = lane % 8); class rows >= C read the zero row
    const int lm = lane >> 3, lr = lane & 7;
    const int c1 = min((lm >> 1) * 8 + lr, C), c2 = min((lm & 1) * 8 + lr, C);
    const uint32_t off1 = (uint32_t)(c1 * ldw + (lm & 1) * 8) * 2u;   // pass 1: matrices (classes 0-7 | 8-15) x (k 0-7 | 8-15)
    const uint32_t off2 = (uint32_t)(c2 * ldw + (lm >> 1) * 8) * 2u;  // pass 2: matrices (classes 0-7 | 8-15) x (8 columns | next 8)
    const uint32_t wo_hi_a = smem_u32(wo_hi), wo_lo_a = smem_u32(wo_lo);
    const int cls[4] = {2 * q, 2 * q + 1, 8 + 2 * q, 9 + 2 * q};      // classes of this thread's 4 accumulator slots per row
    for (int i = et; i < ldw; i += kEpiWarps * 32) {                 // the zero row
      wo_hi[C * ldw + i] = __float2half_rn(0.f);
      wo_lo[C * ldw + i] = __float2half_rn(0.f);
    }
    uint32_t it = 0;
    long long w_tfull = 0, c_stage = 0, c_p1 = 0, c_head = 0, c_p2 = 0, t_mark = clock64();
    const long long t_begin = t_mark;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int z = item / p.m_tiles, m_idx = item % p.m_tiles;
      const float* __restrict__ wrow = p.bank + (long long)(p.z_row0 + z) * p.P;
      { const long long t = clock64(); c_p2 += t - t_mark; t_mark = t; }
      // ---------------- stage Wo_z (fp16 hi/lo of s_wo * Wo), b1_z, bo_z ----------------
      float wv[kWoPerThread];
      float wmax = 0.f, bmax = 0.f;
#pragma unroll
      for (int u = 0; u < kWoPerThread; ++u) {
        const int i = et + u * kEpiWarps * 32;
        wv[u] = i < C * H ? __ldg(wrow + p.wo_off + i) : 0.f;
        wmax = fmaxf(wmax, fabsf(wv[u]));
      }
      float b1v = 0.f;
      if (et < H) { b1v = __ldg(wrow + p.b1_off + et); bmax = fabsf(b1v); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
      }
      epi_bar();                                // everyone is done with the previous item's parameters
      if (lane == 0) { red[ew] = wmax; red[kEpiWarps + ew] = bmax; }
      if (et < H) b1_s[et] = b1v;
      if (et < C) bo_s[et] = __ldg(wrow + p.bo_off + et);
      if (et == 0) *wl_count = 0u;
      epi_bar();
      wmax = 0.f; bmax = 0.f;
#pragma unroll
      for (int i = 0; i < kEpiWarps; ++i) { wmax = fmaxf(wmax, red[i]); bmax = fmaxf(bmax, red[kEpiWarps + i]); }
      const float s_wo = pow2_scale(wmax);
#pragma unroll
      for (int u = 0; u < kWoPerThread; ++u) {
        const int i = et + u * kEpiWarps * 32;
        if (i < C * H) {
          const int c = i / H, j = i - c * H;
          const float v = wv[u] * s_wo;
          const __half h = __float2half_rn(v);
          wo_hi[c * ldw + j] = h;
          wo_lo[c * ldw + j] = __float2half_rn(v - __half2float(h));
        }
      }
      epi_bar();
      // this thread's 2 rows: quad*32 + lhalf*16 + {0,8} + rsub
      int brow[2];
      bool rok[2];
      float guard[2], s_a[2];
      const float wn = __ldg(p.wnorm + p.z_row0 + z);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        brow[r] = m_idx * kBM + quad * 32 + lhalf * 16 + 8 * r + rsub;
        rok[r] = brow[r] < p.B;
        const float bound = rok[r] ? wn * __ldg(p.xnorm + brow[r]) : 0.f;   // |<x_b, w_zj>| <= ||x_b|| max_j ||w_zj||
        guard[r] = p.eps * bound;
        s_a[r] = pow2_scale(bound + bmax);     // activations of this row -> [.., 2^9)
      }
      unsigned long long* wl = p.worklist ? p.worklist + (long long)item * kWorkPerItem : nullptr;
      float acc[2][4];                          // logits: [class group][row 0: c, c+1 | row 1: c, c+1], scaled by s_a[row] s_wo
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[g][i] = 0.f;
      uint32_t mbits[kMaskWords];               // one word per 16x32 block: bit (r*8 + k*2 + e)
#pragma unroll
      for (int i = 0; i < kMaskWords; ++i) mbits[i] = 0u;

      { const long long t = clock64(); c_stage += t - t_mark; t_mark = t; }
      // ---------------- pass 1 ----------------
      for (int n = 0; n < p.n_tiles; ++n, ++it) {
        const uint32_t as = it & 1u;
        { const long long t0 = clock64(); mbar_wait(tfull0 + 8 * as, (it >> 1) & 1u); w_tfull += clock64() - t0; }
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (uint32_t)kBNMax;
        if (active) {
#pragma unroll 1
          for (int cc = 0; cc < nblocks; ++cc) {
            const int c0 = half * cols_half + cc * 32;            // column inside the n-tile
            uint32_t v[16];                                       // [4 * colgroup + {r0c0, r0c1, r1c0, r1c1}]
            tmem_ld_16x256b_x4(taddr + ((uint32_t)(lhalf * 16) << 16) + (uint32_t)c0, v);
            // B fragments of the two 16-column K-blocks while the TMEM load is in flight
            const uint32_t jb = (uint32_t)(n * p.BN + c0) * 2u;
            uint32_t bh[2][4], bl[2][4];
            ldsm_x4(wo_hi_a + off1 + jb, bh[0]);
            ldsm_x4(wo_lo_a + off1 + jb, bl[0]);
            ldsm_x4(wo_hi_a + off1 + jb + 32u, bh[1]);
            ldsm_x4(wo_lo_a + off1 + jb + 32u, bl[1]);
            tmem_ld_wait();
            if (p.debug & 4) continue;
            uint32_t bits = 0u;
            const int jbase = n * p.BN + c0 + 2 * q;
            uint32_t ahi[4][2], alo[4][2];                        // [column group k][row]: fp16 pairs of the activations
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int j = jbase + 8 * k;                        // this thread's column pair (j, j+1)
              const float2 bb = *reinterpret_cast<const float2*>(b1_s + j);
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                float h[2];
                if (MODE == MODE_F16X3) {
                  h[0] = fmaf(__uint_as_float(v[4 * k + 2 * r]), unscale, bb.x);
                  h[1] = fmaf(__uint_as_float(v[4 * k + 2 * r + 1]), unscale, bb.y);
                } else {
                  h[0] = __uint_as_float(v[4 * k + 2 * r]) + bb.x;
                  h[1] = __uint_as_float(v[4 * k + 2 * r + 1]) + bb.y;
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  bool pos = h[e] > 0.f;
                  if (fabsf(h[e]) < guard[r]) {
                    const uint32_t slot = atomicAdd(wl_count, 1u);
                    if (slot < (uint32_t)kWorkPerItem) {
                      wl[slot] = pack_entry(z, brow[r], j + e, pos);
                    } else {                                      // item budget exhausted: settle it here
                      pos = exact_positive_serial(p.x + (long long)brow[r] * p.D, wrow + (long long)(j + e) * p.D,
                                                  e ? bb.y : bb.x, p.D);
                    }
                  }
                  if (pos) bits |= 1u << (r * 8 + k * 2 + e);
                  h[e] = (pos ? h[e] : h[e] * kSlopeF) * s_a[r];
                }
                split_pair(h[0], h[1], ahi[k][r], alo[k][r]);
              }
            }
            const int word = n * nblocks + cc;
#pragma unroll
            for (int i = 0; i < kMaskWords; ++i)
              if (i == word) mbits[i] = bits;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {                      // 16 hidden units each
              const uint32_t Ah[4] = {ahi[2 * kb][0], ahi[2 * kb][1], ahi[2 * kb + 1][0], ahi[2 * kb + 1][1]};
              const uint32_t Al[4] = {alo[2 * kb][0], alo[2 * kb][1], alo[2 * kb + 1][0], alo[2 * kb + 1][1]};
#pragma unroll
              for (int g = 0; g < 2; ++g) {                       // classes 0-7 | 8-15
                if (g == 1 && C <= 8) continue;
                hmma_16816(acc[g], Al, bh[kb][2 * g], bh[kb][2 * g + 1]);
                hmma_16816(acc[g], Ah, bl[kb][2 * g], bl[kb][2 * g + 1]);
                hmma_16816(acc[g], Ah, bh[kb][2 * g], bh[kb][2 * g + 1]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      }

      { const long long t = clock64(); c_p1 += t - t_mark; t_mark = t; }
      // ---------------- logits: sum of the two column halves, then the loss head on this thread's 2 rows x 4 classes ----------------
      float dl[2][4];                           // [row][class slot]: logits -> dlogits
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float inv = 1.f / (s_a[r] * s_wo);
        dl[r][0] = acc[0][2 * r] * inv; dl[r][1] = acc[0][2 * r + 1] * inv;
        dl[r][2] = acc[1][2 * r] * inv; dl[r][3] = acc[1][2 * r + 1] * inv;
      }
      const int rit0 = quad * 32 + lhalf * 16 + rsub;
      if (half == 1) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int i = 0; i < 4; ++i) xchg[(rit0 + 8 * r) * 16 + cls[i]] = dl[r][i];
      }
      epi_bar();
      if (half == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int rit = rit0 + 8 * r;
          bool valid[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            valid[i] = cls[i] < C;
            dl[r][i] += (p.BN >= 64 ? xchg[rit * 16 + cls[i]] : 0.f) + (valid[i] ? bo_s[cls[i]] : 0.f);
          }
          if (p.head < 0) {
            if (rok[r]) {
              float* out = p.logits + ((long long)z * p.B + brow[r]) * C;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (valid[i]) out[cls[i]] = dl[r][i];
            }
          } else {
            head_quad(p.head, dl[r], valid, cls, rok[r] ? p.labels[brow[r]] : 0,
                      (rok[r] && p.pbar) ? p.pbar + (long long)brow[r] * C : nullptr);
#pragma unroll
            for (int i = 0; i < 4; ++i) xchg[rit * 16 + cls[i]] = dl[r][i];
          }
        }
      }
      epi_bar();
      // sentinel-fill the unused worklist slots of this item
      if (wl) {
        const uint32_t used = min(*wl_count, (uint32_t)kWorkPerItem);
        for (uint32_t i = used + et; i < (uint32_t)kWorkPerItem; i += kEpiWarps * 32) wl[i] = kSentinel;
      }
      if (p.head == -2 && active) {            // keep mode: the LeakyReLU masks of this item for the gradient pass
        uint32_t* mb = p.maskbuf + (long long)item * (kMaskWords * kEpiWarps * 32);
#pragma unroll
        for (int i = 0; i < kMaskWords; ++i)
          if (i < p.n_tiles * nblocks) mb[i * (kEpiWarps * 32) + et] = mbits[i];
      }
      if (p.head < 0) continue;
      if (half == 1) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int i = 0; i < 4; ++i) dl[r][i] = xchg[(rit0 + 8 * r) * 16 + cls[i]];
      }

      { const long long t = clock64(); c_head += t - t_mark; t_mark = t; }
      // ---------------- pass 2: dH = (dlogits . Wo) * leaky'(H) for this thread's 2 rows x column pairs ----------------
      if (active && !(p.debug & 2)) {
        uint32_t Dh[4], Dl[4];
        float mul[2];
        long long orow[2];
        dlogits_frags(dl, s_wo, dh_scale, Dh, Dl, mul);
#pragma unroll
        for (int r = 0; r < 2; ++r) orow[r] = ((long long)z * p.B + brow[r]) * H;
        for (int n = 0; n < p.n_tiles; ++n) {
#pragma unroll 1
          for (int cc = 0; cc < nblocks; ++cc) {
            const int j0 = n * p.BN + half * cols_half + cc * 32;
            const int word = n * nblocks + cc;
            uint32_t bits = 0u;
#pragma unroll
            for (int i = 0; i < kMaskWords; ++i)
              if (i == word) bits = mbits[i];
            pass2_block<MODE>(Dh, Dl, mul, bits, wo_hi_a, wo_lo_a, off2, j0, q, rok, orow, p.dh_hi, p.dh_lo, p.dh_bf,
                              (p.debug & 1) != 0);
          }
        }
      }
    }
    if ((p.debug & 8) && blockIdx.x == 0 && et == 0)
      printf("[fused cta0] epilogue warp 2: total %lld cyc: staging %lld, pass 1 %lld (of which waiting for the accumulator %lld), "
             "head %lld, pass 2 %lld\n", clock64() - t_begin, c_stage, c_p1, w_tfull, c_head, c_p2 + (clock64() - t_mark));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// exact sign of b1[j] + <x_b, w_j>: fp64 accumulation of the exact fp32 products, one warp (all lanes return it)
__device__ __forceinline__ bool exact_positive_warp(const float* __restrict__ xr, const float* __restrict__ w, float bias,
                                                    int D, int lane) {
  double s = 0.0;
  // bank rows are only 4-byte aligned in general (P floats apart), hence scalar loads.  (An 8-deep software pipeline
  // with two accumulation chains -- which halved the unfused route's refine_kernel -- made this kernel 40 % slower.)
  for (int d = lane; d < D; d += 32) s = fma((double)__ldg(xr + d), (double)__ldg(w + d), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return (float)(s + (double)bias) > 0.f;
}

// Worklist fix-up.  An item's guard-band entries sit in the FIRST slots of its kWorkPerItem-slot segment, so a
// slot-parallel sweep leaves one warp with all the work of an item: instead one block per item (grid-stride), every
// warp reads the segment 32 slots at a time and takes the entries whose index is congruent to its own.
//   MASK = false: dH[z, b, j] is rescaled by slope^(+-1) in place when the sign the epilogue assumed was wrong
//   MASK = true : keep mode, the stored LeakyReLU mask bit is flipped instead
template <bool F16, bool MASK>
__global__ void __launch_bounds__(256)
fixup_kernel(const unsigned long long* __restrict__ wl, int num_items, const float* __restrict__ x,
             const float* __restrict__ bank, long long P, long long b1_off, int z_row0, int B, int D, int H,
             void* __restrict__ dh_hi_v, void* __restrict__ dh_lo_v, int BN, int cols_half, uint32_t* __restrict__ masks) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m_tiles = (B + kBM - 1) / kBM, nblocks = cols_half / 32;
  for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
    const unsigned long long* __restrict__ seg = wl + (long long)item * kWorkPerItem;
    for (int base = 0; base < kWorkPerItem; base += 32) {
      const unsigned long long mine = seg[base + lane];
      unsigned valid = __ballot_sync(0xffffffffu, mine != kSentinel);
      if (!valid) break;                                     // entries are packed from slot 0; the rest is sentinels
      while (valid) {
        const int src = __ffs(valid) - 1;
        valid &= valid - 1;
        if (((base + src) & 7) != warp) continue;
        const unsigned long long e = __shfl_sync(0xffffffffu, mine, src);
        const int z = (int)(e >> 44), b = (int)((e >> 20) & 0xFFFFFF), j = (int)((e >> 4) & 0xFFFF);
        const bool assumed_pos = (e & 1ull) != 0;
        const float* __restrict__ wrow = bank + (long long)(z_row0 + z) * P;
        const bool pos = exact_positive_warp(x + (long long)b * D, wrow + (long long)j * D, __ldg(wrow + b1_off + j), D, lane);
        if (pos == assumed_pos || lane != 0) continue;
        if (MASK) {
          // (b, j) -> (item, mask word, epilogue thread, bit) of the fused kernel's fragment layout
          const int rt = b % kBM;
          const int quad = rt >> 5, lhalf = (rt >> 4) & 1, r = (rt >> 3) & 1, rsub = rt & 7;
          const int n = j / BN, jn = j % BN, half = jn / cols_half, jc = jn % cols_half;
          const int cc = jc >> 5, k = (jc >> 3) & 3, qq = (jc >> 1) & 3, ee = jc & 1;
          const int ew = half * 8 + lhalf * 4 + ((quad + 2) & 3);
          const int et = ew * 32 + rsub * 4 + qq;
          atomicXor(masks + ((long long)(z * m_tiles + b / kBM) * kMaskWords + (n * nblocks + cc)) * (kEpiWarps * 32) + et,
                    1u << (r * 8 + k * 2 + ee));
        } else {
          const long long o = ((long long)z * B + b) * H + j;
          if (F16) {                                   // scaled fp16 hi/lo pair (the scale is a power of two: it commutes)
            __half* dh_hi = reinterpret_cast<__half*>(dh_hi_v);
            __half* dh_lo = reinterpret_cast<__half*>(dh_lo_v);
            float v = __half2float(dh_hi[o]) + __half2float(dh_lo[o]);
            v = pos ? v * 100.f : v * kSlopeF;
            const __half hi = __float2half_rn(v);
            dh_hi[o] = hi;
            dh_lo[o] = __float2half_rn(v - __half2float(hi));
          } else {
            float* dh_hi = reinterpret_cast<float*>(dh_hi_v);
            float* dh_lo = reinterpret_cast<float*>(dh_lo_v);
            float v = dh_hi[o] + dh_lo[o];
            v = pos ? v * 100.f : v * kSlopeF;          // undo / apply the LeakyReLU slope
            const float hi = to_tf32_rn(v);
            dh_hi[o] = hi;
            dh_lo[o] = v - hi;
          }
        }
      }
    }
  }
}

// Gradient pass of a kept forward (attacks: adversarialAttacks.py:74-78 evaluates BNN.forward once and differentiates it):
// dH of every (sample, input) from the stored logits and LeakyReLU masks -- the loss head and pass 2 of the fused
// kernel without its GEMM.  One CTA of 16 warps per work item at a time, same thread <-> (row, column) mapping as the
// fused epilogue (the mask words are stored per epilogue thread).
template <int MODE>
__global__ void __launch_bounds__(kEpiWarps * 32, 2)
dh_from_kept_kernel(int B, int H, int C, int num_items, int m_tiles, int head, const float* __restrict__ bank, long long P,
                    long long wo_off, int z_row0, const int32_t* __restrict__ labels, const float* __restrict__ pbar,
                    const float* __restrict__ logits, const uint32_t* __restrict__ masks, void* dh_hi, void* dh_lo,
                    __nv_bfloat16* dh_bf, const float* __restrict__ dh_scale_p) {
  __shared__ __align__(16) __half wo16[2 * kWoHalfs];
  __shared__ float red[kEpiWarps];
  __half* wo_hi = wo16;
  __half* wo_lo = wo16 + kWoHalfs;
  const int et = threadIdx.x, ew = et >> 5, lane = et & 31;
  const int quad = (ew + 2) & 3, lhalf = (ew >> 2) & 1, half = ew >> 3;
  const int BN = H <= 256 ? H : 256, n_tiles = H <= 256 ? 1 : H / 256;
  const int cols_half = BN >= 64 ? BN / 2 : BN;
  const bool active = BN >= 64 || half == 0;
  const int nblocks = cols_half / 32;
  const int q = lane & 3, rsub = lane >> 2;
  const int ldw = H + 8;
  const float dh_scale = (MODE == MODE_F16X3) ? __ldg(dh_scale_p) : 1.f;
  const int lm = lane >> 3, lr = lane & 7;
  const int c2 = min((lm & 1) * 8 + lr, C);
  const uint32_t off2 = (uint32_t)(c2 * ldw + (lm >> 1) * 8) * 2u;
  const uint32_t wo_hi_a = smem_u32(wo_hi), wo_lo_a = smem_u32(wo_lo);
  const int cls[4] = {2 * q, 2 * q + 1, 8 + 2 * q, 9 + 2 * q};
  for (int i = et; i < ldw; i += kEpiWarps * 32) {
    wo_hi[C * ldw + i] = __float2half_rn(0.f);
    wo_lo[C * ldw + i] = __float2half_rn(0.f);
  }
  for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
    const int z = item / m_tiles, m_idx = item % m_tiles;
    const float* __restrict__ wrow = bank + (long long)(z_row0 + z) * P;
    float wv[kWoPerThread];
    float wmax = 0.f;
#pragma unroll
    for (int u = 0; u < kWoPerThread; ++u) {
      const int i = et + u * kEpiWarps * 32;
      wv[u] = i < C * H ? __ldg(wrow + wo_off + i) : 0.f;
      wmax = fmaxf(wmax, fabsf(wv[u]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    __syncthreads();                          // everyone is done with the previous item's Wo16
    if (lane == 0) red[ew] = wmax;
    __syncthreads();
    wmax = 0.f;
#pragma unroll
    for (int i = 0; i < kEpiWarps; ++i) wmax = fmaxf(wmax, red[i]);
    const float s_wo = pow2_scale(wmax);
#pragma unroll
    for (int u = 0; u < kWoPerThread; ++u) {
      const int i = et + u * kEpiWarps * 32;
      if (i < C * H) {
        const int c = i / H, j = i - c * H;
        const float v = wv[u] * s_wo;
        const __half h = __float2half_rn(v);
        wo_hi[c * ldw + j] = h;
        wo_lo[c * ldw + j] = __float2half_rn(v - __half2float(h));
      }
    }
    __syncthreads();
    int brow[2];
    bool rok[2];
    float dl[2][4];
    long long orow[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      brow[r] = m_idx * kBM + quad * 32 + lhalf * 16 + 8 * r + rsub;
      rok[r] = brow[r] < B;
      bool valid[4];
      const float* __restrict__ lrow = logits + ((long long)z * B + brow[r]) * C;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        valid[i] = cls[i] < C;
        dl[r][i] = (valid[i] && rok[r]) ? __ldg(lrow + cls[i]) : 0.f;
      }
      head_quad(head, dl[r], valid, cls, rok[r] ? labels[brow[r]] : 0,
                (rok[r] && pbar) ? pbar + (long long)brow[r] * C : nullptr);
      orow[r] = ((long long)z * B + brow[r]) * H;
    }
    if (!active) continue;
    uint32_t Dh[4], Dl[4];
    float mul[2];
    dlogits_frags(dl, s_wo, dh_scale, Dh, Dl, mul);
    const uint32_t* __restrict__ mb = masks + (long long)item * (kMaskWords * kEpiWarps * 32) + et;
    for (int n = 0; n < n_tiles; ++n) {
#pragma unroll 1
      for (int cc = 0; cc < nblocks; ++cc) {
        const int j0 = n * BN + half * cols_half + cc * 32;
        const uint32_t bits = __ldg(mb + (n * nblocks + cc) * (kEpiWarps * 32));
        pass2_block<MODE>(Dh, Dl, mul, bits, wo_hi_a, wo_lo_a, off2, j0, q, rok, orow, dh_hi, dh_lo, dh_bf, false);
      }
    }
  }
}

}  // namespace

bool fused_supported(int H, int C) {
  if (C > kCMax || H < 32 || (H & 15)) return false;
  if (H > 256 && (H % 256)) return false;
  const int n_tiles = H <= 256 ? 1 : H / 256;
  const int bn = H <= 256 ? H : 256;
  const int cols_half = bn >= 64 ? bn / 2 : bn;
  if (cols_half % 32) return false;
  if (n_tiles * (cols_half / 32) > kMaskWords) return false;
  return H <= kHMax;
}

size_t fused_worklist_slots(int B, int Z) { return (size_t)Z * ((B + kBM - 1) / kBM) * kWorkPerItem; }

int fused_forward_head(const FusedDesc& d, cudaStream_t st, std::string* err) {
  std::string local;
  if (!err) err = &local;
  const bool bf16 = d.mode == MODE_BF16, f16x3 = d.mode == MODE_F16X3;
  const int dt = bf16 ? DT_BF16 : (f16x3 ? DT_F16 : DT_F32);
  if (d.B <= 0 || d.Z <= 0) return 0;
  if (!fused_supported(d.H, d.C)) { *err = "fused_forward_head: unsupported hidden / class size"; return 1; }
  if (d.mode < 0 || d.mode > MODE_F16X3) { *err = "fused_forward_head: unknown mode"; return 1; }
  if (f16x3 && (!d.unscale || !d.dh_scale)) { *err = "fused_forward_head: F16X3 needs the scale scalars"; return 1; }
  const int kbb = d.kblock_bytes == 128 ? 128 : 64;
  typedef void (*kern_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const FParams);
  static const kern_t kerns[3][2] = {{fc_fused_kernel<MODE_TF32X3, 64>, fc_fused_kernel<MODE_TF32X3, 128>},
                                     {fc_fused_kernel<MODE_BF16, 64>, fc_fused_kernel<MODE_BF16, 128>},
                                     {fc_fused_kernel<MODE_F16X3, 64>, fc_fused_kernel<MODE_F16X3, 128>}};
  kern_t kern = kerns[d.mode][kbb == 128];
  static bool attr_done[3][2] = {};
  if (!attr_done[d.mode][kbb == 128]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem);
    if (e != cudaSuccess) { *err = std::string("fused_forward_head: cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    attr_done[d.mode][kbb == 128] = true;
  }
  FParams p{};
  p.B = d.B; p.D = d.D; p.H = d.H; p.C = d.C; p.Z = d.Z;
  p.BN = d.H <= 256 ? d.H : 256;
  p.n_tiles = d.H <= 256 ? 1 : d.H / 256;
  p.m_tiles = (d.B + kBM - 1) / kBM;
  p.num_items = p.m_tiles * d.Z;
  const int kbe = kbb / (dt == DT_F32 ? 4 : 2);
  p.num_kb = (d.D + kbe - 1) / kbe;
  p.head = d.head;
  p.bank = d.bank; p.P = d.P; p.b1_off = d.b1_off; p.wo_off = d.wo_off; p.bo_off = d.bo_off; p.z_row0 = d.z_row0;
  p.labels = d.labels; p.pbar = d.pbar;
  p.x = d.x; p.xnorm = d.xnorm; p.wnorm = d.wnorm;
  p.eps = (bf16 || !d.worklist || d.head == -1) ? 0.f : d.eps;      // keep mode (-2) needs exact masks too
  p.maskbuf = d.maskbuf;
  p.dh_hi = d.dh_hi; p.dh_lo = d.dh_lo; p.dh_bf = reinterpret_cast<__nv_bfloat16*>(d.dh_bf); p.logits = d.logits;
  p.worklist = p.eps > 0.f ? d.worklist : nullptr;
  p.unscale = d.unscale; p.dh_scale = d.dh_scale;
  {
    static const int dbg = getenv("RBNN_FUSED_DEBUG") ? atoi(getenv("RBNN_FUSED_DEBUG")) : 0;
    p.debug = dbg;
  }
  if (d.head >= 0 && (bf16 ? !d.dh_bf : (!d.dh_hi || !d.dh_lo))) { *err = "fused_forward_head: missing dH output"; return 1; }
  if (d.head < 0 && !d.logits) { *err = "fused_forward_head: missing logits output"; return 1; }
  if (d.head == -2 && !d.maskbuf) { *err = "fused_forward_head: keep mode without a mask buffer"; return 1; }
  if (d.head >= 0 && d.head != RBNN_HEAD_MEAN_OF_GRADS && d.head != RBNN_HEAD_LOGITS_CE && !d.pbar) {
    *err = "fused_forward_head: this head needs pbar";
    return 1;
  }
  CUtensorMap mAh, mAl, mBh, mBl;
  if (make_map(&mAh, d.X.hi, dt, d.D, d.B, 1, d.X.ld, 0, kBM, kbb, err)) return 1;
  if (make_map(&mBh, d.W1.hi, dt, d.D, d.H, d.Z, d.W1.ld, d.W1.zstride, p.BN, kbb, err)) return 1;
  if (!bf16) {
    if (make_map(&mAl, d.X.lo, dt, d.D, d.B, 1, d.X.ld, 0, kBM, kbb, err)) return 1;
    if (make_map(&mBl, d.W1.lo, dt, d.D, d.H, d.Z, d.W1.ld, d.W1.zstride, p.BN, kbb, err)) return 1;
  } else {
    mAl = mAh;
    mBl = mBh;
  }
  const int grid = p.num_items < d.sm_count ? p.num_items : d.sm_count;
  kern<<<grid, kThreadsF, kFusedSmem, st>>>(mAh, mAl, mBh, mBl, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("fused_forward_head launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

int fused_fixup(const FusedDesc& d, cudaStream_t st, std::string* err) {
  if (d.mode == MODE_BF16 || !d.worklist || d.head < 0 || d.eps <= 0.f || d.B <= 0 || d.Z <= 0) return 0;
  const int items = d.Z * ((d.B + kBM - 1) / kBM);
  const unsigned blocks = (unsigned)std::min(items, d.sm_count * 8);
  if (d.mode == MODE_F16X3)
    fixup_kernel<true, false><<<blocks, 256, 0, st>>>(d.worklist, items, d.x, d.bank, d.P, d.b1_off, d.z_row0, d.B, d.D, d.H,
                                                      d.dh_hi, d.dh_lo, 0, 32, nullptr);
  else
    fixup_kernel<false, false><<<blocks, 256, 0, st>>>(d.worklist, items, d.x, d.bank, d.P, d.b1_off, d.z_row0, d.B, d.D, d.H,
                                                       d.dh_hi, d.dh_lo, 0, 32, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("fused_fixup launch: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

size_t keep_mask_words(int B, int Z) { return (size_t)Z * ((B + kBM - 1) / kBM) * kMaskWords * kEpiWarps * 32; }

int fused_keep_fixup(const FusedDesc& d, uint32_t* masks, cudaStream_t st, std::string* err) {
  if (d.mode == MODE_BF16 || !d.worklist || d.eps <= 0.f || d.B <= 0 || d.Z <= 0) return 0;
  const int items = d.Z * ((d.B + kBM - 1) / kBM);
  const unsigned blocks = (unsigned)std::min(items, d.sm_count * 8);
  const int BN = d.H <= 256 ? d.H : 256;
  fixup_kernel<false, true><<<blocks, 256, 0, st>>>(d.worklist, items, d.x, d.bank, d.P, d.b1_off, d.z_row0, d.B, d.D, d.H,
                                                    nullptr, nullptr, BN, BN >= 64 ? BN / 2 : BN, masks);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("fused_keep_fixup launch: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

int dh_from_kept(const KeptDesc& d, cudaStream_t st, std::string* err) {
  std::string local;
  if (!err) err = &local;
  if (d.B <= 0 || d.Z <= 0) return 0;
  if (!fused_supported(d.H, d.C)) { *err = "dh_from_kept: unsupported hidden / class size"; return 1; }
  if (d.head < 0 || !d.logits || !d.masks) { *err = "dh_from_kept: missing kept forward"; return 1; }
  if (d.mode == MODE_F16X3 && !d.dh_scale) { *err = "dh_from_kept: F16X3 needs the dH scale"; return 1; }
  const int m_tiles = (d.B + kBM - 1) / kBM, items = m_tiles * d.Z;
  const int grid = std::min(items, d.sm_count * 2);
  __nv_bfloat16* bf = reinterpret_cast<__nv_bfloat16*>(d.dh_bf);
#define RBNN_KEPT_LAUNCH(M)                                                                                          \
  dh_from_kept_kernel<M><<<grid, kEpiWarps * 32, 0, st>>>(d.B, d.H, d.C, items, m_tiles, d.head, d.bank, d.P, d.wo_off, \
                                                          d.z_row0, d.labels, d.pbar, d.logits, d.masks, d.dh_hi, d.dh_lo, \
                                                          bf, d.dh_scale)
  if (d.mode == MODE_F16X3) RBNN_KEPT_LAUNCH(MODE_F16X3);
  else if (d.mode == MODE_BF16) RBNN_KEPT_LAUNCH(MODE_BF16);
  else RBNN_KEPT_LAUNCH(MODE_TF32X3);
#undef RBNN_KEPT_LAUNCH
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("dh_from_kept launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace tc
}  // namespace rbnn
