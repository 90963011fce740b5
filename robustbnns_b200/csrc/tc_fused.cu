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
//   warp 1      tcgen05.mma issuer (kind::tf32 x3 passes or kind::f16)
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

#include "../../include/rbnn.h"
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace rbnn {
namespace tc {

namespace {

constexpr float kSlopeF = 0.01f;
constexpr int kRingF = 196608;
constexpr int kEpiWarps = 8;
constexpr int kThreadsF = 64 + kEpiWarps * 32;     // 320
constexpr int kCMax = 16;
constexpr int kParamFloatsMax = 5760;              // Wo_z [C*H] + bo_z [C] + b1_z [H]  (<= 22.5 KB)
constexpr int kXchgFloats = kBM * kCMax;           // partial logits / dlogits exchange between column halves
constexpr int kFusedSmem = kRingF + 1024 + 256 + (kParamFloatsMax + kXchgFloats) * 4 + 16;
constexpr unsigned long long kSentinel = ~0ull;

struct FParams {
  int B, D, H, C, Z, BN, n_tiles, m_tiles, num_items, num_kb;
  int head;
  const float* bank; long long P, b1_off, wo_off, bo_off; int z_row0;
  const int32_t* labels; const float* pbar;
  const float* x; const float* xnorm; const float* wnorm; float eps;
  float* dh_hi; float* dh_lo; __nv_bfloat16* dh_bf; float* logits;
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

template <bool BF16, int KBB>
__global__ void __launch_bounds__(kThreadsF, 1)
fc_fused_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                const FParams p) {
  constexpr int NARR = BF16 ? 1 : 2;
  constexpr int kATileF = kBM * KBB, kBTileF = kBNMax * KBB;
  constexpr int STAGE = NARR * (kATileF + kBTileF);
  constexpr int NSTAGE = kRingF / STAGE;
  constexpr int KBE = KBB / (BF16 ? 2 : 4);
  constexpr int KSTEPS = KBB / 32;
  constexpr uint32_t FMT = BF16 ? 1u : 2u;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  const uint32_t bars = ring + kRingF;
  const uint32_t full0 = bars, empty0 = bars + 8 * NSTAGE;
  const uint32_t tfull0 = bars + 16 * NSTAGE, tempty0 = tfull0 + 16;
  uint8_t* gen = smem_raw + (bars - raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 16 * NSTAGE + 32);
  uint32_t* wl_count = reinterpret_cast<uint32_t*>(gen + 16 * NSTAGE + 40);
  float* wo_s = reinterpret_cast<float*>(gen + 256);      // [C][H]
  float* bo_s = wo_s + p.C * p.H;                          // [C]
  float* b1_s = wo_s + ((p.C * p.H + p.C + 3) & ~3);       // [H], 16-byte aligned
  float* xchg = wo_s + kParamFloatsMax;                    // [128][kCMax]

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
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int z = item / p.m_tiles, m_idx = item % p.m_tiles;
        for (int n = 0; n < p.n_tiles; ++n) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1u);
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
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(kBM >> 4) << 24);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        for (int n = 0; n < p.n_tiles; ++n, ++it) {
          const uint32_t as = it & 1u;
          mbar_wait(tempty0 + 8 * as, ((it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * (uint32_t)kBNMax;
          uint32_t accumulate = 0;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(full0 + 8 * stage, phase);
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
                tc_mma<false>(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, accumulate);
                tc_mma<false>(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                tc_mma<false>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
                accumulate = 1;
              }
            }
            tc_commit(empty0 + 8 * stage);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
          }
          tc_commit(tfull0 + 8 * as);
        }
      }
    }
  } else {
    // ===================== epilogue: warps 2..9 =====================
    const int ew = warp - 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                  // which half of an n-tile's columns
    const int et = threadIdx.x - 64;           // 0..255
    const int cols_half = p.BN >= 64 ? p.BN / 2 : p.BN;   // columns per (n-tile, half); BN < 64: half 1 idles
    const bool active = p.BN >= 64 || half == 0;
    const int row_in_tile = quad * 32 + lane;
    const int C = p.C, H = p.H;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int z = item / p.m_tiles, m_idx = item % p.m_tiles;
      const float* __restrict__ wrow = p.bank + (long long)(p.z_row0 + z) * p.P;
      epi_bar();                                // everyone is done with the previous item's parameters
      for (int i = et; i < C * H; i += kEpiWarps * 32) wo_s[i] = __ldg(wrow + p.wo_off + i);
      for (int i = et; i < C; i += kEpiWarps * 32) bo_s[i] = __ldg(wrow + p.bo_off + i);
      for (int i = et; i < H; i += kEpiWarps * 32) b1_s[i] = __ldg(wrow + p.b1_off + i);
      if (et == 0) *wl_count = 0u;
      epi_bar();
      const int b = m_idx * kBM + row_in_tile;
      const bool row_ok = b < p.B;
      const float guard = (row_ok && p.eps > 0.f) ? p.eps * __ldg(p.xnorm + b) * __ldg(p.wnorm + p.z_row0 + z) : 0.f;
      unsigned long long* wl = p.worklist ? p.worklist + (long long)item * kWorkPerItem : nullptr;
      float logit[kCMax];
#pragma unroll
      for (int c = 0; c < kCMax; ++c) logit[c] = 0.f;
      uint32_t mbits[16];                       // mask bits of this thread's columns: n-tile n, chunk cc -> word
#pragma unroll
      for (int i = 0; i < 16; ++i) mbits[i] = 0u;

      // ---------------- pass 1 ----------------
      for (int n = 0; n < p.n_tiles; ++n, ++it) {
        const uint32_t as = it & 1u;
        mbar_wait(tfull0 + 8 * as, (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (uint32_t)kBNMax;
        if (active) {
          const int nchunks = cols_half / 32 > 0 ? cols_half / 32 : 1;
#pragma unroll 1
          for (int cc = 0; cc < nchunks; ++cc) {
            const int c0 = half * cols_half + cc * 32;            // column inside the n-tile
            uint32_t r[32];
            tmem_ld32(taddr + (uint32_t)c0, r);
            uint32_t bits = 0u;
            const int jbase = n * p.BN + c0;
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              if (c0 + q < p.BN) {                                // BN % 16 == 0: groups of 4 are all-in / all-out
                const int j = jbase + q;
                const float4 bb = *reinterpret_cast<const float4*>(b1_s + j);
                float h[4] = {__uint_as_float(r[q]) + bb.x, __uint_as_float(r[q + 1]) + bb.y,
                              __uint_as_float(r[q + 2]) + bb.z, __uint_as_float(r[q + 3]) + bb.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  bool pos = h[e] > 0.f;
                  if (fabsf(h[e]) < guard) {
                    const uint32_t slot = atomicAdd(wl_count, 1u);
                    if (slot < (uint32_t)kWorkPerItem) {
                      wl[slot] = pack_entry(z, b, j + e, pos);
                    } else {                                      // item budget exhausted: settle it here
                      pos = exact_positive_serial(p.x + (long long)b * p.D, wrow + (long long)(j + e) * p.D,
                                                  b1_s[j + e], p.D);
                    }
                  }
                  if (pos) bits |= 1u << (q + e);
                  h[e] = pos ? h[e] : h[e] * kSlopeF;
                }
#pragma unroll
                for (int c = 0; c < kCMax; ++c)
                  if (c < C) {
                    const float4 w = *reinterpret_cast<const float4*>(wo_s + c * H + j);
                    logit[c] = fmaf(h[0], w.x, logit[c]);
                    logit[c] = fmaf(h[1], w.y, logit[c]);
                    logit[c] = fmaf(h[2], w.z, logit[c]);
                    logit[c] = fmaf(h[3], w.w, logit[c]);
                  }
              }
            }
            const int word = n * nchunks + cc;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i == word) mbits[i] = bits;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      }

      // ---------------- exchange: half 1 -> half 0 partial logits; half 0 -> half 1 dlogits ----------------
      if (half == 1) {
#pragma unroll
        for (int c = 0; c < kCMax; ++c)
          if (c < C) xchg[row_in_tile * kCMax + c] = logit[c];
      }
      epi_bar();
      if (half == 0) {
#pragma unroll
        for (int c = 0; c < kCMax; ++c)
          if (c < C) logit[c] += (p.BN >= 64 ? xchg[row_in_tile * kCMax + c] : 0.f) + bo_s[c];
        if (p.head < 0) {
          if (row_ok) {
            float* out = p.logits + ((long long)z * p.B + b) * C;
#pragma unroll
            for (int c = 0; c < kCMax; ++c)
              if (c < C) out[c] = logit[c];
          }
        } else {
          // loss head (same algebra and operation order as head.cu::dlogits_kernel)
          const int y = row_ok ? p.labels[b] : 0;
          float g[kCMax];
          softmax_r<kCMax>(logit, C);
          if (p.head == RBNN_HEAD_LOGITS_CE) {
#pragma unroll
            for (int c = 0; c < kCMax; ++c) logit[c] = logit[c] - (c == y ? 1.f : 0.f);
          } else {
            if (p.head == RBNN_HEAD_MEAN_OF_GRADS) {
#pragma unroll
              for (int c = 0; c < kCMax; ++c) g[c] = logit[c];
            } else {
#pragma unroll
              for (int c = 0; c < kCMax; ++c) g[c] = (c < C && row_ok) ? __ldg(p.pbar + (long long)b * C + c) : 0.f;
            }
            if (p.head != RBNN_HEAD_UPSTREAM) {
              softmax_r<kCMax>(g, C);
#pragma unroll
              for (int c = 0; c < kCMax; ++c) g[c] -= (c == y ? 1.f : 0.f);
            }
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < kCMax; ++c)
              if (c < C) dot = fmaf(logit[c], g[c], dot);
#pragma unroll
            for (int c = 0; c < kCMax; ++c) logit[c] = logit[c] * (g[c] - dot);
          }
#pragma unroll
          for (int c = 0; c < kCMax; ++c)
            if (c < C) xchg[row_in_tile * kCMax + c] = logit[c];
        }
      }
      epi_bar();
      // sentinel-fill the unused worklist slots of this item
      if (wl) {
        const uint32_t used = min(*wl_count, (uint32_t)kWorkPerItem);
        for (uint32_t i = used + et; i < (uint32_t)kWorkPerItem; i += kEpiWarps * 32) wl[i] = kSentinel;
      }
      if (p.head < 0) continue;
      if (half == 1) {
#pragma unroll
        for (int c = 0; c < kCMax; ++c)
          if (c < C) logit[c] = xchg[row_in_tile * kCMax + c];
      }

      // ---------------- pass 2: dH = (dlogits . Wo) * leaky'(H) for this thread's columns ----------------
      if (active && row_ok) {
        const long long orow = ((long long)z * p.B + b) * H;
        const int nchunks = cols_half / 32 > 0 ? cols_half / 32 : 1;
        for (int n = 0; n < p.n_tiles; ++n) {
#pragma unroll 1
          for (int cc = 0; cc < nchunks; ++cc) {
            const int c0 = half * cols_half + cc * 32;
            const int jbase = n * p.BN + c0;
            const int word = n * nchunks + cc;
            uint32_t bits = 0u;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i == word) bits = mbits[i];
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              if (c0 + q < p.BN) {
                const int j = jbase + q;
                float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < kCMax; ++c)
                  if (c < C) {
                    const float4 w = *reinterpret_cast<const float4*>(wo_s + c * H + j);
                    d[0] = fmaf(logit[c], w.x, d[0]);
                    d[1] = fmaf(logit[c], w.y, d[1]);
                    d[2] = fmaf(logit[c], w.z, d[2]);
                    d[3] = fmaf(logit[c], w.w, d[3]);
                  }
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (!((bits >> (q + e)) & 1u)) d[e] *= kSlopeF;
                if (BF16) {
                  const __nv_bfloat162 a = __floats2bfloat162_rn(d[0], d[1]), bb2 = __floats2bfloat162_rn(d[2], d[3]);
                  uint2 pk;
                  pk.x = *reinterpret_cast<const uint32_t*>(&a);
                  pk.y = *reinterpret_cast<const uint32_t*>(&bb2);
                  __stcs(reinterpret_cast<uint2*>(p.dh_bf + orow + j), pk);          // streaming: read once by the next kernel
                } else {
                  float4 hi4, lo4;
                  hi4.x = to_tf32_rn(d[0]); hi4.y = to_tf32_rn(d[1]); hi4.z = to_tf32_rn(d[2]); hi4.w = to_tf32_rn(d[3]);
                  lo4.x = d[0] - hi4.x; lo4.y = d[1] - hi4.y; lo4.z = d[2] - hi4.z; lo4.w = d[3] - hi4.w;
                  __stcs(reinterpret_cast<float4*>(p.dh_hi + orow + j), hi4);        // streaming: do not displace X / W1 in L2
                  __stcs(reinterpret_cast<float4*>(p.dh_lo + orow + j), lo4);
                }
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// One warp per worklist slot batch: exact re-evaluation of the queued pre-activations; where the sign the
// epilogue assumed was wrong, dH[z, b, j] is rescaled by slope^(+-1) in place.
__global__ void __launch_bounds__(256)
fixup_kernel(const unsigned long long* __restrict__ wl, long long nslots, const float* __restrict__ x,
             const float* __restrict__ bank, long long P, long long b1_off, int z_row0, int B, int D, int H,
             float* __restrict__ dh_hi, float* __restrict__ dh_lo) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long base = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < nslots;
       base += nwarps * 32) {
    const unsigned long long mine = base + lane < nslots ? wl[base + lane] : kSentinel;
    unsigned valid = __ballot_sync(0xffffffffu, mine != kSentinel);
    while (valid) {
      const int src = __ffs(valid) - 1;
      valid &= valid - 1;
      const unsigned long long e = __shfl_sync(0xffffffffu, mine, src);
      const int z = (int)(e >> 44), b = (int)((e >> 20) & 0xFFFFFF), j = (int)((e >> 4) & 0xFFFF);
      const bool assumed_pos = (e & 1ull) != 0;
      const float* __restrict__ xr = x + (long long)b * D;
      const float* __restrict__ wrow = bank + (long long)(z_row0 + z) * P;
      const float* __restrict__ w = wrow + (long long)j * D;            // W1 is the first tensor of a bank row
      double s = 0.0;
      for (int d = lane; d < D; d += 32) s = fma((double)__ldg(xr + d), (double)__ldg(w + d), s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const bool pos = (float)(s + (double)__ldg(wrow + b1_off + j)) > 0.f;
      if (pos != assumed_pos && lane == 0) {
        const long long o = ((long long)z * B + b) * H + j;
        float v = dh_hi[o] + dh_lo[o];
        v = pos ? v * 100.f : v * kSlopeF;          // undo / apply the LeakyReLU slope
        const float hi = to_tf32_rn(v);
        dh_hi[o] = hi;
        dh_lo[o] = v - hi;
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
  const int words = n_tiles * ((bn >= 64 ? bn / 2 : bn) / 32 > 0 ? (bn >= 64 ? bn / 2 : bn) / 32 : 1);
  if (words > 16) return false;
  if (bn >= 64 && ((bn / 2) % 32)) return false;
  return ((C * H + C + 3) & ~3) + H <= kParamFloatsMax;
}

size_t fused_worklist_slots(int B, int Z) { return (size_t)Z * ((B + kBM - 1) / kBM) * kWorkPerItem; }

int fused_forward_head(const FusedDesc& d, cudaStream_t st, std::string* err) {
  std::string local;
  if (!err) err = &local;
  const bool bf16 = d.mode == MODE_BF16;
  if (d.B <= 0 || d.Z <= 0) return 0;
  if (!fused_supported(d.H, d.C)) { *err = "fused_forward_head: unsupported hidden / class size"; return 1; }
  const int kbb = d.kblock_bytes == 128 ? 128 : 64;
  auto kern = bf16 ? (kbb == 128 ? fc_fused_kernel<true, 128> : fc_fused_kernel<true, 64>)
                   : (kbb == 128 ? fc_fused_kernel<false, 128> : fc_fused_kernel<false, 64>);
  static bool attr_done[2][2] = {{false, false}, {false, false}};
  if (!attr_done[bf16][kbb == 128]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem);
    if (e != cudaSuccess) { *err = std::string("fused_forward_head: cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    attr_done[bf16][kbb == 128] = true;
  }
  FParams p{};
  p.B = d.B; p.D = d.D; p.H = d.H; p.C = d.C; p.Z = d.Z;
  p.BN = d.H <= 256 ? d.H : 256;
  p.n_tiles = d.H <= 256 ? 1 : d.H / 256;
  p.m_tiles = (d.B + kBM - 1) / kBM;
  p.num_items = p.m_tiles * d.Z;
  const int kbe = kbb / (bf16 ? 2 : 4);
  p.num_kb = (d.D + kbe - 1) / kbe;
  p.head = d.head;
  p.bank = d.bank; p.P = d.P; p.b1_off = d.b1_off; p.wo_off = d.wo_off; p.bo_off = d.bo_off; p.z_row0 = d.z_row0;
  p.labels = d.labels; p.pbar = d.pbar;
  p.x = d.x; p.xnorm = d.xnorm; p.wnorm = d.wnorm;
  p.eps = (bf16 || !d.worklist || d.head < 0) ? 0.f : d.eps;
  p.dh_hi = d.dh_hi; p.dh_lo = d.dh_lo; p.dh_bf = reinterpret_cast<__nv_bfloat16*>(d.dh_bf); p.logits = d.logits;
  p.worklist = p.eps > 0.f ? d.worklist : nullptr;
  if (d.head >= 0 && (bf16 ? !d.dh_bf : (!d.dh_hi || !d.dh_lo))) { *err = "fused_forward_head: missing dH output"; return 1; }
  if (d.head < 0 && !d.logits) { *err = "fused_forward_head: missing logits output"; return 1; }
  if (d.head >= 0 && d.head != RBNN_HEAD_MEAN_OF_GRADS && d.head != RBNN_HEAD_LOGITS_CE && !d.pbar) {
    *err = "fused_forward_head: this head needs pbar";
    return 1;
  }
  CUtensorMap mAh, mAl, mBh, mBl;
  if (make_map(&mAh, d.X.hi, bf16, d.D, d.B, 1, d.X.ld, 0, kBM, kbb, err)) return 1;
  if (make_map(&mBh, d.W1.hi, bf16, d.D, d.H, d.Z, d.W1.ld, d.W1.zstride, p.BN, kbb, err)) return 1;
  if (!bf16) {
    if (make_map(&mAl, d.X.lo, false, d.D, d.B, 1, d.X.ld, 0, kBM, kbb, err)) return 1;
    if (make_map(&mBl, d.W1.lo, false, d.D, d.H, d.Z, d.W1.ld, d.W1.zstride, p.BN, kbb, err)) return 1;
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
  const long long nslots = (long long)fused_worklist_slots(d.B, d.Z);
  const long long warps = (nslots + 31) / 32;
  const unsigned blocks = (unsigned)std::min<long long>((warps + 7) / 8, (long long)d.sm_count * 8);
  fixup_kernel<<<blocks, 256, 0, st>>>(d.worklist, nslots, d.x, d.bank, d.P, d.b1_off, d.z_row0, d.B, d.D, d.H,
                                       d.dh_hi, d.dh_lo);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("fused_fixup launch: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

}  // namespace tc
}  // namespace rbnn
