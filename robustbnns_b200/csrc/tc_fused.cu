// Fused forward + loss head for arch fc on tcgen05 (see FusedDesc in tc_gemm.cuh).
//
// Replaces, for every posterior sample z and every input b at once, the reference's
//   output = net.forward(x, n_samples=1, seeds=[z]); loss = CE(output, y); loss.backward()
// up to the gradient w.r.t. the hidden layer (lossGradients.py:33-36 / adversarialAttacks.py:74-78;
// model_nn.py:77-82 for the network).  The hidden activations never leave the SM:
//
//   work item = (sample z, tile of 128 inputs); per item the hidden dimension is cut into n-tiles of
//   BN <= 256 columns that alternate between the two TMEM accumulator stages.
//   warp 0       TMA producer (X tile + W1_z tile K-blocks, SWIZZLE_128B ring; X is loaded with an L2 evict_last
//                policy: it is re-read for every sample while dH streams through the cache)
//   warp 1       tcgen05.mma issuer (kind::tf32 x3 passes, kind::f16 x3 passes on scaled fp16 hi/lo, or kind::f16 on bf16)
//   warps 2..17  epilogue, warp = (TMEM lane quadrant, column quarter): a thread owns ONE input row (tcgen05.ld 32x32b)
//                and 32-column chunks of it.
//                Pass 1 (per n-tile, overlapping the next n-tile's MMAs): +b1 -> guard-band test -> LeakyReLU -> mask
//                bits, partial logits += act . Wo^T with packed fp32 FMAs (fma.rn.f32x2: two hidden units per
//                instruction, Wo_z as fp32 in shared memory, broadcast 16-byte loads).
//                Head: the four column quarters of a row add their partial logits through shared memory (fixed order),
//                one thread per row evaluates the loss head (softmax -> g -> dlogits).
//                Pass 2: dH = (dlogits . Wo) * leaky'(H), row-cooperative: a warp walks the 32 input rows of its TMEM lane
//                quadrant and, for one row, lane l owns 4 consecutive hidden columns (n-tile l / 16, columns
//                64 cq + 4 (l % 16) ..) whose Wo entries it keeps in registers for the whole item; the row's dlogits are a
//                broadcast shared-memory load, its mask words come from the owner lane by shuffle.  16 lanes write one
//                full 128-byte line of the pre-split output (fp16 hi/lo, tf32 hi/lo or bf16) per store instruction.
//                (With thread = row in pass 2 a store instruction touched 32 different lines: 64 L1 wavefronts per KB
//                written, and the dH stores alone cost 2.7 ms of a 10.2 ms chunk.)
//   (Round 1 ran both small GEMMs of the head as mma.sync on fp16 hi/lo fragments: on sm_100 a legacy HMMA issues every
//   ~32 cycles per sub-core, 24.5 k cycles per item for them alone against 38 k cycles of tcgen05 MMAs; the packed
//   fp32 FMAs need ~10 k and no fragment shuffling, staging conversions or operand scaling.)
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
constexpr int kEpiWarps = 16;                     // 4 TMEM lane quadrants x 4 column quarters
constexpr int kEpiThreads = kEpiWarps * 32;       // 512
constexpr int kThreadsF = 64 + kEpiThreads;       // 576
constexpr int kCMax = 10;                         // classes the fused epilogue covers (registers per row)
constexpr int kHMax = 512;                        // hidden units the staged parameters cover
constexpr int kMaskWords = 4;                     // 32-column chunks per thread: n_tiles * ceil(BN / 128) <= 4
constexpr int kMaskWordsItem = kHMax / 32;        // mask words per input row of an item (keep mode): word = j / 32
constexpr int kXStride = kCMax + 1;               // floats per row of the exchange buffers (odd: conflict-free)
constexpr int kDStride = 12;                      // floats per row of the dlogits buffer pass 2 reads (16-byte aligned rows)
// shared memory after the ring and the barriers: Wo [C][H] fp32 | b1 [kHMax] | bo [16] | xa [128][11] | xb [128][12]
constexpr int kFusedSmem = kRingF + 1024 + 256 + (kCMax * kHMax + kHMax + 16 + kBM * kXStride + kBM * kDStride) * 4 + 16;
static_assert(kFusedSmem <= 232448, "fused kernel exceeds the 227 KB shared memory of an sm_100 CTA");
constexpr unsigned long long kSentinel = ~0ull;

struct FParams {
  int B, D, H, C, Z, BN, n_tiles, m_tiles, m_units, num_items, num_kb;   // m_units = m_tiles, or ceil(m_tiles / 2) for CTA pairs
  int head;
  const float* bank; long long P, b1_off, wo_off, bo_off; int z_row0;
  const int32_t* labels; const float* pbar;
  const float* x; const float* xnorm; const float* wnorm; float eps;
  void* dh_hi; void* dh_lo; __nv_bfloat16* dh_bf; float* logits;
  const float* unscale; const float* dh_scale;      // F16X3 device scalars (see FusedDesc)
  const unsigned* xlo_zero;                          // F16X3 device flag: the lo part of the inputs is zero -> two passes
  uint32_t* maskbuf;                                 // head == -2 (keep mode): [item][kMaskWordsItem][128] LeakyReLU mask words
  int debug;                                         // timing experiments (RBNN_FUSED_DEBUG): 1 no dH stores, 2 no pass 2, 4 no pass-1 math, 8 phase timers, 16 no L2 hint
  unsigned long long* worklist;
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

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

// (x, y) -> packed fp16 pairs hi = rn(x, y), lo = rn((x, y) - hi); the lower half holds x
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - f.x, y - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// packed fp32 pairs (sm_100: fma.rn.f32x2 = two FMAs per issue slot)
__device__ __forceinline__ uint64_t pack2(float x, float y) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// 32 consecutive bytes, one full sector per lane; streaming (evict-first) so dH does not displace X / W1 in L2
__device__ __forceinline__ void st_cs_u8(void* ptr, const uint32_t* a) {
  asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a[0]), "r"(a[1]), "r"(a[2]),
               "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
               : "memory");
}
// 32 TMEM lanes x 32 consecutive fp32 columns, no wait inside (pair with tmem_ld_wait)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 TMEM lanes x 16 consecutive fp32 columns, no wait inside
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// TMA load with an L2 cache policy (createpolicy)
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                 int c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
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

// Loss head on one input row held by one thread: l = logits (bias included) -> l = dL/dlogits.  Same algebra and
// operation order as head.cu::dlogits_kernel.  Slots c >= C come out as 0.
__device__ __forceinline__ void head_row(int head, float (&l)[kCMax], int C, int y, const float* __restrict__ pbar_row) {
  float g[kCMax];
  if (head == RBNN_HEAD_LOGITS_UPSTREAM) {     // loss of the mean LOGITS (ensembles, deterministic nets): dlogits = d_pbar
#pragma unroll
    for (int c = 0; c < kCMax; ++c) l[c] = (c < C && pbar_row) ? __ldg(pbar_row + c) : 0.f;
    return;
  }
  softmax_r<kCMax>(l, C);
  if (head == RBNN_HEAD_LOGITS_CE) {
#pragma unroll
    for (int c = 0; c < kCMax; ++c) l[c] = c < C ? l[c] - (c == y ? 1.f : 0.f) : 0.f;
    return;
  }
  if (head == RBNN_HEAD_MEAN_OF_GRADS) {
#pragma unroll
    for (int c = 0; c < kCMax; ++c) g[c] = l[c];
  } else {
#pragma unroll
    for (int c = 0; c < kCMax; ++c) g[c] = (c < C && pbar_row) ? __ldg(pbar_row + c) : 0.f;
  }
  if (head != RBNN_HEAD_UPSTREAM) {
    softmax_r<kCMax>(g, C);
#pragma unroll
    for (int c = 0; c < kCMax; ++c) g[c] -= (c == y ? 1.f : 0.f);
  }
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < kCMax; ++c)
    if (c < C) dot = fmaf(l[c], g[c], dot);
#pragma unroll
  for (int c = 0; c < kCMax; ++c) l[c] = c < C ? l[c] * (g[c] - dot) : 0.f;
}

// ---- pass 2, row-cooperative -------------------------------------------------------------------------------------
// For one input row, lane l of epilogue warp (quad, cq) owns the 4 hidden columns j .. j+3 with
//   n-tile n = l / 16, column inside the n-tile c = 64 cq + 4 (l % 16)      (the two 32-column chunks 2 cq, 2 cq + 1
// of each n-tile that the warp also owns in pass 1), so 16 lanes cover one 128-byte line of fp16 output.
struct P2Lane {
  int j;          // first hidden unit of this lane
  int word;       // which of the owner thread's mask words holds its bits: n * 2 + chunk parity
  int shift;      // bit of column j inside that word
  bool valid;     // inside the hidden layer
};
__device__ __forceinline__ P2Lane p2_lane(int lane, int cq, int BN, int n_tiles) {
  P2Lane g;
  const int n = lane >> 4, c = cq * 64 + 4 * (lane & 15);
  g.valid = n < n_tiles && c < BN;
  g.j = n * BN + c;
  g.word = n * 2 + ((lane & 15) >> 3);
  g.shift = 4 * (lane & 7);
  return g;
}
// The lane's slice of Wo_z as column pairs, exactly as they lie in shared memory (so they stay in aligned register pairs):
// wq[c][0] = (Wo[c][j], Wo[c][j+1]), wq[c][1] = (Wo[c][j+2], Wo[c][j+3]); classes >= C are zero.
__device__ __forceinline__ void p2_load_wo(uint64_t (&wq)[kCMax][2], const float* __restrict__ wo_s, int H, int C,
                                           const P2Lane& g) {
#pragma unroll
  for (int c = 0; c < kCMax; ++c) {
    ulonglong2 t = make_ulonglong2(0ull, 0ull);
    if (g.valid && c < C) t = *reinterpret_cast<const ulonglong2*>(wo_s + c * H + g.j);
    wq[c][0] = t.x;
    wq[c][1] = t.y;
  }
}
__device__ __forceinline__ void st_cs_u2(void* ptr, uint32_t a, uint32_t b) {
  asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(ptr), "r"(a), "r"(b) : "memory");
}
// dH[row][j .. j+3] = (sum_c dl[row][c] Wo[c][j ..]) * leaky'(H) for the 32 rows of TMEM lane quadrant `quad`, written as
// fp16 hi/lo (F16X3), tf32 hi/lo (TF32X3) or bf16.  xd: [128][kDStride] dlogits (already carrying the F16X3 range scale,
// slots >= C zero); mbits: this THREAD's mask words (thread = row `lane` of the quadrant), fetched by shuffle.
// b0: input index of the quadrant's first row; zrow: z * B.  Two columns per FFMA2 (fma.rn.f32x2 with the row's dlogit
// broadcast), 10-deep chains, two rows in flight.
template <int MODE>
__device__ __forceinline__ void pass2_rows(const uint64_t (&wq)[kCMax][2], const P2Lane& g, const float* xd,
                                           const uint32_t (&mbits)[kMaskWords], int quad, int b0, int B, int H,
                                           long long zrow, void* dh_hi, void* dh_lo, __nv_bfloat16* dh_bf, bool no_store,
                                           int dbg = 0) {
  // byte pointers of (first row of the quadrant, column j) in the output arrays, advanced one row per iteration
  constexpr int ES = MODE == MODE_TF32X3 ? 4 : 2;
  long long o0 = ((zrow + b0) * H + g.j) * ES;
  if (dbg & 64) o0 &= (long long)((32 << 20) - 1);          // timing experiment: every store lands in the first 32 MB (L2-resident)
  char* ph = reinterpret_cast<char*>(MODE == MODE_BF16 ? (void*)dh_bf : dh_hi) + o0;
  char* pl = MODE == MODE_BF16 ? nullptr : reinterpret_cast<char*>(dh_lo) + o0;
  const int row_bytes = H * ES;
  const int rows_ok = (g.valid && !no_store) ? min(32, B - b0) : 0;
#pragma unroll 2
  for (int rr = 0; rr < 32; ++rr, ph += row_bytes, pl += row_bytes) {
    uint32_t w = 0u;
#pragma unroll
    for (int k = 0; k < kMaskWords; ++k) {
      const uint32_t t = __shfl_sync(0xffffffffu, mbits[k], rr);
      if (k == g.word) w = t;
    }
    const uint32_t bits = w >> g.shift;
    const float* d = xd + (quad * 32 + rr) * kDStride;
    const float4 d0 = *reinterpret_cast<const float4*>(d);
    const float4 d1 = *reinterpret_cast<const float4*>(d + 4);
    const float2 d2 = *reinterpret_cast<const float2*>(d + 8);
    const float dl[kCMax] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y};
    uint64_t a0 = 0ull, a1 = 0ull;
#pragma unroll
    for (int c = 0; c < kCMax; ++c) {
      const uint64_t dd = pack2(dl[c], dl[c]);
      a0 = ffma2(dd, wq[c][0], a0);
      a1 = ffma2(dd, wq[c][1], a1);
    }
    float v[4];
    unpack2(a0, v[0], v[1]);
    unpack2(a1, v[2], v[3]);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (!((bits >> q) & 1u)) v[q] *= kSlopeF;
    if (rr < rows_ok) {
      if (MODE == MODE_TF32X3) {
        float4 hi4, lo4;
        hi4.x = to_tf32_rn(v[0]); hi4.y = to_tf32_rn(v[1]); hi4.z = to_tf32_rn(v[2]); hi4.w = to_tf32_rn(v[3]);
        lo4.x = v[0] - hi4.x; lo4.y = v[1] - hi4.y; lo4.z = v[2] - hi4.z; lo4.w = v[3] - hi4.w;
        __stcs(reinterpret_cast<float4*>(ph), hi4);   // streaming: keep X / W1 in L2
        __stcs(reinterpret_cast<float4*>(pl), lo4);
      } else if (MODE == MODE_BF16) {
        const __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
        st_cs_u2(ph, *reinterpret_cast<const uint32_t*>(&t0), *reinterpret_cast<const uint32_t*>(&t1));
      } else {
        uint32_t h0, l0, h1, l1;
        split_pair(v[0], v[1], h0, l0);
        split_pair(v[2], v[3], h1, l1);
        st_cs_u2(ph, h0, h1);
        if (!(dbg & 32)) st_cs_u2(pl, l0, l1);                // timing experiment: hi stores only
      }
    }
  }
}

// Wo_z [C][H], b1_z [H], bo_z [C] of one bank row -> shared memory (fp32 as is).  Bank rows are 8-byte aligned
// (P even) and every offset is even for even H: 8-byte loads.
__device__ __forceinline__ void stage_head_params(const float* __restrict__ wrow, long long wo_off, long long b1_off,
                                                  long long bo_off, int C, int H, float* wo_s, float* b1_s, float* bo_s,
                                                  int t, int nthreads) {
  const bool al8 = ((reinterpret_cast<uintptr_t>(wrow + wo_off) & 7u) == 0) && !((C * H) & 1);
  if (al8) {
    const float2* __restrict__ src = reinterpret_cast<const float2*>(wrow + wo_off);
    float2* dst = reinterpret_cast<float2*>(wo_s);
    for (int i = t; i < C * H / 2; i += nthreads) dst[i] = __ldg(src + i);
  } else {
    for (int i = t; i < C * H; i += nthreads) wo_s[i] = __ldg(wrow + wo_off + i);
  }
  if (b1_s)
    for (int i = t; i < H; i += nthreads) b1_s[i] = __ldg(wrow + b1_off + i);
  if (bo_s && t < C) bo_s[t] = __ldg(wrow + bo_off + t);
}

// CT: class count known at compile time (0 = use p.C): the headline net has 10 classes, and with runtime bounds every
// FFMA2 of pass 1 was predicated and shadowed by two MOVs.
// PAIR: two CTAs of a cluster (the SM pair of a TPC) work on one sample and two adjacent 128-input tiles with
// tcgen05.mma.cta_group::2 (M = 256): each CTA stages its own inputs and HALF of the W1 n-tile, the leader issues the
// MMAs for both, each CTA's TMEM receives its own 128 accumulator rows and each CTA runs its own epilogue.  Per SM this
// removes a third of the operand bytes fetched from L2 and a third of the tensor cores' shared-memory operand reads --
// the shared-memory data pipe (MMA operand reads + the epilogue's loads) is what bounds the single-CTA kernel.  The
// barrier protocol is tc_gemm.cu's relay variant.
template <int MODE, int KBB, int CT, bool PAIR>
__global__ void __launch_bounds__(kThreadsF, 1)
fc_fused_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                const FParams p) {
  constexpr bool BF16 = MODE == MODE_BF16;
  constexpr bool F16K = MODE != MODE_TF32X3;            // kind::f16 MMAs on 2-byte operands
  constexpr int NARR = BF16 ? 1 : 2;
  constexpr int kATileF = kBM * KBB, kBTileF = (PAIR ? kBNMax / 2 : kBNMax) * KBB;
  constexpr int NCTA = PAIR ? 2 : 1;
  constexpr int STAGE = NARR * (kATileF + kBTileF);
  constexpr int NSTAGE = kRingF / STAGE;
  // two-pass forward (F16X3, inputs on the pixel grid: X.lo is zero and never loaded): a stage shrinks by the X.lo tile
  // and the ring holds more of them (CTA pairs: 4 x 48 KB instead of 3 x 64 KB)
  constexpr int STAGE2 = MODE == MODE_F16X3 ? STAGE - kATileF : STAGE;
  constexpr int NBAR = kRingF / STAGE2;                   // barriers are laid out for the larger stage count
  constexpr int KBE = KBB / (F16K ? 2 : 4);
  constexpr int KSTEPS = KBB / 32;
  constexpr uint32_t FMT = MODE == MODE_BF16 ? 1u : (MODE == MODE_F16X3 ? 0u : 2u);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  const uint32_t bars = ring + kRingF;
  const uint32_t full0 = bars, empty0 = bars + 8 * NBAR;
  const uint32_t tfull0 = bars + 16 * NBAR, tempty0 = tfull0 + 16;
  uint8_t* gen = smem_raw + (bars - raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 16 * NBAR + 32);
  uint32_t* wl_count = reinterpret_cast<uint32_t*>(gen + 16 * NBAR + 40);
  const uint32_t pfull0 = bars + 16 * NBAR + 48;         // pair protocol: "the peer's stage is full" (leader's copy)
  static_assert(16 * NBAR + 48 + 8 * NBAR <= 256, "barrier block");
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs)
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // scheduling unit: CTA or CTA pair
  const int num_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  float* wo_s = reinterpret_cast<float*>(gen + 256);       // [C][H] fp32
  float* b1_s = wo_s + kCMax * kHMax;                       // [H]
  float* bo_s = b1_s + kHMax;                               // [16]
  float* xa = bo_s + 16;                                    // [128][kXStride] partial logits / final dlogits
  float* xb = xa + kBM * kXStride;                          // [128][kXStride]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NBAR; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
      if (PAIR) mbar_init(pfull0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, kEpiWarps * NCTA);           // one arrival per epilogue warp (of both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
    if (!BF16) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
    }
  }
  static_assert(kTmemCols == 512, "tmem_alloc allocates 512 columns");
  if (warp == 1) tmem_alloc<PAIR>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                            // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int bn_cta = PAIR ? p.BN / 2 : p.BN;               // W1 rows (hidden units) staged by this CTA per n-tile
  const uint32_t stage_tx = (uint32_t)NARR * (uint32_t)(kATileF + bn_cta * KBB);
  const bool two_pass = MODE == MODE_F16X3 && p.xlo_zero && __ldg(p.xlo_zero) != 0u;   // same word for every role / CTA
  const int nst = two_pass ? NBAR : NSTAGE;
  const uint32_t stage_bytes = two_pass ? (uint32_t)STAGE2 : (uint32_t)STAGE;
  const uint32_t b_off = (two_pass ? 1u : (uint32_t)NARR) * (uint32_t)kATileF;            // B tiles follow the A tile(s)

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint64_t pol_x;      // the inputs are re-read for every posterior sample: keep them in L2 while dH streams through
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_x));
      const bool hint = !(p.debug & 16);
      const bool a_lo = !BF16 && !two_pass;              // X.lo needed
      const uint32_t tx = a_lo || BF16 ? stage_tx : stage_tx - (uint32_t)kATileF;
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0;
      const long long t_begin = clock64();
      for (int item = unit; item < p.num_items; item += num_units) {
        const int z = item / p.m_units, m_idx = PAIR ? 2 * (item % p.m_units) + (int)rank : item % p.m_units;
        for (int n = 0; n < p.n_tiles; ++n) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            const long long t0 = clock64();
            mbar_wait(empty0 + 8 * stage, phase ^ 1u);
            w_empty += clock64() - t0;
            const uint32_t fb = full0 + 8 * stage;
            mbar_expect_tx(fb, tx);
            const uint32_t sa = ring + stage * stage_bytes;
            const uint32_t sb = sa + b_off;
            if (hint) tma_load_3d_hint(sa, &tmAh, fb, kb * KBE, m_idx * kBM, 0, pol_x);
            else tma_load_3d(sa, &tmAh, fb, kb * KBE, m_idx * kBM, 0);
            tma_load_3d(sb, &tmBh, fb, kb * KBE, n * p.BN + (int)rank * bn_cta, z);
            if (!BF16) {
              if (a_lo) {
                if (hint) tma_load_3d_hint(sa + kATileF, &tmAl, fb, kb * KBE, m_idx * kBM, 0, pol_x);
                else tma_load_3d(sa + kATileF, &tmAl, fb, kb * KBE, m_idx * kBM, 0);
              }
              tma_load_3d(sb + kBTileF, &tmBl, fb, kb * KBE, n * p.BN + (int)rank * bn_cta, z);
            }
            if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if ((p.debug & 8) && blockIdx.x == 0)
        printf("[fused cta0] TMA producer: total %lld cyc, waiting for an empty stage %lld\n", clock64() - t_begin, w_empty);
    }
  } else if (warp == 1) {
    if (PAIR && lane == 0 && rank == 1) {
      // ===================== peer relay: own stage full -> tell the leader =====================
      const uint32_t pf = mapa_u32(pfull0, 0u);
      uint32_t stage = 0, phase = 0;
      for (int item = unit; item < p.num_items; item += num_units) {
        const int total_kb = p.n_tiles * p.num_kb;
        for (int i = 0; i < total_kb; ++i) {
          mbar_wait(full0 + 8 * stage, phase);
          mbar_arrive_cluster(pf + 8 * stage);
          if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1u; }
        }
      }
    }
    if (lane == 0 && rank == 0) {
      // ===================== MMA issuer (leader CTA of a pair) =====================
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)((kBM * NCTA) >> 4) << 24);
      const bool a_lo_used = !two_pass;
      uint32_t stage = 0, phase = 0, it = 0;
      long long w_tempty = 0, w_full = 0;
      const long long t_begin = clock64();
      for (int item = unit; item < p.num_items; item += num_units) {
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
            if (PAIR) mbar_wait(pfull0 + 8 * stage, phase);
            w_full += clock64() - t0;
            tc_fence_after();
            const uint32_t sa = ring + stage * stage_bytes;
            const uint32_t sb = sa + b_off;
            const uint64_t a_hi = smem_desc<KBB>(sa), b_hi = smem_desc<KBB>(sb);
            if (BF16) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                if (PAIR) tc_mma_pair<true>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, accumulate);
                else tc_mma<true>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, accumulate);
                accumulate = 1;
              }
            } else {
              const uint64_t a_lo = smem_desc<KBB>(sa + kATileF), b_lo = smem_desc<KBB>(sb + kBTileF);
              if (a_lo_used) {
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  if (PAIR) {
                    tc_mma_pair<F16K>(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, accumulate);
                    tc_mma_pair<F16K>(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                    tc_mma_pair<F16K>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
                  } else {
                    tc_mma<F16K>(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, accumulate);
                    tc_mma<F16K>(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                    tc_mma<F16K>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
                  }
                  accumulate = 1;
                }
              } else {          // inputs on the pixel grid: X_lo == 0, the X_lo . W1_hi pass drops out
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  if (PAIR) {
                    tc_mma_pair<F16K>(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, accumulate);
                    tc_mma_pair<F16K>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
                  } else {
                    tc_mma<F16K>(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, accumulate);
                    tc_mma<F16K>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
                  }
                  accumulate = 1;
                }
              }
            }
            if (PAIR) tc_commit_pair(empty0 + 8 * stage, (uint16_t)3);   // stage reusable in both CTAs once these MMAs retire
            else tc_commit(empty0 + 8 * stage);
            if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1u; }
          }
          if (PAIR) tc_commit_pair(tfull0 + 8 * as, (uint16_t)3);        // accumulator complete (both CTAs' epilogues)
          else tc_commit(tfull0 + 8 * as);
        }
      }
      if ((p.debug & 8) && blockIdx.x == 0)
        printf("[fused cta0] MMA issuer: total %lld cyc, waiting for a free accumulator %lld, for operands %lld\n",
               clock64() - t_begin, w_tempty, w_full);
    }
  } else {
    // ===================== epilogue: warps 2..17 =====================
    // warp = (TMEM lane quadrant `quad` = warp % 4, column quarter `cq`); in pass 1 thread = input row quad*32 + lane.
    // Of an n-tile's BN / 32 chunks of 32 columns this warp takes chunks 2 cq and 2 cq + 1 (64 contiguous columns).
    const int ew = warp - 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int cq = ew >> 2;                    // column quarter
    const int et = threadIdx.x - 64;           // 0..511
    const int C = CT ? CT : p.C, H = p.H;
    const int nchunks = p.BN / 32;             // chunks per n-tile (1..8)
    const int row = quad * 32 + lane;          // row inside the 128-input tile
    const float unscale = (MODE == MODE_F16X3) ? __ldg(p.unscale) : 1.f;     // 1 / (s_X s_W1)
    const float dh_scale = (MODE == MODE_F16X3) ? __ldg(p.dh_scale) : 1.f;   // dH -> fp16 range
    const P2Lane g2 = p2_lane(lane, cq, p.BN, p.n_tiles);
    uint32_t it = 0;
#ifdef RBNN_FUSED_TIMERS
    long long w_tfull = 0, c_stage = 0, c_p1 = 0, c_head = 0, c_p2 = 0, t_mark = clock64();
    const long long t_begin = t_mark;
#define RBNN_TMARK(acc) { const long long t = clock64(); acc += t - t_mark; t_mark = t; }
#else
#define RBNN_TMARK(acc)
#endif
    const uint32_t tempty_leader = PAIR ? mapa_u32(tempty0, 0u) : tempty0;
    for (int item = unit; item < p.num_items; item += num_units) {
      const int z = item / p.m_units, m_idx = PAIR ? 2 * (item % p.m_units) + (int)rank : item % p.m_units;
      const bool tile_ok = m_idx < p.m_tiles;             // a pair's second tile may lie entirely past the last input
      const long long item_w = (long long)z * p.m_tiles + m_idx;   // this tile's slot in the worklist / mask buffer
      const float* __restrict__ wrow = p.bank + (long long)(p.z_row0 + z) * p.P;
      RBNN_TMARK(c_p2)
      // ---------------- stage Wo_z, b1_z, bo_z (fp32) ----------------
      epi_bar();                                // everyone is done with the previous item's parameters
      stage_head_params(wrow, p.wo_off, p.b1_off, p.bo_off, C, H, wo_s, b1_s, bo_s, et, kEpiThreads);
      if (et == 0) *wl_count = 0u;
      epi_bar();
      const int b = m_idx * kBM + row;
      const bool rok = b < p.B;
      const float guard = (rok && p.eps > 0.f) ? p.eps * __ldg(p.wnorm + p.z_row0 + z) * __ldg(p.xnorm + b) : 0.f;
      unsigned long long* wl = (p.worklist && tile_ok) ? p.worklist + item_w * kWorkPerItem : nullptr;
      uint64_t acc2[kCMax];                     // partial logits: (even hidden units, odd hidden units) of this thread's chunks
#pragma unroll
      for (int c = 0; c < kCMax; ++c) acc2[c] = 0ull;
      uint32_t mbits[kMaskWords];               // word n * 2 + i: chunk 2 cq + i of n-tile n
#pragma unroll
      for (int i = 0; i < kMaskWords; ++i) mbits[i] = 0u;

      RBNN_TMARK(c_stage)
      // ---------------- pass 1 ----------------
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        if (n >= p.n_tiles) break;
        const uint32_t as = it & 1u;
#ifdef RBNN_FUSED_TIMERS
        { const long long t0 = clock64(); mbar_wait(tfull0 + 8 * as, (it >> 1) & 1u); w_tfull += clock64() - t0; }
#else
        mbar_wait(tfull0 + 8 * as, (it >> 1) & 1u);
#endif
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (uint32_t)kBNMax;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ch = 2 * cq + i;
          uint32_t bits = 0u;
          if (ch < nchunks) {
            const int c0 = ch * 32;                                 // column inside the n-tile
            const int j0 = n * p.BN + c0;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {                        // 16 columns at a time (register budget: 96 per thread)
              uint32_t v[16];
              tmem_ld16_nowait(taddr + (uint32_t)(c0 + 16 * hf), v);
              const int jh = j0 + 16 * hf;
              tmem_ld_wait();
              if (p.debug & 4) continue;
              uint32_t hb = 0u;
              float amin = INFINITY;
#pragma unroll
              for (int q = 0; q < 16; q += 4) {
                const float4 bb = *reinterpret_cast<const float4*>(b1_s + jh + q);
                float h[4];
                if (MODE == MODE_F16X3) {
                  h[0] = fmaf(__uint_as_float(v[q]), unscale, bb.x);
                  h[1] = fmaf(__uint_as_float(v[q + 1]), unscale, bb.y);
                  h[2] = fmaf(__uint_as_float(v[q + 2]), unscale, bb.z);
                  h[3] = fmaf(__uint_as_float(v[q + 3]), unscale, bb.w);
                } else {
                  h[0] = __uint_as_float(v[q]) + bb.x;
                  h[1] = __uint_as_float(v[q + 1]) + bb.y;
                  h[2] = __uint_as_float(v[q + 2]) + bb.z;
                  h[3] = __uint_as_float(v[q + 3]) + bb.w;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  amin = fminf(amin, fabsf(h[e]));
                  if (h[e] > 0.f) hb |= 1u << (q + e);
                  v[q + e] = __float_as_uint(h[e]);                 // keep the pre-activation for the guard-band path
                  h[e] = fmaxf(h[e], h[e] * kSlopeF);               // LeakyReLU (slope < 1)
                }
                const uint64_t a01 = pack2(h[0], h[1]), a23 = pack2(h[2], h[3]);
#pragma unroll
                for (int c = 0; c < kCMax; ++c)
                  if (c < C) {
                    const float4 w = *reinterpret_cast<const float4*>(wo_s + c * H + jh + q);
                    acc2[c] = ffma2(a01, pack2(w.x, w.y), acc2[c]);
                    acc2[c] = ffma2(a23, pack2(w.z, w.w), acc2[c]);
                  }
              }
              if (amin < guard) {
                // rare (one 16-column group in ten): some pre-activation lies inside the guard band.  Kept short on
                // purpose -- the warps of an item meet at a barrier after pass 1, so a slow rare path stalls all of them.
                uint32_t inband = 0u;
#pragma unroll
                for (int q = 0; q < 16; ++q)
                  if (fabsf(__uint_as_float(v[q])) < guard) inband |= 1u << q;
                while (inband) {
                  const int q = __ffs(inband) - 1;
                  inband &= inband - 1u;
                  const bool pos = (hb >> q) & 1u;
                  const uint32_t slot = atomicAdd(wl_count, 1u);
                  if (slot < (uint32_t)kWorkPerItem) {
                    wl[slot] = pack_entry(z, b, jh + q, pos);
                  } else {                                          // item budget exhausted: settle it here
                    const bool ex = exact_positive_serial(p.x + (long long)b * p.D, wrow + (long long)(jh + q) * p.D,
                                                          b1_s[jh + q], p.D);
                    if (ex != pos) hb ^= 1u << q;
                  }
                }
              }
              bits |= hb << (16 * hf);
            }
          }
          mbits[n * 2 + i] = bits;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_relaxed_cluster(tempty_leader + 8 * as);
          else mbar_arrive_relaxed(tempty0 + 8 * as);
        }
        ++it;
      }

      RBNN_TMARK(c_p1)
      // ---------------- logits: quarters 1, 3 -> quarters 0, 2; quarter 2 -> quarter 0; loss head on quarter 0 ----------------
      float dl[kCMax];                          // partial logits -> logits -> dlogits
#pragma unroll
      for (int c = 0; c < kCMax; ++c) {
        float a, bsum;
        unpack2(acc2[c], a, bsum);
        dl[c] = a + bsum;
      }
      if (cq & 1) {
        float* dst = (cq == 1 ? xa : xb) + row * kXStride;
#pragma unroll
        for (int c = 0; c < kCMax; ++c) dst[c] = dl[c];
      }
      epi_bar();
      if (!(cq & 1)) {
        const float* src = (cq == 0 ? xa : xb) + row * kXStride;
#pragma unroll
        for (int c = 0; c < kCMax; ++c) dl[c] += src[c];
      }
      epi_bar();
      if (cq == 2) {
#pragma unroll
        for (int c = 0; c < kCMax; ++c) xa[row * kXStride + c] = dl[c];
      }
      epi_bar();
      if (cq == 0) {
#pragma unroll
        for (int c = 0; c < kCMax; ++c) dl[c] += xa[row * kXStride + c] + (c < C ? bo_s[c] : 0.f);
        if (p.head < 0) {
          if (rok) {
            float* out = p.logits + ((long long)z * p.B + b) * C;
#pragma unroll
            for (int c = 0; c < kCMax; ++c)
              if (c < C) out[c] = dl[c];
          }
        } else {
          head_row(p.head, dl, C, rok ? p.labels[b] : 0, (rok && p.pbar) ? p.pbar + (long long)b * C : nullptr);
          // dlogits (x the F16X3 range scale, a power of two: exact) for pass 2: rows of kDStride floats, slots >= C zero
          float4* dst = reinterpret_cast<float4*>(xb + row * kDStride);
          dst[0] = make_float4(dl[0] * dh_scale, dl[1] * dh_scale, dl[2] * dh_scale, dl[3] * dh_scale);
          dst[1] = make_float4(dl[4] * dh_scale, dl[5] * dh_scale, dl[6] * dh_scale, dl[7] * dh_scale);
          dst[2] = make_float4(dl[8] * dh_scale, dl[9] * dh_scale, 0.f, 0.f);
        }
      }
      epi_bar();
      // sentinel-fill the unused worklist slots of this item
      if (wl) {
        const uint32_t used = min(*wl_count, (uint32_t)kWorkPerItem);
        for (uint32_t i = used + et; i < (uint32_t)kWorkPerItem; i += kEpiThreads) wl[i] = kSentinel;
      }
      if (p.head == -2 && tile_ok) {            // keep mode: the LeakyReLU masks of this item for the gradient pass
        uint32_t* mb = p.maskbuf + item_w * (kMaskWordsItem * kBM) + row;
#pragma unroll
        for (int k = 0; k < kMaskWords; ++k) {
          const int n = k >> 1, ch = 2 * cq + (k & 1);
          if (n < p.n_tiles && ch < nchunks) mb[(n * nchunks + ch) * kBM] = mbits[k];
        }
      }
      if (p.head < 0) continue;

      RBNN_TMARK(c_head)
      // ---------------- pass 2: dH = (dlogits . Wo) * leaky'(H), row-cooperative ----------------
      if (!(p.debug & 2)) {
        uint64_t wq[kCMax][2];                  // the lane's slice of Wo (wo_s is stable until the next item's staging)
        p2_load_wo(wq, wo_s, H, C, g2);
        pass2_rows<MODE>(wq, g2, xb, mbits, quad, m_idx * kBM + quad * 32, p.B, H, (long long)z * p.B, p.dh_hi, p.dh_lo,
                         p.dh_bf, (p.debug & 1) != 0, p.debug);
      }
    }
#ifdef RBNN_FUSED_TIMERS
    if ((p.debug & 8) && blockIdx.x == 0 && et == 0)
      printf("[fused cta0] epilogue warp 2: total %lld cyc: staging %lld, pass 1 %lld (of which waiting for the accumulator %lld), "
             "head %lld, pass 2 %lld\n", clock64() - t_begin, c_stage, c_p1, w_tfull, c_head, c_p2 + (clock64() - t_mark));
#endif
#undef RBNN_TMARK
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                            // no CTA leaves while its peer may still signal / read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<PAIR>(tmem_base);
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
             void* __restrict__ dh_hi_v, void* __restrict__ dh_lo_v, uint32_t* __restrict__ masks) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m_tiles = (B + kBM - 1) / kBM;
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
          // (b, j) -> word j / 32 of row b % 128 of the item, bit j % 32 (the fused kernel's keep-mode layout)
          atomicXor(masks + ((long long)(z * m_tiles + b / kBM) * kMaskWordsItem + (j >> 5)) * kBM + (b % kBM), 1u << (j & 31));
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
// kernel without its GEMM.  One CTA of 512 threads per work item at a time: thread = (input row, column quarter) as in
// the fused epilogue.
// Work item = (sample, 128-input tile, half of the tile): CTAs of 256 threads, two resident per SM, so that one CTA's
// staging of Wo_z overlaps the other's row loop.
constexpr int kKeptThreads = 256;
template <int MODE>
__global__ void __launch_bounds__(kKeptThreads, 2)
dh_from_kept_kernel(int B, int H, int C, int num_items, int m_tiles, int head, const float* __restrict__ bank, long long P,
                    long long wo_off, int z_row0, const int32_t* __restrict__ labels, const float* __restrict__ pbar,
                    const float* __restrict__ logits, const uint32_t* __restrict__ masks, void* dh_hi, void* dh_lo,
                    __nv_bfloat16* dh_bf, const float* __restrict__ dh_scale_p) {
  extern __shared__ float ksm[];
  float* wo_s = ksm;                            // [C][H]
  float* xd = ksm + kCMax * kHMax;              // [128][kDStride] dlogits (only this half's 64 rows are used)
  const int et = threadIdx.x, warp = et >> 5, lane = et & 31;
  const int cq = warp >> 1;                     // as in the fused epilogue: warp = (32-row group, column quarter)
  const int BN = H <= 256 ? H : 256, n_tiles = H <= 256 ? 1 : H / 256, nchunks = BN / 32;
  const P2Lane g2 = p2_lane(lane, cq, BN, n_tiles);
  const float dh_scale = (MODE == MODE_F16X3) ? __ldg(dh_scale_p) : 1.f;
  for (int item2 = blockIdx.x; item2 < 2 * num_items; item2 += gridDim.x) {
    const int item = item2 >> 1, quad = (item2 & 1) * 2 + (warp & 1);
    const int row = quad * 32 + lane;
    const int z = item / m_tiles, m_idx = item % m_tiles;
    const float* __restrict__ wrow = bank + (long long)(z_row0 + z) * P;
    __syncthreads();                            // everyone is done with the previous item's Wo / dlogits
    stage_head_params(wrow, wo_off, 0, 0, C, H, wo_s, nullptr, nullptr, et, kKeptThreads);
    const int b = m_idx * kBM + row;
    const bool rok = b < B;
    if (cq == 0) {
      float dl[kCMax];
      const float* __restrict__ lrow = logits + ((long long)z * B + b) * C;
#pragma unroll
      for (int c = 0; c < kCMax; ++c) dl[c] = (c < C && rok) ? __ldg(lrow + c) : 0.f;
      head_row(head, dl, C, rok ? labels[b] : 0, (rok && pbar) ? pbar + (long long)b * C : nullptr);
      float4* dst = reinterpret_cast<float4*>(xd + row * kDStride);
      dst[0] = make_float4(dl[0] * dh_scale, dl[1] * dh_scale, dl[2] * dh_scale, dl[3] * dh_scale);
      dst[1] = make_float4(dl[4] * dh_scale, dl[5] * dh_scale, dl[6] * dh_scale, dl[7] * dh_scale);
      dst[2] = make_float4(dl[8] * dh_scale, dl[9] * dh_scale, 0.f, 0.f);
    }
    // this thread's mask words (thread = row): chunks 2 cq, 2 cq + 1 of every n-tile, the fused kernel's keep-mode layout
    uint32_t mbits[kMaskWords];
    const uint32_t* __restrict__ mb = masks + (long long)item * (kMaskWordsItem * kBM) + row;
#pragma unroll
    for (int k = 0; k < kMaskWords; ++k) {
      const int n = k >> 1, ch = 2 * cq + (k & 1);
      mbits[k] = (n < n_tiles && ch < nchunks) ? __ldg(mb + (n * nchunks + ch) * kBM) : 0u;
    }
    __syncthreads();
    uint64_t wq[kCMax][2];
    p2_load_wo(wq, wo_s, H, C, g2);
    pass2_rows<MODE>(wq, g2, xd, mbits, quad, m_idx * kBM + quad * 32, B, H, (long long)z * B, dh_hi, dh_lo, dh_bf, false);
  }
}

}  // namespace

bool fused_supported(int H, int C) {
  if (C > kCMax || C < 1 || H < 32 || (H & 31) || H > kHMax) return false;
  if (H > 256 && (H % 256)) return false;
  return true;
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
  static const int env_kbb = getenv("RBNN_FUSED_KBB") ? atoi(getenv("RBNN_FUSED_KBB")) : 0;   // experiments
  const int kbb = env_kbb == 64 ? 64 : (env_kbb == 128 ? 128 : (d.kblock_bytes == 128 ? 128 : 64));
  typedef void (*kern_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const FParams);
  // [mode][runtime class count / 10 classes at compile time][variant: 64-byte K-blocks, 128-byte K-blocks, CTA pairs]
  static const kern_t kerns[3][2][3] = {
      {{fc_fused_kernel<MODE_TF32X3, 64, 0, false>, fc_fused_kernel<MODE_TF32X3, 128, 0, false>, fc_fused_kernel<MODE_TF32X3, 128, 0, true>},
       {fc_fused_kernel<MODE_TF32X3, 64, 10, false>, fc_fused_kernel<MODE_TF32X3, 128, 10, false>, fc_fused_kernel<MODE_TF32X3, 128, 10, true>}},
      {{fc_fused_kernel<MODE_BF16, 64, 0, false>, fc_fused_kernel<MODE_BF16, 128, 0, false>, fc_fused_kernel<MODE_BF16, 128, 0, true>},
       {fc_fused_kernel<MODE_BF16, 64, 10, false>, fc_fused_kernel<MODE_BF16, 128, 10, false>, fc_fused_kernel<MODE_BF16, 128, 10, true>}},
      {{fc_fused_kernel<MODE_F16X3, 64, 0, false>, fc_fused_kernel<MODE_F16X3, 128, 0, false>, fc_fused_kernel<MODE_F16X3, 128, 0, true>},
       {fc_fused_kernel<MODE_F16X3, 64, 10, false>, fc_fused_kernel<MODE_F16X3, 128, 10, false>, fc_fused_kernel<MODE_F16X3, 128, 10, true>}}};
  const int m_tiles = (d.B + kBM - 1) / kBM;
  // CTA pairs pay off once every SM pair has several work items; small problems keep one CTA per tile
  static const int env_pair = getenv("RBNN_FUSED_PAIR") ? atoi(getenv("RBNN_FUSED_PAIR")) : -1;   // experiments
  bool pair = kbb == 128 && m_tiles >= 2 && d.sm_count >= 2 && (long long)m_tiles * d.Z >= 4LL * d.sm_count;
  if (env_pair == 0) pair = false;
  if (env_pair == 1 && kbb == 128 && m_tiles >= 2 && d.sm_count >= 2) pair = true;
  const int ct = d.C == 10, variant = pair ? 2 : (kbb == 128 ? 1 : 0);
  kern_t kern = kerns[d.mode][ct][variant];
  static bool attr_done[3][2][3] = {};
  if (!attr_done[d.mode][ct][variant]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem);
    if (e != cudaSuccess) { *err = std::string("fused_forward_head: cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    attr_done[d.mode][ct][variant] = true;
  }
  FParams p{};
  p.B = d.B; p.D = d.D; p.H = d.H; p.C = d.C; p.Z = d.Z;
  p.BN = d.H <= 256 ? d.H : 256;
  p.n_tiles = d.H <= 256 ? 1 : d.H / 256;
  p.m_tiles = m_tiles;
  p.m_units = pair ? (m_tiles + 1) / 2 : m_tiles;
  p.num_items = p.m_units * d.Z;
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
  p.unscale = d.unscale; p.dh_scale = d.dh_scale; p.xlo_zero = d.xlo_zero;
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
  if (make_map(&mBh, d.W1.hi, dt, d.D, d.H, d.Z, d.W1.ld, d.W1.zstride, pair ? p.BN / 2 : p.BN, kbb, err)) return 1;
  if (!bf16) {
    if (make_map(&mAl, d.X.lo, dt, d.D, d.B, 1, d.X.ld, 0, kBM, kbb, err)) return 1;
    if (make_map(&mBl, d.W1.lo, dt, d.D, d.H, d.Z, d.W1.ld, d.W1.zstride, pair ? p.BN / 2 : p.BN, kbb, err)) return 1;
  } else {
    mAl = mAh;
    mBl = mBh;
  }
  const int units = pair ? d.sm_count / 2 : d.sm_count;
  const int grid_units = p.num_items < units ? p.num_items : units;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pair ? 2 * grid_units : grid_units);
  cfg.blockDim = dim3(kThreadsF);
  cfg.dynamicSmemBytes = kFusedSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  {
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, mAh, mAl, mBh, mBl, p);
    if (le != cudaSuccess) { *err = std::string("fused_forward_head launch: ") + cudaGetErrorString(le); return 1; }
  }
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
                                                      d.dh_hi, d.dh_lo, nullptr);
  else
    fixup_kernel<false, false><<<blocks, 256, 0, st>>>(d.worklist, items, d.x, d.bank, d.P, d.b1_off, d.z_row0, d.B, d.D, d.H,
                                                       d.dh_hi, d.dh_lo, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("fused_fixup launch: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

size_t keep_mask_words(int B, int Z) { return (size_t)Z * ((B + kBM - 1) / kBM) * kMaskWordsItem * kBM; }

int fused_keep_fixup(const FusedDesc& d, uint32_t* masks, cudaStream_t st, std::string* err) {
  if (d.mode == MODE_BF16 || !d.worklist || d.eps <= 0.f || d.B <= 0 || d.Z <= 0) return 0;
  const int items = d.Z * ((d.B + kBM - 1) / kBM);
  const unsigned blocks = (unsigned)std::min(items, d.sm_count * 8);
  fixup_kernel<false, true><<<blocks, 256, 0, st>>>(d.worklist, items, d.x, d.bank, d.P, d.b1_off, d.z_row0, d.B, d.D, d.H,
                                                    nullptr, nullptr, masks);
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
  const int grid = std::min(2 * items, d.sm_count * 2);
  const int smem = (kCMax * kHMax + kBM * kDStride) * 4;
  __nv_bfloat16* bf = reinterpret_cast<__nv_bfloat16*>(d.dh_bf);
#define RBNN_KEPT_LAUNCH(M)                                                                                          \
  do {                                                                                                               \
    static bool attr = false;                                                                                        \
    if (!attr) {                                                                                                     \
      cudaError_t ea = cudaFuncSetAttribute(dh_from_kept_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
      if (ea != cudaSuccess) { *err = std::string("dh_from_kept: cudaFuncSetAttribute: ") + cudaGetErrorString(ea); return 1; } \
      attr = true;                                                                                                   \
    }                                                                                                                \
    dh_from_kept_kernel<M><<<grid, kKeptThreads, smem, st>>>(d.B, d.H, d.C, items, m_tiles, d.head, d.bank, d.P, d.wo_off, \
                                                            d.z_row0, d.labels, d.pbar, d.logits, d.masks, d.dh_hi,   \
                                                            d.dh_lo, bf, d.dh_scale);                                 \
  } while (0)
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
