// tcgen05 / TMA / TMEM grouped GEMM kernel for sm_100a (see tc_gemm.cuh for what it computes).
//
// Per CTA (persistent, tiles handed out round-robin):
//   warp 0, lane 0 : TMA producer  - cp.async.bulk.tensor.3d of the A / B K-blocks (128 B of K per row,
//                    SWIZZLE_128B) into a ring of shared-memory stages, completion on `full` mbarriers
//   warp 1, lane 0 : MMA issuer    - tcgen05.mma.cta_group::1 (kind::tf32 x3 passes, kind::f16 x3 passes on
//                    scaled fp16 hi/lo operands, or one kind::f16 pass on bf16),
//                    128 x BN x (8|16) per instruction, accumulating in TMEM; tcgen05.commit frees the
//                    stage (`empty`) and, after the last K-block, publishes the accumulator (`tmem_full`)
//   warps 2..5     : epilogue      - tcgen05.ld 32 lanes x 32 columns at a time, bias / LeakyReLU / mask,
//                    fp32 / tf32-split / bf16 stores; `tmem_empty` hands the accumulator stage back
// The accumulator is double buffered (2 x 256 TMEM columns), so the epilogue of tile i overlaps the
// MMAs of tile i+1.
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <map>
#include <mutex>
#include <tuple>

namespace rbnn {
namespace tc {

constexpr float kSlope = 0.01f;            // nn.LeakyReLU() default (model_nn.py:68-69)
// operand ring: as many K-block stages as fit.  A stage holds NARR x (128 + BNT) rows of 128 (64) bytes, BNT = the
// B-tile capacity the kernel was instantiated for: 256 rows -> 2 x 96 KB, 160 rows -> 3 x 72 KB, 128 -> 3 x 64 KB.
// The MMAs of a stage and its TMA refill cannot overlap, so with S stages the K-block period is
// max(T_mma, (T_mma + T_tma) / S): narrower tiles with a third stage beat wide tiles with two.
constexpr int kRingBytes = 221184;
constexpr int kSmemBytes = kRingBytes + 1024 /*alignment slack*/ + 256 /*barriers*/ + 4 * 2048 /*epilogue staging tiles*/;
static_assert(kSmemBytes <= 232448, "tc_gemm_kernel exceeds the 227 KB shared memory of an sm_100 CTA");

size_t smem_bytes() { return kSmemBytes; }

struct KParams {
  int M, N, K, Z, BN;
  int m_tiles, m_units, n_tiles, num_tiles, num_kb;   // m_units = m_tiles (single CTA) or ceil(m_tiles / 2) (CTA pair)
  int reduce_z, slots, a_per_z, epi, skip_mma, relay, spin;
  int conv_a;                 // A tiles are 5-D boxes of a channels-last 12x12x32 map (implicit GEMM of the 5x5 convolution)
  const float* unscale;
  const float* bias; long long bias_zstride;
  const float* act; long long act_zstride, act_ld;
  float* out; float* out_lo; __nv_bfloat16* out_bf;
  long long out_ld, out_zstride;
  unsigned* group_max;
};

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
// PAIR = true: two CTAs of a cluster (a TPC's SM pair) work on one 256-row tile with tcgen05.mma.cta_group::2:
// each CTA stages its own 128 rows of A and HALF of the B tile, the leader CTA issues the MMAs for both,
// each CTA's TMEM receives its 128 accumulator rows.  Per SM that halves the B traffic from L2 and the B reads
// from shared memory, which is what bounds the single-CTA kernel.
template <int MODE, int KBB, bool PAIR, int BNT>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const KParams p) {
  constexpr bool BF16 = MODE == MODE_BF16;                 // single pass, one array per operand
  constexpr bool F16K = MODE != MODE_TF32X3;               // kind::f16 MMAs on 2-byte operands
  constexpr int NARR = BF16 ? 1 : 2;                       // arrays per operand (hi[, lo])
  constexpr int kATile = kBM * KBB;                        // bytes of one A K-block tile (128 rows)
  constexpr int kBTile = (PAIR ? BNT / 2 : BNT) * KBB;     // B K-block tile held by this CTA
  constexpr int STAGE = NARR * (kATile + kBTile);
  constexpr int NSTAGE = (kRingBytes / STAGE) > 8 ? 8 : (kRingBytes / STAGE);
  constexpr int KBE = KBB / (F16K ? 2 : 4);                // elements of K per K-block
  constexpr int KSTEPS = KBB / 32;                         // UMMA K-steps (32 bytes of K each) per K-block
  constexpr uint32_t FMT = MODE == MODE_BF16 ? 1u : (MODE == MODE_F16X3 ? 0u : 2u);   // UMMA operand format: F16 = 0, BF16 = 1, TF32 = 2
  constexpr int NCTA = PAIR ? 2 : 1;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;            // swizzled tiles need 1024 B alignment
  const uint32_t bars = ring + kRingBytes;
  const uint32_t full0 = bars, empty0 = bars + 8 * NSTAGE;
  const uint32_t tfull0 = bars + 16 * NSTAGE, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (bars - raw) + 16 * NSTAGE + 32);
  const uint32_t pfull0 = bars + 16 * NSTAGE + 48;         // pair/relay protocol: "the peer's stage is full" (leader's copy)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs)
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // scheduling unit: CTA or CTA pair
  const int num_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, (PAIR && !p.relay) ? 2 : 1); // direct protocol: one arrival per producer of the pair
      mbar_init(empty0 + 8 * s, 1);
      if (PAIR) mbar_init(pfull0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, 4 * NCTA);                // one arrival per epilogue warp (of both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
    if (!BF16) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
    }
  }
  if (warp == 1) tmem_alloc<PAIR>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                            // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int bn_cta = PAIR ? p.BN / 2 : p.BN;               // B rows staged by this CTA
  const uint32_t stage_tx = (uint32_t)((PAIR && !p.relay) ? 2 * NARR : NARR) * (uint32_t)(kATile + bn_cta * KBB);   // bytes per full barrier
  const int tiles_mn = p.m_units * p.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs of a pair) =====================
      uint32_t stage = 0, phase = 0;
      for (int t = unit; t < p.num_tiles; t += num_units) {
        const int n_idx = t % p.n_tiles, m_unit = (t / p.n_tiles) % p.m_units, zz = t / tiles_mn;
        const int m_idx = PAIR ? 2 * m_unit + (int)rank : m_unit;
        const int z0 = p.reduce_z ? (int)((long long)zz * p.Z / p.slots) : zz;
        const int z1 = p.reduce_z ? (int)((long long)(zz + 1) * p.Z / p.slots) : zz + 1;
        for (int z = z0; z < z1; ++z) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait_mode(empty0 + 8 * stage, phase ^ 1u, p.spin);
            const uint32_t sa = ring + stage * STAGE;
            const uint32_t sb = sa + NARR * kATile;
            const int za = p.a_per_z ? z : 0;
            if (PAIR && p.relay) {
              // relay protocol: every CTA fills its own stage behind its own barrier; the peer's idle MMA warp
              // forwards "full" to the leader with one remote arrive per stage
              const uint32_t fb = full0 + 8 * stage;
              mbar_expect_tx(fb, stage_tx);
              if (!BF16 && KBB == 128 && p.conv_a) {       // implicit GEMM (see the single-CTA branch): this CTA's two images
                int cx, cy;
                if (MODE == MODE_F16X3) { cy = kb / 3; cx = 2 * (kb - 3 * cy); }
                else { cy = kb / 5; cx = kb - 5 * cy; }
                tma_load_5d(sa, &tmAh, fb, 0, cx, cy, 2 * m_idx, z);
                tma_load_5d(sa + kATile, &tmAl, fb, 0, cx, cy, 2 * m_idx, z);
              } else {
                tma_load_3d(sa, &tmAh, fb, kb * KBE, m_idx * kBM, za);
                if (!BF16) tma_load_3d(sa + kATile, &tmAl, fb, kb * KBE, m_idx * kBM, za);
              }
              tma_load_3d(sb, &tmBh, fb, kb * KBE, n_idx * p.BN + (int)rank * bn_cta, z);
              if (!BF16) tma_load_3d(sb + kBTile, &tmBl, fb, kb * KBE, n_idx * p.BN + (int)rank * bn_cta, z);
            } else if (PAIR) {
              const uint32_t fb = mapa_u32(full0 + 8 * stage, 0u);          // the leader's barrier collects both CTAs' bytes
              if (rank == 0) mbar_expect_tx(full0 + 8 * stage, stage_tx);
              else mbar_arrive_cluster(fb);
              tma_load_3d_pair(sa, &tmAh, fb, kb * KBE, m_idx * kBM, za);
              tma_load_3d_pair(sb, &tmBh, fb, kb * KBE, n_idx * p.BN + (int)rank * bn_cta, z);
              if (!BF16) {
                tma_load_3d_pair(sa + kATile, &tmAl, fb, kb * KBE, m_idx * kBM, za);
                tma_load_3d_pair(sb + kBTile, &tmBl, fb, kb * KBE, n_idx * p.BN + (int)rank * bn_cta, z);
              }
            } else if ((MODE == MODE_TF32X3 && KBB == 128 || MODE == MODE_F16X3) && p.conv_a) {
              // implicit GEMM: K-block kb = filter tap (ky, kx); the 128 tile rows are the 8x8 output positions of two
              // images, i.e. the box {32 channels, x in [kx, kx+8), y in [ky, ky+8), images 2 m_idx .. +1} of the map
              // (32 channels = 128 bytes of fp32 / tf32, 64 bytes of fp16).  F16X3 with 128-byte K-blocks: kb = (ky, tap
              // pair), the box is {64 elements = pixels x, x + 1} at x offset 0 / 2 / 4 of the pixel-pair layout.
              const uint32_t fb = full0 + 8 * stage;
              int ky, kx;
              if (MODE == MODE_F16X3 && KBB == 128) { ky = kb / 3; kx = 2 * (kb - 3 * ky); }
              else { ky = kb / 5; kx = kb - 5 * ky; }
              mbar_expect_tx(fb, stage_tx);
              tma_load_5d(sa, &tmAh, fb, 0, kx, ky, 2 * m_idx, z);
              tma_load_3d(sb, &tmBh, fb, kb * KBE, n_idx * p.BN, z);
              tma_load_5d(sa + kATile, &tmAl, fb, 0, kx, ky, 2 * m_idx, z);
              tma_load_3d(sb + kBTile, &tmBl, fb, kb * KBE, n_idx * p.BN, z);
            } else {
              const uint32_t fb = full0 + 8 * stage;
              mbar_expect_tx(fb, stage_tx);
              tma_load_3d(sa, &tmAh, fb, kb * KBE, m_idx * kBM, za);
              tma_load_3d(sb, &tmBh, fb, kb * KBE, n_idx * p.BN, z);
              if (!BF16) {
                tma_load_3d(sa + kATile, &tmAl, fb, kb * KBE, m_idx * kBM, za);
                tma_load_3d(sb + kBTile, &tmBl, fb, kb * KBE, n_idx * p.BN, z);
              }
            }
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (PAIR && p.relay && lane == 0 && rank == 1) {
      // ===================== peer relay: own stage full -> tell the leader =====================
      const uint32_t pf = mapa_u32(pfull0, 0u);
      uint32_t stage = 0, phase = 0;
      for (int t = unit; t < p.num_tiles; t += num_units) {
        const int zz = t / tiles_mn;
        const int z0 = p.reduce_z ? (int)((long long)zz * p.Z / p.slots) : zz;
        const int z1 = p.reduce_z ? (int)((long long)(zz + 1) * p.Z / p.slots) : zz + 1;
        const int total_kb = (z1 - z0) * p.num_kb;
        for (int i = 0; i < total_kb; ++i) {
          mbar_wait_mode(full0 + 8 * stage, phase, p.spin);
          mbar_arrive_cluster(pf + 8 * stage);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
    if (lane == 0 && rank == 0) {
      // ===================== MMA issuer (leader CTA only) =====================
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)((kBM * NCTA) >> 4) << 24);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int t = unit; t < p.num_tiles; t += num_units, ++it) {
        const int zz = t / tiles_mn;
        const int z0 = p.reduce_z ? (int)((long long)zz * p.Z / p.slots) : zz;
        const int z1 = p.reduce_z ? (int)((long long)(zz + 1) * p.Z / p.slots) : zz + 1;
        const uint32_t as = it & 1u;
        mbar_wait_mode(tempty0 + 8 * as, ((it >> 1) & 1u) ^ 1u, p.spin);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)kBNMax;
        uint32_t accumulate = 0;
        const int total_kb = (z1 - z0) * p.num_kb;
        for (int i = 0; i < total_kb; ++i) {
          mbar_wait_mode(full0 + 8 * stage, phase, p.spin);
          if (PAIR && p.relay) mbar_wait_mode(pfull0 + 8 * stage, phase, p.spin);
          tc_fence_after();
          const uint32_t sa = ring + stage * STAGE;
          const uint32_t sb = sa + NARR * kATile;
          const uint64_t a_hi = smem_desc<KBB>(sa), b_hi = smem_desc<KBB>(sb);
          if (p.skip_mma) {
            // debug: operand pipeline only (measures what TMA alone sustains)
          } else if (BF16) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              if (PAIR) tc_mma_pair<true>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, accumulate);
              else tc_mma<true>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, accumulate);
              accumulate = 1;
            }
          } else {
            const uint64_t a_lo = smem_desc<KBB>(sa + kATile), b_lo = smem_desc<KBB>(sb + kBTile);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {      // small cross terms first, then the leading term
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
          }
          // stage reusable (in both CTAs) once these MMAs retire
          if (PAIR) tc_commit_pair(empty0 + 8 * stage, (uint16_t)3);
          else tc_commit(empty0 + 8 * stage);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
        if (PAIR) tc_commit_pair(tfull0 + 8 * as, (uint16_t)3);   // accumulator complete (both CTAs' epilogues)
        else tc_commit(tfull0 + 8 * as);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 <-> TMEM lane quadrants 2,3,0,1) =====================
    const int quad = warp & 3;
    const uint32_t tempty_leader = PAIR ? mapa_u32(tempty0, 0u) : tempty0;
    const float unscale = p.unscale ? __ldg(p.unscale) : 1.f;   // F16X3: 1 / (operand scales), a power of two
    float* stg = reinterpret_cast<float*>(smem_raw + (bars - raw) + 256) + (warp - 2) * 512;   // 32 rows x 16 floats per warp
    uint32_t it = 0;
    for (int t = unit; t < p.num_tiles; t += num_units, ++it) {
      const int n_idx = t % p.n_tiles, m_unit = (t / p.n_tiles) % p.m_units, zz = t / tiles_mn;
      const int m_idx = PAIR ? 2 * m_unit + (int)rank : m_unit;
      const uint32_t as = it & 1u;
      mbar_wait_mode(tfull0 + 8 * as, (it >> 1) & 1u, p.spin);
      tc_fence_after();
      // The accumulator arrives with lane = tile row (tcgen05.ld 32x32b).  Stored that way a warp's 16-byte stores hit 32
      // different lines per instruction and the bias loads sit between them, one dependent load per 4 columns (the
      // conv2 epilogue took 41 k cycles per tile against 23 k of MMAs).  So 16 columns at a time go through a 2 KB
      // per-warp staging tile (16-byte chunks XOR-swizzled: conflict-free both ways) and come back transposed: lane =
      // (row srow + 8 i, columns 4 sk .. 4 sk + 3), i.e. 4 lanes write 64 contiguous bytes of a row, the bias of a lane's
      // 4 columns is loaded once per 16-column group (and prefetched one group ahead), the mask operand of EPI_MASK is
      // read with the same coalesced pattern.
      const int m_base = m_idx * kBM + quad * 32;
      const int srow = lane >> 2, sk = lane & 3;
      const float* __restrict__ bias = p.bias ? p.bias + (long long)zz * p.bias_zstride : nullptr;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (uint32_t)kBNMax;
      const bool has_bias = p.epi == EPI_BIAS_LEAKY || p.epi == EPI_BIAS;
      float gmx = 0.f;                                   // max |pre-activation| seen by this lane (p.group_max)
      float bnext[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const int n = n_idx * p.BN + 4 * sk;
        if (has_bias && 4 * sk < p.BN && n < p.N) {
#pragma unroll
          for (int q = 0; q < 4; ++q) bnext[q] = __ldg(bias + n + q);   // scalar: bank rows are only 4-byte aligned
        }
      }
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          const int cc = c0 + 16 * hlf + 4 * sk;           // this lane's first column inside the tile
          const int n = n_idx * p.BN + cc;
          const bool col_ok = cc < p.BN && n < p.N;        // BN % 16 == 0 and N % 4 == 0: groups of 4 are all-in or all-out
          float bv[4] = {bnext[0], bnext[1], bnext[2], bnext[3]};
          if (has_bias && cc + 16 < p.BN && n + 16 < p.N) {
#pragma unroll
            for (int q = 0; q < 4; ++q) bnext[q] = __ldg(bias + n + 16 + q);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(stg + lane * 16 + ((k ^ ((lane >> 1) & 3)) << 2)) =
                make_uint4(r[16 * hlf + 4 * k], r[16 * hlf + 4 * k + 1], r[16 * hlf + 4 * k + 2], r[16 * hlf + 4 * k + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = srow + 8 * i;
            const uint4 t = *reinterpret_cast<const uint4*>(stg + rr * 16 + ((sk ^ ((rr >> 1) & 3)) << 2));
            const int m = m_base + rr;
            if (!col_ok || m >= p.M) continue;
            const long long orow = (long long)zz * p.out_zstride + (long long)m * p.out_ld;
            float v[4] = {__uint_as_float(t.x), __uint_as_float(t.y), __uint_as_float(t.z), __uint_as_float(t.w)};
            if (MODE == MODE_F16X3) { v[0] *= unscale; v[1] *= unscale; v[2] *= unscale; v[3] *= unscale; }
            if (has_bias) {
              v[0] += bv[0]; v[1] += bv[1]; v[2] += bv[2]; v[3] += bv[3];
              if (p.epi == EPI_BIAS_LEAKY) {
                gmx = fmaxf(fmaxf(gmx, fmaxf(fabsf(v[0]), fabsf(v[1]))), fmaxf(fabsf(v[2]), fabsf(v[3])));
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = v[q] > 0.f ? v[q] : v[q] * kSlope;
              }
            } else if (p.epi == EPI_MASK) {
              const long long arow = (long long)zz * p.act_zstride + (long long)m * p.act_ld;
              const float4 h = __ldg(reinterpret_cast<const float4*>(p.act + arow + n));
              v[0] = h.x > 0.f ? v[0] : v[0] * kSlope;
              v[1] = h.y > 0.f ? v[1] : v[1] * kSlope;
              v[2] = h.z > 0.f ? v[2] : v[2] * kSlope;
              v[3] = h.w > 0.f ? v[3] : v[3] * kSlope;
            }
            if (p.out_lo) {
              float hi[4], lo[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) { hi[q] = to_tf32_rn(v[q]); lo[q] = v[q] - hi[q]; }
              *reinterpret_cast<float4*>(p.out + orow + n) = make_float4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<float4*>(p.out_lo + orow + n) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            } else if (p.out) {
              *reinterpret_cast<float4*>(p.out + orow + n) = make_float4(v[0], v[1], v[2], v[3]);
            }
            if (p.out_bf) {
              __nv_bfloat162 b01 = __floats2bfloat162_rn(v[0], v[1]);
              __nv_bfloat162 b23 = __floats2bfloat162_rn(v[2], v[3]);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&b01);
              pk.y = *reinterpret_cast<uint32_t*>(&b23);
              *reinterpret_cast<uint2*>(p.out_bf + orow + n) = pk;
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_relaxed_cluster(tempty_leader + 8 * as);
        else mbar_arrive_relaxed(tempty0 + 8 * as);
      }
      if (p.group_max) {           // the 32 rows of a warp lie in one group of 64 (tile rows start at multiples of 128)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmx = fmaxf(gmx, __shfl_xor_sync(0xffffffffu, gmx, o));
        if (lane == 0 && m_base < p.M && gmx == gmx)
          atomicMax(p.group_max + (long long)zz * ((p.M + 63) / 64) + (m_base >> 6), __float_as_uint(gmx));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                            // no CTA leaves while its peer may still signal / read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<PAIR>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}

// 3-D map over [Z][rows][K] (K innermost); box = one K-block (128 B) x box_rows x 1; zero fill out of bounds.
int make_map(CUtensorMap* map, const void* base, int dtype, int64_t K, int64_t rows, int64_t Z, int64_t ld,
             int64_t zstride, int box_rows, int kb_bytes, std::string* err) {
  auto fn = encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled is not available from the driver"; return 1; }
  const int64_t es = dtype == DT_F32 ? 4 : 2;
  if (zstride == 0) Z = 1;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)Z};
  cuuint64_t strides[2] = {(cuuint64_t)(ld * es), (cuuint64_t)((zstride ? zstride : rows * ld) * es)};
  cuuint32_t box[3] = {(cuuint32_t)(kb_bytes / es), (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15)) {
    *err = "tc::gemm: operand base / strides must be 16-byte aligned";
    return 1;
  }
  const CUtensorMapDataType dt = dtype == DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, 3,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  kb_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return 1;
  }
  return 0;
}

// 5-D map over channels-last activations [Z][images][12][12][32] fp32: box = {32 channels, 8, 8, 2 images, 1}, i.e. the
// 128 rows x 128 bytes of one filter tap of the implicit GEMM (rows ordered image, oy, ox), SWIZZLE_128B.
static int make_map_conv_a(CUtensorMap* map, const void* base, int dtype, int kb_bytes, int64_t images, int64_t Z, std::string* err) {
  auto fn = encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled is not available from the driver"; return 1; }
  // fp32: 32 channels = 128 bytes per pixel.  fp16, kb_bytes 64: 32 channels = 64 bytes per pixel (one tap per K-block);
  // fp16, kb_bytes 128: the pixel-pair layout of p1_split_hwc_kernel, 64 elements = 128 bytes per pixel (two taps)
  const bool pairs = dtype != DT_F32 && kb_bytes == 128;
  const cuuint64_t px = (dtype == DT_F32 || pairs) ? 128u : 64u;      // bytes between pixels
  cuuint64_t dims[5] = {pairs ? 64u : 32u, 12u, 12u, (cuuint64_t)images, (cuuint64_t)Z};
  cuuint64_t strides[4] = {px, 12u * px, 144u * px, (cuuint64_t)images * 144u * px};
  cuuint32_t box[5] = {pairs ? 64u : 32u, 8u, 8u, 2u, 1u};
  cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  if (reinterpret_cast<uintptr_t>(base) & 127) { *err = "tc::gemm: conv activations must be 128-byte aligned"; return 1; }
  CUresult r = fn(map, dtype == DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  px == 128u ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled (conv activations) failed with CUresult " + std::to_string((int)r);
    return 1;
  }
  return 0;
}

int gemm(const GemmDesc& d, cudaStream_t st, std::string* err) {
  std::string local;
  if (!err) err = &local;
  const bool bf16 = d.mode == MODE_BF16, f16x3 = d.mode == MODE_F16X3;
  const int dt = bf16 ? DT_BF16 : (f16x3 ? DT_F16 : DT_F32);
  if (d.M <= 0 || d.N <= 0 || d.K <= 0 || d.Z <= 0) return 0;
  if (d.mode != MODE_TF32X3 && d.mode != MODE_BF16 && d.mode != MODE_F16X3) { *err = "tc::gemm: unknown mode"; return 1; }
  if (f16x3 && !d.unscale) { *err = "tc::gemm: F16X3 needs the unscale scalar"; return 1; }
  if (f16x3 && d.out_lo) { *err = "tc::gemm: F16X3 has no pre-split output"; return 1; }
  if (d.BN < 16 || d.BN > kBNMax || (d.BN & 15)) { *err = "tc::gemm: BN must be a multiple of 16 in [16,256]"; return 1; }
  if (d.N & 3) { *err = "tc::gemm: N must be a multiple of 4"; return 1; }
  if (!d.A.hi || !d.B.hi || (!bf16 && (!d.A.lo || !d.B.lo))) { *err = "tc::gemm: missing operand array"; return 1; }
  if (d.reduce_z && (d.slots < 1 || d.slots > d.Z)) { *err = "tc::gemm: slots must be in [1, Z]"; return 1; }

  const int kbb = d.kblock_bytes == 128 ? 128 : 64;
  // implicit-GEMM conv operand: TF32X3 / 128-byte K-blocks and F16X3 / 64-byte K-blocks walk the 25 taps (K = 800);
  // F16X3 / 128-byte K-blocks walk 15 tap PAIRS (ky, {0,1} {2,3} {4,pad}) of the pixel-pair layout: K = 960, the B rows
  // carry zeros for the padding tap
  const bool conv_pairs = d.conv_images > 0 && d.mode == MODE_F16X3 && kbb == 128;
  if (d.conv_images > 0 && (!((d.mode == MODE_TF32X3 && kbb == 128) || d.mode == MODE_F16X3) ||
                            d.K != (conv_pairs ? 960 : 800) || d.reduce_z || d.M != d.conv_images * 64 ||
                            (d.pair && ((d.conv_images & 3) || !d.pair_relay)))) {
    *err = "tc::gemm: the implicit-GEMM conv operand needs TF32X3 with 128-byte K-blocks or F16X3 (K = 800; K = 960 with "
           "128-byte K-blocks of two taps), M = 64 * images, CTA pairs only with images % 4 == 0 and the relay protocol";
    return 1;
  }
  const bool pair = d.pair && (d.BN % 32 == 0 || d.BN % 16 == 0) && ((d.BN / 2) % 8 == 0) && d.sm_count >= 2;
  typedef void (*kern_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const KParams);
  static const kern_t kerns[3][2][2] = {
      {{tc_gemm_kernel<MODE_TF32X3, 64, false, 256>, tc_gemm_kernel<MODE_TF32X3, 64, true, 256>},
       {tc_gemm_kernel<MODE_TF32X3, 128, false, 256>, tc_gemm_kernel<MODE_TF32X3, 128, true, 256>}},
      {{tc_gemm_kernel<MODE_BF16, 64, false, 256>, tc_gemm_kernel<MODE_BF16, 64, true, 256>},
       {tc_gemm_kernel<MODE_BF16, 128, false, 256>, tc_gemm_kernel<MODE_BF16, 128, true, 256>}},
      {{tc_gemm_kernel<MODE_F16X3, 64, false, 256>, tc_gemm_kernel<MODE_F16X3, 64, true, 256>},
       {tc_gemm_kernel<MODE_F16X3, 128, false, 256>, tc_gemm_kernel<MODE_F16X3, 128, true, 256>}}};
  // narrow-tile instantiations (single CTA, 128 B K-blocks): a smaller B-tile capacity buys more pipeline stages
  static const kern_t narrow[3][3] = {
      {tc_gemm_kernel<MODE_TF32X3, 128, false, 112>, tc_gemm_kernel<MODE_TF32X3, 128, false, 128>, tc_gemm_kernel<MODE_TF32X3, 128, false, 160>},
      {tc_gemm_kernel<MODE_BF16, 128, false, 112>, tc_gemm_kernel<MODE_BF16, 128, false, 128>, tc_gemm_kernel<MODE_BF16, 128, false, 160>},
      {tc_gemm_kernel<MODE_F16X3, 128, false, 112>, tc_gemm_kernel<MODE_F16X3, 128, false, 128>, tc_gemm_kernel<MODE_F16X3, 128, false, 160>}};
  kern_t kern = kerns[d.mode][kbb == 128][pair];
  if (!pair && kbb == 128 && d.BN <= 160) kern = narrow[d.mode][d.BN <= 112 ? 0 : (d.BN <= 128 ? 1 : 2)];
  {
    static std::map<kern_t, bool> attr_set;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!attr_set[kern]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      if (e != cudaSuccess) { *err = std::string("tc::gemm: cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
      attr_set[kern] = true;
    }
  }

  CUtensorMap mAh, mAl, mBh, mBl;
  if (d.conv_images > 0) {
    if (make_map_conv_a(&mAh, d.A.hi, dt, kbb, d.conv_images, d.Z, err)) return 1;
  } else if (make_map(&mAh, d.A.hi, dt, d.K, d.A.rows, d.Z, d.A.ld, d.A.zstride, kBM, kbb, err)) return 1;
  if (make_map(&mBh, d.B.hi, dt, d.K, d.B.rows, d.Z, d.B.ld, d.B.zstride, pair ? d.BN / 2 : d.BN, kbb, err)) return 1;
  if (!bf16) {
    if (d.conv_images > 0) {
      if (make_map_conv_a(&mAl, d.A.lo, dt, kbb, d.conv_images, d.Z, err)) return 1;
    } else if (make_map(&mAl, d.A.lo, dt, d.K, d.A.rows, d.Z, d.A.ld, d.A.zstride, kBM, kbb, err)) return 1;
    if (make_map(&mBl, d.B.lo, dt, d.K, d.B.rows, d.Z, d.B.ld, d.B.zstride, pair ? d.BN / 2 : d.BN, kbb, err)) return 1;
  } else {
    mAl = mAh;
    mBl = mBh;
  }

  KParams p{};
  p.M = d.M; p.N = d.N; p.K = d.K; p.Z = d.Z; p.BN = d.BN;
  p.m_tiles = (d.M + kBM - 1) / kBM;
  p.m_units = pair ? (p.m_tiles + 1) / 2 : p.m_tiles;
  p.n_tiles = (d.N + d.BN - 1) / d.BN;
  p.reduce_z = d.reduce_z ? 1 : 0;
  p.slots = d.reduce_z ? d.slots : 1;
  p.num_tiles = p.m_units * p.n_tiles * (d.reduce_z ? d.slots : d.Z);
  const int kbe = kbb / (dt == DT_F32 ? 4 : 2);
  p.num_kb = (d.K + kbe - 1) / kbe;
  p.a_per_z = d.A.zstride != 0;
  p.conv_a = d.conv_images > 0;
  p.epi = d.epi;
  p.skip_mma = d.debug_skip_mma;
  p.relay = d.pair_relay;
  p.spin = d.spin_wait;
  p.unscale = f16x3 ? d.unscale : nullptr;
  p.bias = d.bias; p.bias_zstride = d.bias_zstride;
  p.act = d.act; p.act_zstride = d.act_zstride; p.act_ld = d.act_ld;
  p.out = d.out; p.out_lo = d.out_lo; p.out_bf = reinterpret_cast<__nv_bfloat16*>(d.out_bf);
  p.out_ld = d.out_ld; p.out_zstride = d.out_zstride;
  p.group_max = d.epi == EPI_BIAS_LEAKY ? d.group_max : nullptr;
  if ((d.epi == EPI_BIAS_LEAKY || d.epi == EPI_BIAS) && !d.bias) { *err = "tc::gemm: bias epilogue without bias"; return 1; }
  if (d.epi == EPI_MASK && !d.act) { *err = "tc::gemm: mask epilogue without activations"; return 1; }
  if ((d.out_ld & 3) || (d.epi == EPI_MASK && (d.act_ld & 3))) { *err = "tc::gemm: leading dimensions must be multiples of 4"; return 1; }

  const int units = pair ? d.sm_count / 2 : d.sm_count;
  const int grid_units = p.num_tiles < units ? p.num_tiles : units;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pair ? 2 * grid_units : grid_units);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
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
    if (le != cudaSuccess) { *err = std::string("tc::gemm launch: ") + cudaGetErrorString(le); return 1; }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("tc::gemm launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace tc
}  // namespace rbnn
