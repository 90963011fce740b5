// tcgen05 / TMA / TMEM grouped GEMM kernel for sm_100a (see tc_gemm.cuh for what it computes).
//
// Per CTA (persistent, tiles handed out round-robin):
//   warp 0, lane 0 : TMA producer  - cp.async.bulk.tensor.3d of the A / B K-blocks (128 B of K per row,
//                    SWIZZLE_128B) into a ring of shared-memory stages, completion on `full` mbarriers
//   warp 1, lane 0 : MMA issuer    - tcgen05.mma.cta_group::1 (kind::tf32 x3 passes, or kind::f16),
//                    128 x BN x (8|16) per instruction, accumulating in TMEM; tcgen05.commit frees the
//                    stage (`empty`) and, after the last K-block, publishes the accumulator (`tmem_full`)
//   warps 2..5     : epilogue      - tcgen05.ld 32 lanes x 32 columns at a time, bias / LeakyReLU / mask,
//                    fp32 / tf32-split / bf16 stores; `tmem_empty` hands the accumulator stage back
// The accumulator is double buffered (2 x 256 TMEM columns), so the epilogue of tile i overlaps the
// MMAs of tile i+1.
#include "tc_gemm.cuh"

#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <map>
#include <mutex>
#include <tuple>

namespace rbnn {
namespace tc {

constexpr float kSlope = 0.01f;            // nn.LeakyReLU() default (model_nn.py:68-69)
constexpr int kATile = kBM * 128;          // bytes: 128 rows x 128 B
constexpr int kBTile = kBNMax * 128;       // bytes: up to 256 rows x 128 B
constexpr int kRingBytes = 196608;         // 2 stages x 96 KB (TF32X3) or 4 stages x 48 KB (BF16)
constexpr int kSmemBytes = kRingBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;

size_t smem_bytes() { return kSmemBytes; }

struct KParams {
  int M, N, K, Z, BN;
  int m_tiles, n_tiles, num_tiles, num_kb;
  int reduce_z, slots, a_per_z, epi;
  const float* bias; long long bias_zstride;
  const float* act; long long act_zstride, act_ld;
  float* out; float* out_lo; __nv_bfloat16* out_bf;
  long long out_ld, out_zstride;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (=> launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) __trap();   // ~10 s
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  if (BF16) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane = tile row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset between 8-row core groups
  d |= (uint64_t)1 << 46;             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ float to_tf32_rn(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const KParams p) {
  constexpr int NPASS = BF16 ? 1 : 3;
  constexpr int NARR = BF16 ? 1 : 2;                       // arrays per operand (hi[, lo])
  constexpr int STAGE = NARR * (kATile + kBTile);          // 96 KB / 48 KB
  constexpr int NSTAGE = kRingBytes / STAGE;               // 2 / 4
  constexpr int KBE = BF16 ? 64 : 32;                      // elements of K per 128-byte K-block
  constexpr uint32_t FMT = BF16 ? 1u : 2u;                 // UMMA operand format: BF16 = 1, TF32 = 2

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;            // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t bars = ring + kRingBytes;
  const uint32_t full0 = bars, empty0 = bars + 8 * NSTAGE;
  const uint32_t tfull0 = bars + 16 * NSTAGE, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (bars - raw) + 16 * NSTAGE + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, 4);                       // one arrival per epilogue warp
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

  const uint32_t stage_tx = (uint32_t)NARR * (uint32_t)(kATile + p.BN * 128);
  const int tiles_mn = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n_idx = t % p.n_tiles, m_idx = (t / p.n_tiles) % p.m_tiles, zz = t / tiles_mn;
        const int z0 = p.reduce_z ? (int)((long long)zz * p.Z / p.slots) : zz;
        const int z1 = p.reduce_z ? (int)((long long)(zz + 1) * p.Z / p.slots) : zz + 1;
        for (int z = z0; z < z1; ++z) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1u);
            const uint32_t fb = full0 + 8 * stage;
            mbar_expect_tx(fb, stage_tx);
            const uint32_t sa = ring + stage * STAGE;
            const uint32_t sb = sa + NARR * kATile;
            const int za = p.a_per_z ? z : 0;
            tma_load_3d(sa, &tmAh, fb, kb * KBE, m_idx * kBM, za);
            tma_load_3d(sb, &tmBh, fb, kb * KBE, n_idx * p.BN, z);
            if (!BF16) {
              tma_load_3d(sa + kATile, &tmAl, fb, kb * KBE, m_idx * kBM, za);
              tma_load_3d(sb + kBTile, &tmBl, fb, kb * KBE, n_idx * p.BN, z);
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
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int zz = t / tiles_mn;
        const int z0 = p.reduce_z ? (int)((long long)zz * p.Z / p.slots) : zz;
        const int z1 = p.reduce_z ? (int)((long long)(zz + 1) * p.Z / p.slots) : zz + 1;
        const uint32_t as = it & 1u;
        mbar_wait(tempty0 + 8 * as, ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)kBNMax;
        uint32_t accumulate = 0;
        const int total_kb = (z1 - z0) * p.num_kb;
        for (int i = 0; i < total_kb; ++i) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = ring + stage * STAGE;
          const uint32_t sb = sa + NARR * kATile;
          const uint64_t a_hi = smem_desc(sa), b_hi = smem_desc(sb);
          if (BF16) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              tc_mma<true>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, accumulate);
              accumulate = 1;
            }
          } else {
            const uint64_t a_lo = smem_desc(sa + kATile), b_lo = smem_desc(sb + kBTile);
#pragma unroll
            for (int k = 0; k < 4; ++k) {      // small cross terms first, then the leading term
              tc_mma<false>(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, accumulate);
              tc_mma<false>(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
              tc_mma<false>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
              accumulate = 1;
            }
          }
          tc_commit(empty0 + 8 * stage);                   // stage reusable once these MMAs retire
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull0 + 8 * as);                        // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 <-> TMEM lane quadrants 2,3,0,1) =====================
    const int quad = warp & 3;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int n_idx = t % p.n_tiles, m_idx = (t / p.n_tiles) % p.m_tiles, zz = t / tiles_mn;
      const uint32_t as = it & 1u;
      mbar_wait(tfull0 + 8 * as, (it >> 1) & 1u);
      tc_fence_after();
      const int m = m_idx * kBM + quad * 32 + lane;
      const bool row_ok = m < p.M;
      const long long orow = (long long)zz * p.out_zstride + (long long)m * p.out_ld;
      const long long arow = (long long)zz * p.act_zstride + (long long)m * p.act_ld;
      const float* bias = p.bias ? p.bias + (long long)zz * p.bias_zstride : nullptr;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (uint32_t)kBNMax;
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        const int n0 = n_idx * p.BN + c0;
        if (!row_ok) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + j;
          if (c0 + j >= p.BN || n >= p.N) break;           // BN % 16 == 0 and N % 4 == 0: groups of 4 are all-in or all-out
          float v[4] = {__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                        __uint_as_float(r[j + 3])};
          if (p.epi == EPI_BIAS_LEAKY || p.epi == EPI_BIAS) {
            // scalar loads: bank rows are P floats apart, so bias_z is only 4-byte aligned
            v[0] += __ldg(bias + n); v[1] += __ldg(bias + n + 1); v[2] += __ldg(bias + n + 2); v[3] += __ldg(bias + n + 3);
            if (p.epi == EPI_BIAS_LEAKY) {
#pragma unroll
              for (int q = 0; q < 4; ++q) v[q] = v[q] > 0.f ? v[q] : v[q] * kSlope;
            }
          } else if (p.epi == EPI_MASK) {
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
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
static int make_map(CUtensorMap* map, const void* base, bool bf16, int64_t K, int64_t rows, int64_t Z, int64_t ld,
                    int64_t zstride, int box_rows, std::string* err) {
  auto fn = encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled is not available from the driver"; return 1; }
  const int64_t es = bf16 ? 2 : 4;
  if (zstride == 0) Z = 1;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)Z};
  cuuint64_t strides[2] = {(cuuint64_t)(ld * es), (cuuint64_t)((zstride ? zstride : rows * ld) * es)};
  cuuint32_t box[3] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15)) {
    *err = "tc::gemm: operand base / strides must be 16-byte aligned";
    return 1;
  }
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return 1;
  }
  return 0;
}

int gemm(const GemmDesc& d, cudaStream_t st, std::string* err) {
  std::string local;
  if (!err) err = &local;
  const bool bf16 = d.mode == MODE_BF16;
  if (d.M <= 0 || d.N <= 0 || d.K <= 0 || d.Z <= 0) return 0;
  if (d.BN < 16 || d.BN > kBNMax || (d.BN & 15)) { *err = "tc::gemm: BN must be a multiple of 16 in [16,256]"; return 1; }
  if (d.N & 3) { *err = "tc::gemm: N must be a multiple of 4"; return 1; }
  if (!d.A.hi || !d.B.hi || (!bf16 && (!d.A.lo || !d.B.lo))) { *err = "tc::gemm: missing operand array"; return 1; }
  if (d.reduce_z && (d.slots < 1 || d.slots > d.Z)) { *err = "tc::gemm: slots must be in [1, Z]"; return 1; }

  static bool attr_done[2] = {false, false};
  if (!attr_done[bf16]) {
    cudaError_t e = bf16 ? cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)
                         : cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) { *err = std::string("tc::gemm: cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    attr_done[bf16] = true;
  }

  CUtensorMap mAh, mAl, mBh, mBl;
  if (make_map(&mAh, d.A.hi, bf16, d.K, d.A.rows, d.Z, d.A.ld, d.A.zstride, kBM, err)) return 1;
  if (make_map(&mBh, d.B.hi, bf16, d.K, d.B.rows, d.Z, d.B.ld, d.B.zstride, d.BN, err)) return 1;
  if (!bf16) {
    if (make_map(&mAl, d.A.lo, false, d.K, d.A.rows, d.Z, d.A.ld, d.A.zstride, kBM, err)) return 1;
    if (make_map(&mBl, d.B.lo, false, d.K, d.B.rows, d.Z, d.B.ld, d.B.zstride, d.BN, err)) return 1;
  } else {
    mAl = mAh;
    mBl = mBh;
  }

  KParams p{};
  p.M = d.M; p.N = d.N; p.K = d.K; p.Z = d.Z; p.BN = d.BN;
  p.m_tiles = (d.M + kBM - 1) / kBM;
  p.n_tiles = (d.N + d.BN - 1) / d.BN;
  p.reduce_z = d.reduce_z ? 1 : 0;
  p.slots = d.reduce_z ? d.slots : 1;
  p.num_tiles = p.m_tiles * p.n_tiles * (d.reduce_z ? d.slots : d.Z);
  const int kbe = bf16 ? 64 : 32;
  p.num_kb = (d.K + kbe - 1) / kbe;
  p.a_per_z = d.A.zstride != 0;
  p.epi = d.epi;
  p.bias = d.bias; p.bias_zstride = d.bias_zstride;
  p.act = d.act; p.act_zstride = d.act_zstride; p.act_ld = d.act_ld;
  p.out = d.out; p.out_lo = d.out_lo; p.out_bf = reinterpret_cast<__nv_bfloat16*>(d.out_bf);
  p.out_ld = d.out_ld; p.out_zstride = d.out_zstride;
  if ((d.epi == EPI_BIAS_LEAKY || d.epi == EPI_BIAS) && !d.bias) { *err = "tc::gemm: bias epilogue without bias"; return 1; }
  if (d.epi == EPI_MASK && !d.act) { *err = "tc::gemm: mask epilogue without activations"; return 1; }
  if ((d.out_ld & 3) || (d.epi == EPI_MASK && (d.act_ld & 3))) { *err = "tc::gemm: leading dimensions must be multiples of 4"; return 1; }

  const int grid = p.num_tiles < d.sm_count ? p.num_tiles : d.sm_count;
  if (bf16)
    tc_gemm_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(mAh, mAl, mBh, mBl, p);
  else
    tc_gemm_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(mAh, mAl, mBh, mBl, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("tc::gemm launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace tc
}  // namespace rbnn
