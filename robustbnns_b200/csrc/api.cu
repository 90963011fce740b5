// C ABI of librbnn.so (include/rbnn.h): handle management, the posterior-sample bank, and the
// per-architecture orchestration of the forward / input-gradient passes over bank rows.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace rbnn {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

constexpr size_t kTailBytes = (size_t)64 << 20;  // split-z partials of gemm_simt live at the arena's end

int ws_reserve(rbnn_net* net, size_t bytes) {
  bytes += kTailBytes;
  if (net->ws_bytes >= bytes) return 0;
  if (net->ws) {
    RBNN_CUDA(cudaDeviceSynchronize());
    RBNN_CUDA(cudaFree(net->ws));
    net->ws = nullptr;
    net->ws_bytes = 0;
  }
  RBNN_CUDA(cudaMalloc(&net->ws, bytes));
  net->ws_bytes = bytes;
  net->alloc_epoch++;
  return 0;
}

int timing_begin(rbnn_net* n, int cls, cudaStream_t st) {
  if (!n->timing || cls <= 0 || cls > 2) return 0;
  cudaEvent_t e;
  RBNN_CUDA(cudaEventCreate(&e));
  RBNN_CUDA(cudaEventRecord(e, st));
  n->ev[cls][0].push_back(e);
  return 0;
}

int timing_end(rbnn_net* n, int cls, cudaStream_t st) {
  if (!n->timing || cls <= 0 || cls > 2) return 0;
  cudaEvent_t e;
  RBNN_CUDA(cudaEventCreate(&e));
  RBNN_CUDA(cudaEventRecord(e, st));
  n->ev[cls][1].push_back(e);
  return 0;
}

static int build_layout(rbnn_net* n) {
  ParamLayout& L = n->L;
  const int64_t H = n->H, D = n->D, C = n->C;
  int64_t o = 0;
  if (n->arch == RBNN_ARCH_FC || n->arch == RBNN_ARCH_FC2) {
    L.w1 = o; o += H * D;
    L.b1 = o; o += H;
    if (n->arch == RBNN_ARCH_FC2) {
      L.w2 = o; o += H * H;
      L.b2 = o; o += H;
    }
    L.wo = o; o += C * H;
    L.bo = o; o += C;
  } else {
    L.cw1 = o; o += 32 * 25;
    L.cb1 = o; o += 32;
    L.cw2 = o; o += H * 800;
    L.cb2 = o; o += H;
    L.wo = o; o += C * 49 * H;
    L.bo = o; o += C;
  }
  L.P = o;
  return 0;
}

// ------------------------------------------------------------------------------------------
// fc / fc2 on the CUDA-core engine
// ------------------------------------------------------------------------------------------
struct FcBufs {
  float *h1, *h2, *logits, *dlogits, *dtop, *dh1;
};

static size_t fc_bytes_per_z(const rbnn_net* n, int B, bool grad) {
  const bool two = n->arch == RBNN_ARCH_FC2;
  size_t per = pad256((size_t)B * n->H * 4) * (two ? 2 : 1) + pad256((size_t)B * n->C * 4);
  if (grad) per += pad256((size_t)B * n->C * 4) + pad256((size_t)B * n->H * 4) * (two ? 2 : 1);
  return per;
}

static int fc_forward_chunk(rbnn_net* n, const float* x, int B, int z0, int Z, FcBufs& f, cudaStream_t st) {
  const int H = n->H, D = n->D, C = n->C;
  const int64_t P = n->L.P;
  const float* rows = n->bank + (int64_t)z0 * P;
  GemmArgs g{};
  g.A = x; g.lda = D; g.sAz = 0;
  g.B = rows + n->L.w1; g.ldb = D; g.sBz = P;
  g.bias = rows + n->L.b1; g.sbz = P;
  g.C = f.h1; g.ldc = H; g.sCz = (int64_t)B * H;
  g.M = B; g.N = H; g.K = D; g.Z = Z; g.epi = EPI_BIAS_LEAKY; g.act = n->act;
  g.tag = 1;
  RBNN_TRY(gemm_simt(n, g, st));
  g.tag = 0;
  const float* top = f.h1;
  if (n->arch == RBNN_ARCH_FC2) {
    g.A = f.h1; g.lda = H; g.sAz = (int64_t)B * H;
    g.B = rows + n->L.w2; g.ldb = H;
    g.bias = rows + n->L.b2;
    g.C = f.h2; g.K = H;
    RBNN_TRY(gemm_simt(n, g, st));
    top = f.h2;
  }
  g.A = top; g.lda = H; g.sAz = (int64_t)B * H;
  g.B = rows + n->L.wo; g.ldb = H;
  g.bias = rows + n->L.bo;
  g.C = f.logits; g.ldc = C; g.sCz = (int64_t)B * C;
  g.N = C; g.K = H; g.epi = EPI_BIAS;
  RBNN_TRY(gemm_simt(n, g, st));
  return 0;
}

static int fc_grad_simt(rbnn_net* n, int head, const float* x, const int32_t* labels, int B, int s0, int s1,
                        const float* pbar, float* out_sum, cudaStream_t st) {
  const int H = n->H, D = n->D, C = n->C;
  const int64_t P = n->L.P;
  const bool two = n->arch == RBNN_ARCH_FC2;
  const size_t per = fc_bytes_per_z(n, B, true);
  const int zc = (int)std::max<size_t>(1, std::min<size_t>(n->ws_budget / per, (size_t)(s1 - s0)));
  RBNN_TRY(ws_reserve(n, per * zc));
  bool first = true;
  for (int z0 = s0; z0 < s1; z0 += zc) {
    const int Z = std::min(zc, s1 - z0);
    Arena ar(n);
    FcBufs f{};
    f.h1 = ar.take<float>((size_t)Z * B * H);
    f.h2 = two ? ar.take<float>((size_t)Z * B * H) : nullptr;
    f.logits = ar.take<float>((size_t)Z * B * C);
    f.dlogits = ar.take<float>((size_t)Z * B * C);
    f.dtop = ar.take<float>((size_t)Z * B * H);
    f.dh1 = two ? ar.take<float>((size_t)Z * B * H) : nullptr;
    RBNN_TRY(fc_forward_chunk(n, x, B, z0, Z, f, st));
    RBNN_TRY(head_dlogits(n, head, f.logits, labels, pbar, Z, B, C, f.dlogits, st));
    const float* rows = n->bank + (int64_t)z0 * P;
    GemmArgs g{};
    g.b_kn = 1;
    g.A = f.dlogits; g.lda = C; g.sAz = (int64_t)B * C;
    g.B = rows + n->L.wo; g.ldb = H; g.sBz = P;
    g.C = f.dtop; g.ldc = H; g.sCz = (int64_t)B * H;
    g.mask = two ? f.h2 : f.h1; g.ldm = H; g.sMz = (int64_t)B * H;
    g.M = B; g.N = H; g.K = C; g.Z = Z; g.epi = EPI_MASK; g.act = n->act;
    RBNN_TRY(gemm_simt(n, g, st));
    const float* dfirst = f.dtop;
    if (two) {
      g.A = f.dtop; g.lda = H; g.sAz = (int64_t)B * H;
      g.B = rows + n->L.w2; g.ldb = H;
      g.C = f.dh1;
      g.mask = f.h1;
      g.K = H;
      RBNN_TRY(gemm_simt(n, g, st));
      dfirst = f.dh1;
    }
    GemmArgs r{};
    r.b_kn = 1;
    r.A = dfirst; r.lda = H; r.sAz = (int64_t)B * H;
    r.B = rows + n->L.w1; r.ldb = D; r.sBz = P;
    r.C = out_sum; r.ldc = D;
    r.M = B; r.N = D; r.K = H; r.Z = Z; r.epi = EPI_NONE;
    r.reduce_z = 1; r.accumulate = first ? 0 : 1;
    r.tag = 2;
    RBNN_TRY(gemm_simt(n, r, st));
    first = false;
  }
  return 0;
}

static int fc_probs_simt(rbnn_net* n, const float* x, int B, int s0, int s1, float* out_sum, float* out_logits,
                         cudaStream_t st) {
  const int H = n->H, C = n->C;
  const bool two = n->arch == RBNN_ARCH_FC2;
  const size_t per = fc_bytes_per_z(n, B, false);
  const int zc = (int)std::max<size_t>(1, std::min<size_t>(n->ws_budget / per, (size_t)(s1 - s0)));
  RBNN_TRY(ws_reserve(n, per * zc));
  for (int z0 = s0; z0 < s1; z0 += zc) {
    const int Z = std::min(zc, s1 - z0);
    Arena ar(n);
    FcBufs f{};
    f.h1 = ar.take<float>((size_t)Z * B * H);
    f.h2 = two ? ar.take<float>((size_t)Z * B * H) : nullptr;
    f.logits = out_logits ? out_logits : ar.take<float>((size_t)Z * B * C);
    RBNN_TRY(fc_forward_chunk(n, x, B, z0, Z, f, st));
    if (out_sum) RBNN_TRY(head_probs_accumulate(n, f.logits, Z, B, C, out_sum, st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// conv on the CUDA-core engine (im2col + GEMM for conv2)
// ------------------------------------------------------------------------------------------
struct ConvBufs {
  float *p1, *col, *a2, *logits, *lpart, *dlogits, *dz2, *g1, *partial;
  uint8_t* idx1;
  int parts;
};

static size_t conv_bytes_per_zb(const rbnn_net* n, bool grad) {
  const size_t H = n->H, C = n->C;
  size_t per = 4608 * 4 + 4608 + 64 * 800 * 4 + 64 * H * 4 + C * 4 + (H / 128 + 1) * C * 4;
  if (grad) per += C * 4 + 64 * H * 4 + 4608 * 4 + 784 * 4 /* share of the conv1-backward partials */;
  return per + 64;
}

static void conv_carve(rbnn_net* n, Arena& ar, int Z, int B, bool grad, ConvBufs& c) {
  const size_t H = n->H, C = n->C, ZB = (size_t)Z * B;
  c.partial = nullptr; c.parts = 1;
  c.p1 = ar.take<float>((size_t)ZB * 4608);
  c.idx1 = ar.take<uint8_t>((size_t)ZB * 4608);
  c.col = ar.take<float>((size_t)ZB * 64 * 800);
  c.a2 = ar.take<float>((size_t)ZB * 64 * H);
  c.logits = ar.take<float>((size_t)ZB * C);
  c.lpart = ar.take<float>((size_t)pool2_logits_chunks(n) * ZB * C);
  if (grad) {
    c.dlogits = ar.take<float>((size_t)ZB * C);
    c.dz2 = ar.take<float>((size_t)ZB * 64 * H);
    c.g1 = ar.take<float>((size_t)ZB * 4608);
    c.parts = conv1_bwd_parts(n, Z, B);
    if (c.parts > 1) c.partial = ar.take<float>((size_t)c.parts * B * 784);
  }
}

static int conv_forward_chunk(rbnn_net* n, const float* x, int B, int z0, int Z, ConvBufs& c, float* logits,
                              cudaStream_t st) {
  const int H = n->H;
  const int64_t P = n->L.P;
  const float* rows = n->bank + (int64_t)z0 * P;
  RBNN_TRY(conv1_pool_fwd(n, x, n->bank, z0, Z, B, c.p1, c.idx1, st));
  RBNN_TRY(im2col_conv2(n, c.p1, Z * B, c.col, st));
  GemmArgs g{};
  g.A = c.col; g.lda = 800; g.sAz = (int64_t)B * 64 * 800;
  g.B = rows + n->L.cw2; g.ldb = 800; g.sBz = P;
  g.bias = rows + n->L.cb2; g.sbz = P;
  g.C = c.a2; g.ldc = H; g.sCz = (int64_t)B * 64 * H;
  g.M = B * 64; g.N = H; g.K = 800; g.Z = Z; g.epi = EPI_BIAS_LEAKY;
  RBNN_TRY(gemm_simt(n, g, st));
  RBNN_TRY(pool2_logits(n, c.a2, z0, Z, B, logits, c.lpart, st));
  return 0;
}

static int conv_grad_simt(rbnn_net* n, int head, const float* x, const int32_t* labels, int B, int s0, int s1,
                          const float* pbar, float* out_sum, cudaStream_t st) {
  const int H = n->H, C = n->C;
  const int64_t P = n->L.P;
  const size_t per = conv_bytes_per_zb(n, true) * (size_t)B + 4096;
  const int zc = (int)std::max<size_t>(1, std::min<size_t>(n->ws_budget / per, (size_t)(s1 - s0)));
  RBNN_TRY(ws_reserve(n, per * zc));
  bool first = true;
  for (int z0 = s0; z0 < s1; z0 += zc) {
    const int Z = std::min(zc, s1 - z0);
    Arena ar(n);
    ConvBufs c{};
    conv_carve(n, ar, Z, B, true, c);
    RBNN_TRY(conv_forward_chunk(n, x, B, z0, Z, c, c.logits, st));
    RBNN_TRY(head_dlogits(n, head, c.logits, labels, pbar, Z, B, C, c.dlogits, st));
    const float* rows = n->bank + (int64_t)z0 * P;
    RBNN_TRY(pool2_bwd_fused(n, c.a2, c.dlogits, z0, Z, B, c.dz2, nullptr, nullptr, st));
    GemmArgs d{};
    d.b_kn = 1;
    d.A = c.dz2; d.lda = H; d.sAz = (int64_t)B * 64 * H;
    d.B = rows + n->L.cw2; d.ldb = 800; d.sBz = P;
    d.C = c.col; d.ldc = 800; d.sCz = (int64_t)B * 64 * 800;   // dcol reuses the im2col buffer
    d.M = B * 64; d.N = 800; d.K = H; d.Z = Z; d.epi = EPI_NONE;
    RBNN_TRY(gemm_simt(n, d, st));
    RBNN_TRY(col2im_conv2(n, c.col, c.p1, Z * B, c.g1, st));
    RBNN_TRY(conv1_bwd_sum(n, c.g1, c.idx1, n->bank, z0, Z, B, out_sum, first ? 0 : 1, st, c.partial, c.parts));
    first = false;
  }
  return 0;
}

static int conv_probs_simt(rbnn_net* n, const float* x, int B, int s0, int s1, float* out_sum, float* out_logits,
                           cudaStream_t st) {
  const int C = n->C;
  const size_t per = conv_bytes_per_zb(n, false) * (size_t)B + 4096;
  const int zc = (int)std::max<size_t>(1, std::min<size_t>(n->ws_budget / per, (size_t)(s1 - s0)));
  RBNN_TRY(ws_reserve(n, per * zc));
  for (int z0 = s0; z0 < s1; z0 += zc) {
    const int Z = std::min(zc, s1 - z0);
    Arena ar(n);
    ConvBufs c{};
    conv_carve(n, ar, Z, B, false, c);
    float* lg = out_logits ? out_logits : c.logits;
    RBNN_TRY(conv_forward_chunk(n, x, B, z0, Z, c, lg, st));
    if (out_sum) RBNN_TRY(head_probs_accumulate(n, lg, Z, B, C, out_sum, st));
  }
  return 0;
}

// rows of the batch one pass may take so that a single-sample chunk fits the workspace budget
static int batch_chunk(const rbnn_net* n, int B, bool grad) {
  size_t per_row;
  if (n->arch == RBNN_ARCH_CONV) per_row = conv_bytes_per_zb(n, grad);
  else per_row = (size_t)(n->H * 4) * (n->arch == RBNN_ARCH_FC2 ? 2 : 1) * (grad ? 2 : 1) + n->C * 8 + 64;
  const size_t rows = std::max<size_t>(1, n->ws_budget / per_row);
  return (int)std::min<size_t>(rows, (size_t)B);
}

static int check_rows(const rbnn_net* n, int s0, int s1) {
  RBNN_CHECK(n != nullptr, "null net handle");
  RBNN_CHECK(s0 >= 0 && s1 >= s0 && s1 <= n->capacity, "bank rows [%d,%d) outside the reserved capacity %d", s0, s1,
             n->capacity);
  return 0;
}

}  // namespace rbnn

using namespace rbnn;

extern "C" {

int rbnn_abi_version(void) { return 1; }

const char* rbnn_last_error(void) { return g_err.c_str(); }

int rbnn_net_create(rbnn_net** out, int arch, int in_ch, int in_h, int in_w, int hidden, int n_classes,
                    int device) {
  RBNN_CHECK(out != nullptr, "rbnn_net_create: out is NULL");
  *out = nullptr;
  RBNN_CHECK(arch == RBNN_ARCH_FC || arch == RBNN_ARCH_FC2 || arch == RBNN_ARCH_CONV,
             "architecture %d not implemented (model_nn.py:123-124)", arch);
  RBNN_CHECK(hidden >= 16 && (hidden & (hidden - 1)) == 0,
             "hidden size should be a power of 2 greater than 16 (model_nn.py:39-40), got %d", hidden);
  RBNN_CHECK(n_classes >= 2 && n_classes <= 32, "n_classes %d not in [2,32]", n_classes);
  RBNN_CHECK(in_ch >= 1 && in_h >= 1 && in_w >= 1, "bad input shape");
  if (arch == RBNN_ARCH_CONV)
    RBNN_CHECK(in_ch == 1 && in_h == 28 && in_w == 28,
               "conv architecture is defined for 1x28x28 inputs only (model_nn.py:95-106)");
  int ndev = 0;
  RBNN_CUDA(cudaGetDeviceCount(&ndev));
  RBNN_CHECK(device >= 0 && device < ndev, "CUDA device %d not present (%d devices)", device, ndev);
  rbnn_net* n = new rbnn_net();
  n->arch = arch; n->in_ch = in_ch; n->in_h = in_h; n->in_w = in_w;
  n->D = in_ch * in_h * in_w; n->H = hidden; n->C = n_classes; n->device = device;
  build_layout(n);
  if (const char* e = getenv("RBNN_TC_UNFUSED")) n->tc_unfused = atoi(e);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
    n->sm_count = prop.multiProcessorCount;
    n->cc_major = prop.major;
  }
  *out = n;
  return 0;
}

int rbnn_net_destroy(rbnn_net* n) {
  if (!n) return 0;
  DeviceGuard dg(n->device);
  cudaDeviceSynchronize();
  cudaFree(n->bank); cudaFree(n->woutp); cudaFree(n->sigma); cudaFree(n->ws);
  tc_bank_free(n);
  tc_keep_free(n);
  delete n;
  return 0;
}

int64_t rbnn_net_param_count(const rbnn_net* n) { return n ? n->L.P : -1; }

int rbnn_net_set_precision(rbnn_net* n, int prec) {
  RBNN_CHECK(n != nullptr, "null net handle");
  RBNN_CHECK(prec == RBNN_PREC_FP32 || prec == RBNN_PREC_TF32X3 || prec == RBNN_PREC_BF16 || prec == RBNN_PREC_F16X3,
             "unknown precision %d", prec);
  if (prec != RBNN_PREC_FP32 && n->arch == RBNN_ARCH_CONV)
    RBNN_CHECK((prec == RBNN_PREC_TF32X3 || prec == RBNN_PREC_F16X3) && tc_conv_supported(n),
               "arch conv: the tcgen05 engine offers TF32X3 and F16X3 (implicit-GEMM conv2) on sm_100 only");
  else if (prec != RBNN_PREC_FP32)
    RBNN_CHECK(tc_supported(n), "the tcgen05 engine covers arch fc/fc2 with D%%8==0 and H>=32 on sm_100 only");
  if (prec == RBNN_PREC_F16X3 && n->arch != RBNN_ARCH_CONV)
    RBNN_CHECK(tc_f16x3_supported(n), "F16X3 covers arch fc with hidden sizes the fused forward+head kernel supports");
  RBNN_CHECK(prec == RBNN_PREC_FP32 || prec == RBNN_PREC_TF32X3 || n->arch == RBNN_ARCH_CONV || (n->D & 7) == 0,
             "networks with D %% 8 != 0 (half moons) run on RBNN_PREC_FP32 or, arch fc2, RBNN_PREC_TF32X3");
  RBNN_CHECK(prec == RBNN_PREC_FP32 || n->act == RBNN_ACT_LEAKY,
             "the tensor-core engines fuse LeakyReLU: activation %d runs on RBNN_PREC_FP32 only", n->act);
  if (n->prec != prec) n->keep.valid = 0;
  n->prec = prec;
  return 0;
}

int rbnn_net_set_activation(rbnn_net* n, int act) {
  RBNN_CHECK(n != nullptr, "null net handle");
  RBNN_CHECK(act == RBNN_ACT_LEAKY || act == RBNN_ACT_RELU || act == RBNN_ACT_SIGM || act == RBNN_ACT_TANH,
             "unknown activation %d", act);
  RBNN_CHECK(act == RBNN_ACT_LEAKY || n->arch != RBNN_ARCH_CONV,
             "arch conv implements LeakyReLU only (its kernels fuse it with the pooling layers)");
  if (act != n->act) n->keep.valid = 0;
  n->act = act;
  if (act != RBNN_ACT_LEAKY) n->prec = RBNN_PREC_FP32;
  return 0;
}

int rbnn_net_get_precision(const rbnn_net* n) { return n ? n->prec : -1; }

int64_t rbnn_net_launch_count(const rbnn_net* n) { return n ? n->launches : -1; }

int rbnn_net_input_grid(rbnn_net* n) {
  if (!n) return -1;
  if (!n->tc.xgrid_last) return 0;
  unsigned v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpy(&v, n->tc.xgrid_last, sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v ? 1 : 0;
}

int rbnn_net_timing_enable(rbnn_net* n, int on) {
  RBNN_CHECK(n != nullptr, "null net handle");
  n->timing = on ? 1 : 0;
  return 0;
}

int rbnn_net_timing_read(rbnn_net* n, int cls, double* total_ms, int64_t* launches) {
  RBNN_CHECK(n != nullptr && cls >= 1 && cls <= 2, "timing class must be 1 (forward GEMM) or 2 (input-grad GEMM)");
  DeviceGuard dg(n->device);
  RBNN_CUDA(cudaDeviceSynchronize());
  double tot = 0.0;
  const size_t cnt = std::min(n->ev[cls][0].size(), n->ev[cls][1].size());
  for (size_t i = 0; i < cnt; ++i) {
    float ms = 0.f;
    RBNN_CUDA(cudaEventElapsedTime(&ms, n->ev[cls][0][i], n->ev[cls][1][i]));
    tot += ms;
  }
  for (int k = 0; k < 2; ++k) {
    for (cudaEvent_t e : n->ev[cls][k]) cudaEventDestroy(e);
    n->ev[cls][k].clear();
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int64_t)cnt;
  return 0;
}

int rbnn_bank_reserve(rbnn_net* n, int capacity) {
  RBNN_CHECK(n != nullptr, "null net handle");
  RBNN_CHECK(capacity >= 1, "bank capacity must be >= 1");
  if (capacity <= n->capacity) return 0;
  DeviceGuard dg(n->device);
  RBNN_CUDA(cudaDeviceSynchronize());
  float* nb = nullptr;
  RBNN_CUDA(cudaMalloc(&nb, (size_t)capacity * n->L.P * sizeof(float)));
  if (n->bank) {
    RBNN_CUDA(cudaMemcpy(nb, n->bank, (size_t)n->capacity * n->L.P * sizeof(float), cudaMemcpyDeviceToDevice));
    RBNN_CUDA(cudaFree(n->bank));
  }
  n->bank = nb;
  if (n->arch == RBNN_ARCH_CONV) {
    const size_t row = (size_t)conv_class_pitch(n->C) * 49 * n->H;
    float* nw = nullptr;
    RBNN_CUDA(cudaMalloc(&nw, (size_t)capacity * row * sizeof(float)));
    if (n->woutp) {
      RBNN_CUDA(cudaMemcpy(nw, n->woutp, (size_t)n->capacity * row * sizeof(float), cudaMemcpyDeviceToDevice));
      RBNN_CUDA(cudaFree(n->woutp));
    }
    n->woutp = nw;
  }
  n->capacity = capacity;
  n->alloc_epoch++;
  return 0;
}

int rbnn_bank_capacity(const rbnn_net* n) { return n ? n->capacity : -1; }

static void mark_dirty(rbnn_net* n, int s0, int count) {
  if (n->keep.valid && s0 < n->keep.s1 && s0 + count > n->keep.s0) n->keep.valid = 0;   // the kept forward used these rows
  if (n->tc.dirty)
    for (int s = s0; s < s0 + count && s < n->tc.capacity; ++s) n->tc.dirty[s] = 1;
}

int rbnn_bank_upload(rbnn_net* n, const float* weights, int s0, int count, int is_device, void* stream) {
  RBNN_TRY(check_rows(n, s0, s0 + count));
  if (count == 0) return 0;
  DeviceGuard dg(n->device);
  cudaStream_t st = (cudaStream_t)stream;
  RBNN_CUDA(cudaMemcpyAsync(n->bank + (int64_t)s0 * n->L.P, weights, (size_t)count * n->L.P * sizeof(float),
                            is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  RBNN_TRY(conv_permute_wout(n, s0, count, st));
  mark_dirty(n, s0, count);
  return 0;
}

int rbnn_bank_sample_diag(rbnn_net* n, const float* d_loc, const float* d_rho, uint64_t seed, int64_t sample_index0,
                          int64_t sample_index_stride, int s0, int count, void* stream) {
  return rbnn_bank_sample_diag_at(n, d_loc, d_rho, seed, sample_index0, sample_index_stride, s0, count, nullptr, stream);
}

int rbnn_bank_invalidate(rbnn_net* n) {
  RBNN_CHECK(n != nullptr, "null net handle");
  DeviceGuard dg(n->device);
  return tc_bank_invalidate(n);
}

int64_t rbnn_net_alloc_epoch(const rbnn_net* n) { return n ? n->alloc_epoch : -1; }

int rbnn_bank_sample_diag_at(rbnn_net* n, const float* d_loc, const float* d_rho, uint64_t seed, int64_t sample_index0,
                             int64_t sample_index_stride, int s0, int count, const int64_t* d_index_offset,
                             void* stream) {
  RBNN_TRY(check_rows(n, s0, s0 + count));
  RBNN_CHECK(sample_index_stride >= 1, "sample_index_stride must be >= 1");
  RBNN_CHECK(sample_index0 >= 0 && sample_index0 + (int64_t)count * sample_index_stride <= 0xFFFFFFFFLL,
             "sample index out of the 32-bit counter range");
  if (count == 0) return 0;
  DeviceGuard dg(n->device);
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int kMaxRowsPerLaunch = 32768;      // the sample index is a grid dimension (limit 65535)
  for (int c0 = 0; c0 < count; c0 += kMaxRowsPerLaunch) {
    const int cnt = std::min(kMaxRowsPerLaunch, count - c0);
    const int r0 = s0 + c0;
    const int64_t idx0 = sample_index0 + (int64_t)c0 * sample_index_stride;
    int fused = 0;      // F16X3 / arch fc with a fixed weight scale: bank rows and operand copies in one pass
    RBNN_TRY(tc_sample_relayout_f16(n, d_loc, d_rho, seed, idx0, sample_index_stride, r0, cnt, st, &fused, d_index_offset));
    if (fused) {
      if (n->keep.valid && r0 < n->keep.s1 && r0 + cnt > n->keep.s0) n->keep.valid = 0;
      continue;
    }
    RBNN_TRY(sample_diag(n, d_loc, d_rho, seed, idx0, sample_index_stride, r0, cnt, st, d_index_offset));
    RBNN_TRY(conv_permute_wout(n, r0, cnt, st));
    mark_dirty(n, r0, cnt);
  }
  return 0;
}

int rbnn_bank_download(rbnn_net* n, float* h_out, int s0, int count) {
  RBNN_TRY(check_rows(n, s0, s0 + count));
  DeviceGuard dg(n->device);
  RBNN_CUDA(cudaDeviceSynchronize());
  RBNN_CUDA(cudaMemcpy(h_out, n->bank + (int64_t)s0 * n->L.P, (size_t)count * n->L.P * sizeof(float),
                       cudaMemcpyDeviceToHost));
  return 0;
}

int rbnn_forward_probs_sum(rbnn_net* n, const float* d_x, int B, int s0, int s1, float* d_out_sum, void* stream) {
  RBNN_TRY(check_rows(n, s0, s1));
  RBNN_CHECK(B >= 0, "negative batch");
  if (B == 0) return 0;
  DeviceGuard dg(n->device);
  cudaStream_t st = (cudaStream_t)stream;
  RBNN_CUDA(cudaMemsetAsync(d_out_sum, 0, (size_t)B * n->C * sizeof(float), st));
  if (s1 == s0) return 0;
  if (n->prec != RBNN_PREC_FP32 && n->arch == RBNN_ARCH_CONV) return tc_conv_forward(n, d_x, B, s0, s1, d_out_sum, nullptr, st);
  if (n->prec != RBNN_PREC_FP32) return tc_fc_forward(n, d_x, B, s0, s1, d_out_sum, nullptr, st);
  const int bc = batch_chunk(n, B, false);
  for (int b0 = 0; b0 < B; b0 += bc) {
    const int nb = std::min(bc, B - b0);
    if (n->arch == RBNN_ARCH_CONV)
      RBNN_TRY(conv_probs_simt(n, d_x + (int64_t)b0 * n->D, nb, s0, s1, d_out_sum + (int64_t)b0 * n->C, nullptr, st));
    else
      RBNN_TRY(fc_probs_simt(n, d_x + (int64_t)b0 * n->D, nb, s0, s1, d_out_sum + (int64_t)b0 * n->C, nullptr, st));
  }
  return 0;
}

int rbnn_forward_logits_sum(rbnn_net* n, const float* d_x, int B, int s0, int s1, float* d_out_sum, void* stream) {
  RBNN_CHECK(n != nullptr, "null net handle");
  n->sum_logits = 1;                      // the per-chunk accumulation adds raw logits (head.cu)
  const int rc = rbnn_forward_probs_sum(n, d_x, B, s0, s1, d_out_sum, stream);
  n->sum_logits = 0;
  return rc;
}

int rbnn_forward_probs_sum_keep(rbnn_net* n, const float* d_x, int B, int s0, int s1, float* d_out_sum, void* stream) {
  RBNN_TRY(check_rows(n, s0, s1));
  RBNN_CHECK(B >= 0, "negative batch");
  n->keep.valid = 0;
  if (B == 0 || s1 == s0 || n->prec == RBNN_PREC_FP32)
    return rbnn_forward_probs_sum(n, d_x, B, s0, s1, d_out_sum, stream);
  DeviceGuard dg(n->device);
  cudaStream_t st = (cudaStream_t)stream;
  RBNN_CUDA(cudaMemsetAsync(d_out_sum, 0, (size_t)B * n->C * sizeof(float), st));
  if (n->arch == RBNN_ARCH_CONV) return tc_conv_forward_keep(n, d_x, B, s0, s1, d_out_sum, st);
  return tc_fc_forward_keep(n, d_x, B, s0, s1, d_out_sum, st);
}

int rbnn_keep_valid(const rbnn_net* n) { return n ? n->keep.valid : 0; }

int rbnn_input_grad_sum_kept(rbnn_net* n, int head, const int32_t* d_labels, const float* d_pbar, float* d_out_sum,
                             void* stream) {
  RBNN_CHECK(n != nullptr, "null net handle");
  RBNN_CHECK(head == RBNN_HEAD_MEAN_OF_GRADS || head == RBNN_HEAD_GRAD_OF_MEAN || head == RBNN_HEAD_UPSTREAM ||
                 head == RBNN_HEAD_LOGITS_UPSTREAM, "head %d has no kept route", head);
  RBNN_CHECK(head == RBNN_HEAD_MEAN_OF_GRADS || d_pbar != nullptr, "GRAD_OF_MEAN / UPSTREAM / LOGITS_UPSTREAM need d_pbar");
  DeviceGuard dg(n->device);
  if (n->arch == RBNN_ARCH_CONV) return tc_conv_grad_kept(n, head, d_labels, d_pbar, d_out_sum, (cudaStream_t)stream);
  return tc_fc_grad_kept(n, head, d_labels, d_pbar, d_out_sum, (cudaStream_t)stream);
}

int rbnn_forward_logits(rbnn_net* n, const float* d_x, int B, int s, float* d_out, void* stream) {
  RBNN_TRY(check_rows(n, s, s + 1));
  if (B <= 0) return 0;
  DeviceGuard dg(n->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (n->prec != RBNN_PREC_FP32 && n->arch == RBNN_ARCH_CONV) return tc_conv_forward(n, d_x, B, s, s + 1, nullptr, d_out, st);
  if (n->prec != RBNN_PREC_FP32) return tc_fc_forward(n, d_x, B, s, s + 1, nullptr, d_out, st);
  const int bc = batch_chunk(n, B, false);
  for (int b0 = 0; b0 < B; b0 += bc) {
    const int nb = std::min(bc, B - b0);
    if (n->arch == RBNN_ARCH_CONV)
      RBNN_TRY(conv_probs_simt(n, d_x + (int64_t)b0 * n->D, nb, s, s + 1, nullptr, d_out + (int64_t)b0 * n->C, st));
    else
      RBNN_TRY(fc_probs_simt(n, d_x + (int64_t)b0 * n->D, nb, s, s + 1, nullptr, d_out + (int64_t)b0 * n->C, st));
  }
  return 0;
}

int rbnn_input_grad_sum(rbnn_net* n, int head, const float* d_x, const int32_t* d_labels, int B, int s0, int s1,
                        const float* d_pbar, float* d_out_sum, void* stream) {
  RBNN_TRY(check_rows(n, s0, s1));
  RBNN_CHECK(head >= RBNN_HEAD_MEAN_OF_GRADS && head <= RBNN_HEAD_LOGITS_UPSTREAM, "unknown head %d", head);
  RBNN_CHECK((head != RBNN_HEAD_GRAD_OF_MEAN && head != RBNN_HEAD_UPSTREAM && head != RBNN_HEAD_LOGITS_UPSTREAM) ||
                 d_pbar != nullptr, "GRAD_OF_MEAN / UPSTREAM / LOGITS_UPSTREAM need d_pbar");
  RBNN_CHECK(head != RBNN_HEAD_LOGITS_CE || s1 - s0 == 1, "LOGITS_CE takes exactly one bank row");
  if (B <= 0) return 0;
  DeviceGuard dg(n->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (s1 == s0) {
    RBNN_CUDA(cudaMemsetAsync(d_out_sum, 0, (size_t)B * n->D * sizeof(float), st));
    return 0;
  }
  if (n->prec != RBNN_PREC_FP32 && n->arch == RBNN_ARCH_CONV)
    return tc_conv_input_grad_sum(n, head, d_x, d_labels, B, s0, s1, d_pbar, d_out_sum, st);
  if (n->prec != RBNN_PREC_FP32) return tc_fc_input_grad_sum(n, head, d_x, d_labels, B, s0, s1, d_pbar, d_out_sum, st);
  const int bc = batch_chunk(n, B, true);
  for (int b0 = 0; b0 < B; b0 += bc) {
    const int nb = std::min(bc, B - b0);
    const float* pb = d_pbar ? d_pbar + (int64_t)b0 * n->C : nullptr;
    if (n->arch == RBNN_ARCH_CONV)
      RBNN_TRY(conv_grad_simt(n, head, d_x + (int64_t)b0 * n->D, d_labels + b0, nb, s0, s1, pb,
                              d_out_sum + (int64_t)b0 * n->D, st));
    else
      RBNN_TRY(fc_grad_simt(n, head, d_x + (int64_t)b0 * n->D, d_labels + b0, nb, s0, s1, pb,
                            d_out_sum + (int64_t)b0 * n->D, st));
  }
  return 0;
}

int rbnn_loss_gradients_host(rbnn_net* n, const float* h_x, const int32_t* h_labels, int B, int s0, int s1,
                             int n_samples_global, float* h_out) {
  RBNN_TRY(check_rows(n, s0, s1));
  RBNN_CHECK(n_samples_global >= 1, "n_samples_global must be >= 1");
  if (B <= 0) return 0;
  DeviceGuard dg(n->device);
  float *dx = nullptr, *dout = nullptr;
  int32_t* dy = nullptr;
  const size_t xb = (size_t)B * n->D * sizeof(float);
  RBNN_CUDA(cudaMalloc(&dx, xb));
  RBNN_CUDA(cudaMalloc(&dout, xb));
  RBNN_CUDA(cudaMalloc(&dy, (size_t)B * sizeof(int32_t)));
  cudaStream_t st = nullptr;
  int rc = 0;
  do {
    if (cudaMemcpyAsync(dx, h_x, xb, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(dy, h_labels, (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, st) != cudaSuccess) {
      set_error("loss_gradients_host: H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = 1;
      break;
    }
    rc = rbnn_input_grad_sum(n, RBNN_HEAD_MEAN_OF_GRADS, dx, dy, B, s0, s1, nullptr, dout, st);
    if (rc) break;
    rc = scale_inplace(n, dout, 1.f / (float)n_samples_global, (int64_t)B * n->D, st);
    if (rc) break;
    if (cudaMemcpyAsync(h_out, dout, xb, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      set_error("loss_gradients_host: D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = 1;
    }
  } while (0);
  cudaFree(dx); cudaFree(dout); cudaFree(dy);
  return rc;
}

}  // extern "C"
