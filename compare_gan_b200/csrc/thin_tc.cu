// Image-side ("thin") convolutions on the tensor cores — math_mode 1.
//
// The first convolution of every discriminator (3 input channels), the last of every generator (3 output channels) and
// Inception's stem are contractions over kh*kw*3 = 27 values per pixel: far too skinny for an implicit GEMM over
// 32-channel k-blocks (the tap loop would move the 256-channel operand nine times for 3 output channels) and, as fp32
// streaming kernels (thin.cu), bound by the FP32 pipe at 4-6x their HBM time (profiles/r2_launch_summary: 10 % of the
// resnet_cifar10 cycle).  Here each becomes ONE dense 32-wide GEMM on the existing tcgen05 kernels plus a streaming pass:
//
//   cin <= 4   forward   : P = patches(x) [pixels, 32]  ->  y  = P W            (1x1 tcgen05 conv, fused epilogue)
//              filter    : dW = P^T dy                                            (tcgen05 filter-gradient kernel)
//              input grad: T = dy W^T [pixels, 32]      ->  dx = shift_add(T)
//   cout <= 4  forward   : T = x W' [pixels, 32]        ->  y  = shift_add(T) + bias
//              input grad: P = patches(dy)              ->  dx = P W'^T
//              filter    : dW' = x^T P                                            (then re-laid out to HWIO)
//
// patches() gathers the kh*kw*C <= 32 values a pixel's taps see into one 128-byte row (zero where SAME padding applies,
// TF32-rounded: the GEMM skips its operand rounding); shift_add() is its adjoint: out[p, c] = sum_tap T[p + off(tap), tap*C + c].
// Both stream a [pixels, 32] fp32 tensor once (1/4 of a 128-channel activation).  Arithmetic: TF32 operands, fp32
// accumulation — the same as every other tensor-core contraction of math_mode 1 (CGAN_PATH_TCGEN05_TF32).
#include "common.cuh"

int cgan_conv_tc(cgan_ctx* ctx, const float* in, int nviews, const long long* view_off, long long in_sw, long long in_sh,
                 long long in_sn, int n, int h, int w, int gh, int gw, int kdim, const float* wsrc, int taps_total,
                 int transpose_w, int ncols, int ntaps, const int* off_h, const int* off_w, const int* wtap, const int* amap,
                 const float* bias, float* out, long long s_n, long long s_h, long long s_w, long long base, int relu,
                 const int* view_phase_of, int wimg_stride, const TcExtra* ex);
bool cgan_tc_shape_ok(int n, int h, int w, int kdim, int ncols);
bool cgan_wgrad_tc_geometry_ok(int n, int h, int w);
int cgan_wgrad_tc(cgan_ctx* ctx, const cgan_conv_desc* d, const float* x, const float* dy, float* dw, int x_tf32, int dy_tf32);

namespace {

constexpr int TT_K = 32;                        // patch row: 32 floats = one 128-byte TMA / swizzle row
constexpr size_t TT_RESERVE = 16u << 20;        // head of the workspace left to the GEMM kernels (weight prep, split-K partials)
constexpr int TT_MAX_TAPS = 32;

struct TapList {
  int ntaps, c, k;                              // k = ntaps * c <= 32
  int off_h[TT_MAX_TAPS], off_w[TT_MAX_TAPS];
};

__device__ __forceinline__ float tt_rna(float v) {
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// out[(n, gy, gx)][m], m = tap*C + c  =  src[n, gy*stride + off_h[tap], gx*stride + off_w[tap], c]  (0 outside / m >= K).
// One thread per (pixel, 4 consecutive m): eight threads write one 128-byte row.  A thread keeps its quarter-row index for
// the whole grid-stride loop (the stride is a multiple of 8), so the (tap, channel) decode of its four values happens once;
// per pixel it costs one 32-bit decode, four predicated loads from the (L1 / L2 resident) 3-channel source and one STG.128.
__global__ void __launch_bounds__(256)
patch_kernel(float* __restrict__ out, const float* __restrict__ src, TapList t, int n, int gh, int gw, int sh, int sw, int stride) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned q = tid & 7u;
  int dy[4], dx[4], cc[4];
  bool live[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int m = (int)q * 4 + j;
    live[j] = m < t.k;
    const int tap = live[j] ? m / t.c : 0;
    cc[j] = live[j] ? m - tap * t.c : 0;
    dy[j] = t.off_h[tap]; dx[j] = t.off_w[tap];
  }
  const unsigned npix = (unsigned)n * gh * gw;                // < 2^31 (host-checked)
  const unsigned pstep = (gridDim.x * blockDim.x) >> 3;
  for (unsigned pix = tid >> 3; pix < npix; pix += pstep) {
    const unsigned row = pix / (unsigned)gw;
    const int gx = (int)(pix - row * (unsigned)gw);
    const unsigned img = row / (unsigned)gh;
    const int gy = (int)(row - img * (unsigned)gh);
    const float* base = src + (size_t)img * sh * sw * t.c;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int y = gy * stride + dy[j], x = gx * stride + dx[j];
      const bool ok = live[j] && y >= 0 && y < sh && x >= 0 && x < sw;
      v[j] = ok ? tt_rna(__ldg(base + ((size_t)y * sw + x) * t.c + cc[j])) : 0.f;
    }
    *reinterpret_cast<float4*>(out + (size_t)pix * TT_K + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// out[(n, y, x)][c] = bias[c] + sum_tap T[(n, y + off_h[tap], x + off_w[tap])][tap*C + c]   (taps outside the th x tw grid: 0)
__global__ void shift_add_kernel(float* __restrict__ out, const float* __restrict__ t32, const float* __restrict__ bias,
                                 TapList t, int n, int oh, int ow, int th, int tw, int relu) {
  const long long total = (long long)n * oh * ow;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long pix = i;
    const int x = (int)(pix % ow);
    pix /= ow;
    const int y = (int)(pix % oh);
    const int img = (int)(pix / oh);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias)
      for (int c = 0; c < t.c; ++c) acc[c] = bias[c];
    for (int tap = 0; tap < t.ntaps; ++tap) {
      const int ty = y + t.off_h[tap], tx = x + t.off_w[tap];
      if (ty < 0 || ty >= th || tx < 0 || tx >= tw) continue;
      const float* row = t32 + (((long long)img * th + ty) * tw + tx) * TT_K + tap * t.c;
      for (int c = 0; c < t.c; ++c) acc[c] += __ldg(row + c);
    }
    for (int c = 0; c < t.c; ++c) out[i * t.c + c] = relu ? fmaxf(acc[c], 0.f) : acc[c];
  }
}

// HWIO filter w[tap][ci][co] (co < C <= 4)  <->  W'[ci][ld] with W'[ci][tap*C + co]; columns >= ntaps*C are zero
__global__ void wcols_from_hwio_kernel(float* __restrict__ wp, const float* __restrict__ w, int ntaps, int cin, int c, int ld) {
  const int total = cin * ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i / ld, m = i - ci * ld;
    float v = 0.f;
    if (m < ntaps * c) {
      const int tap = m / c, co = m - tap * c;
      v = w[((long long)tap * cin + ci) * c + co];
    }
    wp[i] = v;
  }
}
// dst[rows_dst][cols] = src[rows_src][cols] followed by zero rows
__global__ void pad_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows_src, int rows_dst, int cols) {
  const int total = rows_dst * cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    dst[i] = (i / cols) < rows_src ? src[i] : 0.f;
}
__global__ void hwio_from_wcols_kernel(float* __restrict__ w, const float* __restrict__ wp, int ntaps, int cin, int c, int ld) {
  const int total = ntaps * cin * c;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % c, ci = (i / c) % cin, tap = i / (c * cin);
    w[i] = wp[(long long)ci * ld + tap * c + co];
  }
}

inline int ew_blocks(cgan_ctx* ctx, long long n) {
  long long b = (n + 255) / 256, cap = (long long)ctx->num_sms * 32;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// forward taps of the convolution d: tap (kh, kw) reads the input at output*stride + (kh - pad_t, kw - pad_l)
inline void taps_forward(const cgan_conv_desc* d, int c, TapList* t) {
  t->ntaps = d->kh * d->kw; t->c = c; t->k = t->ntaps * c;
  for (int kh = 0; kh < d->kh; ++kh)
    for (int kw = 0; kw < d->kw; ++kw) { t->off_h[kh * d->kw + kw] = kh - d->pad_t; t->off_w[kh * d->kw + kw] = kw - d->pad_l; }
}
// adjoint taps (stride 1): input pixel ih receives tap kh from output row ih + pad_t - kh
inline void taps_adjoint(const cgan_conv_desc* d, int c, TapList* t) {
  t->ntaps = d->kh * d->kw; t->c = c; t->k = t->ntaps * c;
  for (int kh = 0; kh < d->kh; ++kh)
    for (int kw = 0; kw < d->kw; ++kw) { t->off_h[kh * d->kw + kw] = d->pad_t - kh; t->off_w[kh * d->kw + kw] = d->pad_l - kw; }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// workspace: [0, TT_RESERVE) for the GEMM kernels' own use | patch / T tensor | small weight buffers (2 x 64 KB)
inline int tt_workspace(cgan_ctx* ctx, long long pixels, float** big, float** small0, float** small1) {
  const size_t big_bytes = ((size_t)pixels * TT_K * sizeof(float) + 255) / 256 * 256;
  void* ws = nullptr;
  int rc = cgan_ws(ctx, TT_RESERVE + big_bytes + (128u << 10), &ws);
  if (rc) return rc;
  char* b = reinterpret_cast<char*>(ws) + TT_RESERVE;
  *big = reinterpret_cast<float*>(b);
  *small0 = reinterpret_cast<float*>(b + big_bytes);
  *small1 = reinterpret_cast<float*>(b + big_bytes + (64u << 10));
  return CGAN_OK;
}

inline TcExtra extra_from(const cgan_conv_epilogue* ep, int a_prerounded) {
  TcExtra ex;
  memset(&ex, 0, sizeof(ex));
  ex.a_prerounded = a_prerounded;
  if (ep) {
    ex.round_out = (ep->flags & CGAN_CONV_ROUND_OUT) ? 1 : 0;
    ex.residual = ep->residual; ex.mask = ep->mask; ex.mask_leak = ep->mask_leak;
  }
  return ex;
}

inline bool common_ok(cgan_ctx* ctx, const cgan_conv_desc* d) {
  return ctx->math_mode == 1 && !d->upsample && d->kh * d->kw <= TT_MAX_TAPS && d->n >= 1 &&
         (long long)d->n * d->oh * d->ow < (1ll << 31) && (long long)d->n * d->h * d->w < (1ll << 31);
}
inline bool stride1_same_grid(const cgan_conv_desc* d) { return d->stride == 1 && d->oh == d->h && d->ow == d->w; }

}  // namespace

// ---- cin <= 4 -------------------------------------------------------------------------------------------------------------
bool cgan_thin_tc_cin_ok(cgan_ctx* ctx, const cgan_conv_desc* d) {
  return common_ok(ctx, d) && d->cin >= 1 && d->cin <= 4 && d->kh * d->kw * d->cin <= TT_K && d->cout >= 16 && d->cout % 4 == 0 &&
         d->cout <= 512 && cgan_tc_shape_ok(d->n, d->oh, d->ow, TT_K, d->cout) &&
         (size_t)2 * ctx->num_sms * TT_K * d->cout * 4 <= TT_RESERVE;
}

int cgan_thin_tc_fwd_cin(cgan_ctx* ctx, const cgan_conv_desc* d, const float* x, const float* w, const cgan_conv_epilogue* ep,
                         float* y) {
  const long long pixels = (long long)d->n * d->oh * d->ow;
  float *pt, *s0, *s1;
  int rc = tt_workspace(ctx, pixels, &pt, &s0, &s1);
  if (rc) return rc;
  TapList t;
  taps_forward(d, d->cin, &t);
  patch_kernel<<<ew_blocks(ctx, pixels * 8), 256, 0, ctx->stream>>>(pt, x, t, d->n, d->oh, d->ow, d->h, d->w, d->stride);
  CGAN_LAUNCHED(ctx);
  // HWIO flattened is [K = kh*kw*cin][cout]: padded to 32 rows so that the GEMM sees plain 32-channel operands
  pad_rows_kernel<<<cdiv((long long)TT_K * d->cout, 256), 256, 0, ctx->stream>>>(s0, w, t.k, TT_K, d->cout);
  CGAN_LAUNCHED(ctx);
  const int ldy = (ep && ep->ldy) ? ep->ldy : d->cout;
  const long long zero = 0;
  const int o0 = 0;
  TcExtra ex = extra_from(ep, 1);
  // y = P W: a 1x1 convolution over the 32-channel patch tensor
  return cgan_conv_tc(ctx, pt, 1, &zero, TT_K, (long long)d->ow * TT_K, (long long)d->oh * d->ow * TT_K, d->n, d->oh, d->ow, d->oh,
                      d->ow, TT_K, s0, 1, 1, d->cout, 1, &o0, &o0, &o0, nullptr, ep ? ep->bias : nullptr, y,
                      (long long)d->oh * d->ow * ldy, (long long)d->ow * ldy, ldy, 0, (ep && (ep->flags & CGAN_CONV_RELU)) ? 1 : 0,
                      nullptr, 0, &ex);
}

bool cgan_thin_tc_wgrad_cin_ok(cgan_ctx* ctx, const cgan_conv_desc* d) {
  return cgan_thin_tc_cin_ok(ctx, d) && d->cout <= 256 && cgan_wgrad_tc_geometry_ok(d->n, d->oh, d->ow);
}

int cgan_thin_tc_wgrad_cin(cgan_ctx* ctx, const cgan_conv_desc* d, const float* x, const float* dy, int dy_tf32, float* dw) {
  const long long pixels = (long long)d->n * d->oh * d->ow;
  float *pt, *s0, *s1;
  int rc = tt_workspace(ctx, pixels, &pt, &s0, &s1);
  if (rc) return rc;
  TapList t;
  taps_forward(d, d->cin, &t);
  patch_kernel<<<ew_blocks(ctx, pixels * 8), 256, 0, ctx->stream>>>(pt, x, t, d->n, d->oh, d->ow, d->h, d->w, d->stride);
  CGAN_LAUNCHED(ctx);
  cgan_conv_desc g;          // dW32[32][cout] = P^T dy: the filter gradient of a 1x1 convolution 32 -> cout on the output grid
  memset(&g, 0, sizeof(g));
  g.n = d->n; g.h = g.oh = d->oh; g.w = g.ow = d->ow; g.cin = TT_K; g.cout = d->cout; g.kh = g.kw = 1; g.stride = 1;
  rc = cgan_wgrad_tc(ctx, &g, pt, dy, s0, 1, dy_tf32);
  if (rc) return rc;
  return cgan_copy(ctx, dw, s0, (int64_t)t.k * d->cout);          // rows [0, K) of dW32 are HWIO's [kh][kw][cin][cout]
}

bool cgan_thin_tc_dgrad_cin_ok(cgan_ctx* ctx, const cgan_conv_desc* d) {
  return common_ok(ctx, d) && stride1_same_grid(d) && d->cin >= 1 && d->cin <= 4 && d->kh * d->kw * d->cin <= TT_K && d->cout >= 8 &&
         d->cout % 4 == 0 && cgan_tc_shape_ok(d->n, d->oh, d->ow, d->cout, TT_K);
}

int cgan_thin_tc_dgrad_cin(cgan_ctx* ctx, const cgan_conv_desc* d, const float* dy, const float* w, const cgan_conv_epilogue* ep,
                           float* dx) {
  const long long pixels = (long long)d->n * d->oh * d->ow;
  float *tt, *s0, *s1;
  int rc = tt_workspace(ctx, pixels, &tt, &s0, &s1);
  if (rc) return rc;
  TapList t;
  taps_adjoint(d, d->cin, &t);
  const long long zero = 0;
  const int o0 = 0;
  TcExtra ex = extra_from(nullptr, (ep && (ep->flags & CGAN_CONV_IN_TF32)) ? 1 : 0);
  // T[p][tap*cin + ci] = sum_co dy[p][co] w[tap][ci][co]: HWIO flattened is [ncols = K][kdim = cout]
  rc = cgan_conv_tc(ctx, dy, 1, &zero, d->cout, (long long)d->ow * d->cout, (long long)d->oh * d->ow * d->cout, d->n, d->oh, d->ow,
                    d->oh, d->ow, d->cout, w, 1, 0, t.k, 1, &o0, &o0, &o0, nullptr, nullptr, tt, (long long)d->oh * d->ow * TT_K,
                    (long long)d->ow * TT_K, TT_K, 0, 0, nullptr, 0, &ex);
  if (rc) return rc;
  shift_add_kernel<<<ew_blocks(ctx, (long long)d->n * d->h * d->w), 256, 0, ctx->stream>>>(dx, tt, ep ? ep->bias : nullptr, t, d->n,
                                                                                           d->h, d->w, d->oh, d->ow, 0);
  CGAN_LAUNCHED(ctx);
  if (ep && (ep->residual || ep->mask || (ep->flags & (CGAN_CONV_RELU | CGAN_CONV_ROUND_OUT))))
    return cgan_conv_post_epilogue(ctx, dx, (int64_t)d->n * d->h * d->w, d->cin, d->cin, ep->residual, ep->mask, ep->mask_leak,
                                   (ep->flags & CGAN_CONV_RELU) ? 1 : 0, (ep->flags & CGAN_CONV_ROUND_OUT) ? 1 : 0);
  return CGAN_OK;
}

// ---- cout <= 4 ------------------------------------------------------------------------------------------------------------
bool cgan_thin_tc_cout_ok(cgan_ctx* ctx, const cgan_conv_desc* d) {
  return common_ok(ctx, d) && stride1_same_grid(d) && d->cout >= 1 && d->cout <= 4 && d->kh * d->kw * d->cout <= TT_K &&
         d->cin >= 32 && d->cin % 4 == 0 && d->cin <= 512 && cgan_tc_shape_ok(d->n, d->h, d->w, d->cin, TT_K);
}

int cgan_thin_tc_fwd_cout(cgan_ctx* ctx, const cgan_conv_desc* d, const float* x, const float* w, const cgan_conv_epilogue* ep,
                          float* y) {
  const long long pixels = (long long)d->n * d->h * d->w;
  float *tt, *wp, *s1;
  int rc = tt_workspace(ctx, pixels, &tt, &wp, &s1);
  if (rc) return rc;
  TapList t;
  taps_forward(d, d->cout, &t);
  wcols_from_hwio_kernel<<<cdiv((long long)d->cin * TT_K, 256), 256, 0, ctx->stream>>>(wp, w, t.ntaps, d->cin, d->cout, TT_K);
  CGAN_LAUNCHED(ctx);
  const long long zero = 0;
  const int o0 = 0;
  TcExtra ex = extra_from(nullptr, (ep && (ep->flags & CGAN_CONV_IN_TF32)) ? 1 : 0);
  // T[p][tap*cout + co] = sum_ci x[p][ci] w[tap][ci][co]: every tap's contribution at the INPUT pixel, all taps in one GEMM
  rc = cgan_conv_tc(ctx, x, 1, &zero, d->cin, (long long)d->w * d->cin, (long long)d->h * d->w * d->cin, d->n, d->h, d->w, d->h, d->w,
                    d->cin, wp, 1, 1, TT_K, 1, &o0, &o0, &o0, nullptr, nullptr, tt, (long long)d->h * d->w * TT_K, (long long)d->w * TT_K,
                    TT_K, 0, 0, nullptr, 0, &ex);
  if (rc) return rc;
  const bool post = ep && (ep->residual || ep->mask || (ep->flags & CGAN_CONV_ROUND_OUT));
  const int relu = (ep && (ep->flags & CGAN_CONV_RELU)) ? 1 : 0;
  shift_add_kernel<<<ew_blocks(ctx, pixels), 256, 0, ctx->stream>>>(y, tt, ep ? ep->bias : nullptr, t, d->n, d->oh, d->ow, d->h, d->w,
                                                                   post ? 0 : relu);
  CGAN_LAUNCHED(ctx);
  if (post)
    return cgan_conv_post_epilogue(ctx, y, (int64_t)pixels, d->cout, d->cout, ep->residual, ep->mask, ep->mask_leak, relu,
                                   (ep->flags & CGAN_CONV_ROUND_OUT) ? 1 : 0);
  return CGAN_OK;
}

int cgan_thin_tc_dgrad_cout(cgan_ctx* ctx, const cgan_conv_desc* d, const float* dy, const float* w, const cgan_conv_epilogue* ep,
                            float* dx) {
  const long long pixels = (long long)d->n * d->h * d->w;
  float *pt, *wp, *s1;
  int rc = tt_workspace(ctx, pixels, &pt, &wp, &s1);
  if (rc) return rc;
  TapList t;
  taps_adjoint(d, d->cout, &t);
  patch_kernel<<<ew_blocks(ctx, pixels * 8), 256, 0, ctx->stream>>>(pt, dy, t, d->n, d->h, d->w, d->oh, d->ow, 1);
  CGAN_LAUNCHED(ctx);
  wcols_from_hwio_kernel<<<cdiv((long long)d->cin * TT_K, 256), 256, 0, ctx->stream>>>(wp, w, t.ntaps, d->cin, d->cout, TT_K);
  CGAN_LAUNCHED(ctx);
  const long long zero = 0;
  const int o0 = 0;
  TcExtra ex = extra_from(ep, 1);
  // dx[p][ci] = sum_m P[p][m] W'[ci][m]: W' [ncols = cin][kdim = 32], columns >= K zero
  return cgan_conv_tc(ctx, pt, 1, &zero, TT_K, (long long)d->w * TT_K, (long long)d->h * d->w * TT_K, d->n, d->h, d->w, d->h, d->w, TT_K,
                      wp, 1, 0, d->cin, 1, &o0, &o0, &o0, nullptr, ep ? ep->bias : nullptr, dx, (long long)d->h * d->w * d->cin,
                      (long long)d->w * d->cin, d->cin, 0, (ep && (ep->flags & CGAN_CONV_RELU)) ? 1 : 0, nullptr, 0, &ex);
}

bool cgan_thin_tc_wgrad_cout_ok(cgan_ctx* ctx, const cgan_conv_desc* d) {
  return cgan_thin_tc_cout_ok(ctx, d) && d->cin % 32 == 0 && d->cin >= 64 && cgan_wgrad_tc_geometry_ok(d->n, d->h, d->w) &&
         (size_t)2 * ctx->num_sms * TT_K * d->cin * 4 <= TT_RESERVE && (size_t)d->cin * TT_K * 4 <= (64u << 10);
}

int cgan_thin_tc_wgrad_cout(cgan_ctx* ctx, const cgan_conv_desc* d, const float* x, const float* dy, int x_tf32, float* dw) {
  const long long pixels = (long long)d->n * d->h * d->w;
  float *pt, *s0, *dwp;
  int rc = tt_workspace(ctx, pixels, &pt, &s0, &dwp);
  if (rc) return rc;
  TapList t;
  taps_adjoint(d, d->cout, &t);
  patch_kernel<<<ew_blocks(ctx, pixels * 8), 256, 0, ctx->stream>>>(pt, dy, t, d->n, d->h, d->w, d->oh, d->ow, 1);
  CGAN_LAUNCHED(ctx);
  cgan_conv_desc g;          // dW'[cin][32] = x^T P
  memset(&g, 0, sizeof(g));
  g.n = d->n; g.h = g.oh = d->h; g.w = g.ow = d->w; g.cin = d->cin; g.cout = TT_K; g.kh = g.kw = 1; g.stride = 1;
  rc = cgan_wgrad_tc(ctx, &g, x, pt, dwp, x_tf32, 1);
  if (rc) return rc;
  hwio_from_wcols_kernel<<<cdiv((long long)t.k * d->cin, 256), 256, 0, ctx->stream>>>(dw, dwp, t.ntaps, d->cin, d->cout, TT_K);
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}
