// Batch-norm family (moments / apply / backward, conditional and cross-replica aware) and the
// spectral-norm power iteration.  All HBM-bound rows x channels work over NHWC activations:
// channels are the fastest dimension, so a warp reads 128 contiguous bytes of one row and the
// per-channel sums are reduced over row lanes in shared memory, then over chunks in a second,
// fixed-order (deterministic) pass.
//
// Replaces: standardize_batch (arch_ops.py:194-319), batch_norm (:327-367), conditional_batch_norm
// (:423-445), cross_replica_moments (tpu/tpu_ops.py:94-125), spectral_norm (arch_ops.py:453-535).
#include "common.cuh"

namespace {

constexpr int CR_X = 32, CR_Y = 8;   // column-reduce block: 32 channels x 8 row lanes

struct MomentsF {
  const float* x;
  __device__ __forceinline__ void operator()(long long r, int c, int C, long long, float* v) const {
    float t = x[r * C + c];
    v[0] = t;
    v[1] = t * t;
  }
};
struct SumF {
  const float* x;
  __device__ __forceinline__ void operator()(long long r, int c, int C, long long, float* v) const {
    v[0] = x[r * C + c];
  }
};
struct WeightedSumF {   // sum_r a[r] * w[r,c]  (W^T a)
  const float* w;
  const float* a;
  __device__ __forceinline__ void operator()(long long r, int c, int C, long long, float* v) const {
    v[0] = w[r * C + c] * a[r];
  }
};
struct BnBwdF {         // v0 = sum dy*xhat, v1 = sum dy
  const float* dy;
  const float* x;
  const float* mean_var;
  float eps;
  __device__ __forceinline__ void operator()(long long r, int c, int C, long long, float* v) const {
    float inv = 1.0f / sqrtf(mean_var[C + c] + eps);
    float xh = (x[r * C + c] - mean_var[c]) * inv;
    float g = dy[r * C + c];
    v[0] = g * xh;
    v[1] = g;
  }
};

// partial[((g*chunks + chunk)*NV + v)*C + c].  With `counters` the LAST chunk-block of each (column block, group) to finish
// also does the second stage — the fixed-order sum over the chunk partials that colreduce_final_kernel otherwise does in a
// launch of its own (same order, same result): every block publishes its partials (threadfence), takes a ticket, and the
// holder of ticket chunks-1 reads them all back.  The counter is reset by that block, so launches (and graph replays)
// always start from zero.
template <class F, int NV>
__global__ void colreduce_kernel(F f, float* __restrict__ partial, long long rows_per_group, int C,
                                 long long rows_per_chunk, int chunks, unsigned* counters, float scale, float* out0, float* out1) {
  __shared__ float sh[NV][CR_Y][CR_X + 1];
  __shared__ unsigned s_ticket;
  const int c = blockIdx.x * CR_X + threadIdx.x;
  const int chunk = blockIdx.y, g = blockIdx.z;
  float acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.f;
  if (c < C) {
    long long r0 = (long long)chunk * rows_per_chunk;
    long long r1 = min(rows_per_group, r0 + rows_per_chunk);
    long long base = (long long)g * rows_per_group;
    for (long long r = r0 + threadIdx.y; r < r1; r += CR_Y) {
      float v[NV];
      f(base + r, c, C, r, v);
#pragma unroll
      for (int k = 0; k < NV; ++k) acc[k] += v[k];
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) sh[v][threadIdx.y][threadIdx.x] = acc[v];
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < CR_Y; ++y) s += sh[v][y][threadIdx.x];
      partial[(((long long)g * chunks + chunk) * NV + v) * C + c] = s;
    }
  }
  if (!counters) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) s_ticket = atomicAdd(&counters[(size_t)g * gridDim.x + blockIdx.x], 1u);
  __syncthreads();
  if (s_ticket != (unsigned)(chunks - 1)) return;
  __threadfence();
  float* outs[2] = {out0, out1};
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float s = 0.f;
    if (c < C && outs[v])
      for (int k = threadIdx.y; k < chunks; k += CR_Y) s += __ldcg(&partial[(((long long)g * chunks + k) * NV + v) * C + c]);
    sh[v][threadIdx.y][threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!outs[v]) continue;
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < CR_Y; ++y) s += sh[v][y][threadIdx.x];
      outs[v][(long long)g * C + c] = s * scale;
    }
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) counters[(size_t)g * gridDim.x + blockIdx.x] = 0u;
}

// out_v[g*C + c] = scale * sum_chunk partial.  Block = 32 columns x 8 chunk lanes: the (up to ~150) partials of a column
// are summed by 8 threads in parallel and combined in a fixed order, instead of one long chain of dependent loads.
template <int NV>
__global__ void colreduce_final_kernel(const float* __restrict__ partial, int groups, int chunks, int C, float scale,
                                       float* out0, float* out1) {
  __shared__ float red[NV][8][32];
  const long long i = (long long)blockIdx.x * 32 + threadIdx.x;
  const bool ok = i < (long long)groups * C;
  const int g = ok ? (int)(i / C) : 0, c = ok ? (int)(i % C) : 0;
  float* outs[2] = {out0, out1};
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float s = 0.f;
    if (ok && outs[v])
      for (int k = threadIdx.y; k < chunks; k += 8) s += partial[(((long long)g * chunks + k) * NV + v) * C + c];
    red[v][threadIdx.y][threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.y == 0 && ok) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!outs[v]) continue;
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s += red[v][y][threadIdx.x];
      outs[v][i] = s * scale;
    }
  }
}

template <class F, int NV>
int colreduce(cgan_ctx* ctx, F f, int groups, long long rows_per_group, int C, float scale, float* out0, float* out1) {
  int cblocks = cdiv(C, CR_X);
  long long want = (4ll * ctx->num_sms + (long long)cblocks * groups - 1) / ((long long)cblocks * groups);
  long long maxchunks = (rows_per_group + 4 * CR_Y - 1) / (4 * CR_Y);
  long long chunks = want < 1 ? 1 : want;
  if (chunks > maxchunks) chunks = maxchunks;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  long long rpc = (rows_per_group + chunks - 1) / chunks;
  rpc = (rpc + CR_Y - 1) / CR_Y * CR_Y;
  chunks = (rows_per_group + rpc - 1) / rpc;
  if (chunks < 1) chunks = 1;
  if (groups > 65535) return cgan_fail(ctx, CGAN_ERR_ARG, "%s: too many groups%s", "colreduce");
  void* ws = nullptr;
  int rc = cgan_ws(ctx, (size_t)groups * chunks * NV * C * sizeof(float), &ws);
  if (rc) return rc;
  float* partial = reinterpret_cast<float*>(ws);
  dim3 grid(cblocks, (unsigned)chunks, groups), block(CR_X, CR_Y);
  if (ctx->counters && (long long)cblocks * groups <= CGAN_NUM_COUNTERS) {     // one launch: the last block per column block finishes
    colreduce_kernel<F, NV><<<grid, block, 0, ctx->stream>>>(f, partial, rows_per_group, C, rpc, (int)chunks, ctx->counters, scale,
                                                             out0, out1);
    CGAN_LAUNCHED(ctx);
    return CGAN_OK;
  }
  colreduce_kernel<F, NV><<<grid, block, 0, ctx->stream>>>(f, partial, rows_per_group, C, rpc, (int)chunks, nullptr, scale, out0, out1);
  CGAN_LAUNCHED(ctx);
  long long tot = (long long)groups * C;
  colreduce_final_kernel<NV><<<cdiv(tot, 32), dim3(32, 8), 0, ctx->stream>>>(partial, groups, (int)chunks, C, scale, out0, out1);
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}

__global__ void bn_finalize_kernel(float* mean_var, const float* stats, int C, float* mm, float* mv, float decay) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean = stats[c], msq = stats[C + c];
  float var = msq - mean * mean;
  mean_var[c] = mean;
  mean_var[C + c] = var;
  if (mm) mm[c] -= (mm[c] - mean) * (1.0f - decay);
  if (mv) mv[c] -= (mv[c] - var) * (1.0f - decay);
}

__global__ void bn_accumulate_kernel(float* mean_var, const float* batch, int C, float* am, float* av, float* ac,
                                     const float* upd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  bool update = (*upd == 1.0f);
  float counter = *ac + (update ? 1.0f : 0.0f);
  if (c < C) {
    float m = am[c], v = av[c];
    if (update) {
      m += batch[c];
      v += batch[C + c];
      am[c] = m;
      av[c] = v;
    }
    mean_var[c] = m / counter;
    mean_var[C + c] = v / counter;
  }
}

// runs after bn_accumulate_kernel on the same stream: every block of that kernel read the OLD counter
__global__ void bn_accu_counter_kernel(float* ac, const float* upd) {
  if (*upd == 1.0f) *ac += 1.0f;
}

__device__ __forceinline__ float rna_tf32n(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// act: bit 0 = ReLU, CGAN_ACT_ROUND_TF32 = store TF32-rounded values
__global__ void bn_apply_kernel(float* __restrict__ y, const float* __restrict__ x, long long total, int C,
                                long long rows_per_sample, const float* __restrict__ mean_var, float eps,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int cond, int act) {
  const int relu = act & 1, rnd = (act & CGAN_ACT_ROUND_TF32) ? 1 : 0;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % C);
    long long r = i / C;
    long long pidx = cond ? (r / rows_per_sample) * C + c : c;
    float inv = 1.0f / sqrtf(mean_var[C + c] + eps);
    float v = (x[i] - mean_var[c]) * inv;
    if (gamma) v *= gamma[pidx];
    if (beta) v += beta[pidx];
    if (relu) v = fmaxf(v, 0.f);
    y[i] = rnd ? rna_tf32n(v) : v;
  }
}

// float4 form (C % 4 == 0, fewer than 2^31 vectors, 16-byte aligned tensors): one 32-bit division per four elements, the
// per-channel parameters come in as float4s from L1
__global__ void bn_apply_v4_kernel(float4* __restrict__ y, const float4* __restrict__ x, unsigned total4, unsigned lanes,
                                   unsigned rows_per_sample, const float* __restrict__ mean_var, float eps,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int cond, int act) {
  const int relu = act & 1, rnd = (act & CGAN_ACT_ROUND_TF32) ? 1 : 0;
  const unsigned C = lanes * 4;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const unsigned r = i / lanes, c = (i - r * lanes) * 4;
    const float4 m = *reinterpret_cast<const float4*>(mean_var + c);
    const float4 vv = *reinterpret_cast<const float4*>(mean_var + C + c);
    float4 v = x[i];
    v.x = (v.x - m.x) * (1.0f / sqrtf(vv.x + eps)); v.y = (v.y - m.y) * (1.0f / sqrtf(vv.y + eps));
    v.z = (v.z - m.z) * (1.0f / sqrtf(vv.z + eps)); v.w = (v.w - m.w) * (1.0f / sqrtf(vv.w + eps));
    const size_t pidx = cond ? (size_t)(r / rows_per_sample) * C + c : c;
    if (gamma) { const float4 g = *reinterpret_cast<const float4*>(gamma + pidx); v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w; }
    if (beta) { const float4 b = *reinterpret_cast<const float4*>(beta + pidx); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (rnd) { v.x = rna_tf32n(v.x); v.y = rna_tf32n(v.y); v.z = rna_tf32n(v.z); v.w = rna_tf32n(v.w); }
    y[i] = v;
  }
}

// sums[c] = sum_g gamma(g,c)*B(g,c); sums[C+c] = sum_g gamma(g,c)*A(g,c);  A = sum dy*xhat, B = sum dy (per group)
__global__ void bn_bwd_sums_kernel(float* sums, const float* A, const float* Bv, const float* gamma, int groups, int C,
                                   int cond) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s1 = 0.f, s2 = 0.f;
  for (int g = 0; g < groups; ++g) {
    float gm = gamma ? (cond ? gamma[(long long)g * C + c] : gamma[c]) : 1.0f;
    s1 += gm * Bv[(long long)g * C + c];
    s2 += gm * A[(long long)g * C + c];
  }
  sums[c] = s1;
  sums[C + c] = s2;
}

__global__ void bn_bwd_apply_kernel(float* __restrict__ dx, const float* __restrict__ dy, const float* __restrict__ x,
                                    long long total, int C, long long rows_per_sample,
                                    const float* __restrict__ mean_var, float eps, const float* __restrict__ gamma,
                                    int cond, const float* __restrict__ sums, float inv_count, int rnd) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % C);
    long long r = i / C;
    float inv = 1.0f / sqrtf(mean_var[C + c] + eps);
    float xh = (x[i] - mean_var[c]) * inv;
    float g = gamma ? gamma[cond ? (r / rows_per_sample) * C + c : c] : 1.0f;
    float dxh = dy[i] * g;
    float v = inv * (dxh - sums[c] * inv_count - xh * sums[C + c] * inv_count);
    dx[i] = rnd ? rna_tf32n(v) : v;
  }
}

__global__ void bn_bwd_apply_v4_kernel(float4* __restrict__ dx, const float4* __restrict__ dy, const float4* __restrict__ x,
                                       unsigned total4, unsigned lanes, unsigned rows_per_sample,
                                       const float* __restrict__ mean_var, float eps, const float* __restrict__ gamma,
                                       int cond, const float* __restrict__ sums, float inv_count, int rnd) {
  const unsigned C = lanes * 4;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const unsigned r = i / lanes, c = (i - r * lanes) * 4;
    const float4 m = *reinterpret_cast<const float4*>(mean_var + c);
    const float4 vv = *reinterpret_cast<const float4*>(mean_var + C + c);
    const float4 s1 = *reinterpret_cast<const float4*>(sums + c);
    const float4 s2 = *reinterpret_cast<const float4*>(sums + C + c);
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
    if (gamma) g = *reinterpret_cast<const float4*>(gamma + (cond ? (size_t)(r / rows_per_sample) * C + c : c));
    const float4 xv = x[i], gy = dy[i];
    float4 o;
    {
      const float inv = 1.0f / sqrtf(vv.x + eps), xh = (xv.x - m.x) * inv;
      o.x = inv * (gy.x * g.x - s1.x * inv_count - xh * s2.x * inv_count);
    }
    {
      const float inv = 1.0f / sqrtf(vv.y + eps), xh = (xv.y - m.y) * inv;
      o.y = inv * (gy.y * g.y - s1.y * inv_count - xh * s2.y * inv_count);
    }
    {
      const float inv = 1.0f / sqrtf(vv.z + eps), xh = (xv.z - m.z) * inv;
      o.z = inv * (gy.z * g.z - s1.z * inv_count - xh * s2.z * inv_count);
    }
    {
      const float inv = 1.0f / sqrtf(vv.w + eps), xh = (xv.w - m.w) * inv;
      o.w = inv * (gy.w * g.w - s1.w * inv_count - xh * s2.w * inv_count);
    }
    if (rnd) { o.x = rna_tf32n(o.x); o.y = rna_tf32n(o.y); o.z = rna_tf32n(o.z); o.w = rna_tf32n(o.w); }
    dx[i] = o;
  }
}

// one warp per row: out[r] = sum_c w[r,c]*b[c]
__global__ void gemv_rows_kernel(float* __restrict__ out, const float* __restrict__ w, const float* __restrict__ b,
                                 int rows, int cols) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* row = w + (long long)warp * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += row[c] * b[c];
  s = warp_sum(s);
  if (lane == 0) out[warp] = s;
}

// out = t * rsqrt(max(sum t^2, eps)); optionally sigma = dot(out, t).  Single block.
__global__ void l2_normalize_kernel(float* out, const float* t, int n, float eps, float* sigma) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += t[i] * t[i];
  s = block_sum(s, sh);
  float scale = 1.0f / sqrtf(fmaxf(s, eps));
  float d = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float o = t[i] * scale;
    out[i] = o;
    d += o * t[i];
  }
  if (sigma) {
    d = block_sum(d, sh);
    if (threadIdx.x == 0) *sigma = d;
  }
}

__global__ void sn_bwd_kernel(float* __restrict__ dw, const float* __restrict__ dwbar, int rows, int cols, int left,
                              const float* __restrict__ u, const float* __restrict__ v, const float* sigma,
                              const float* dotp) {
  long long total = (long long)rows * cols;
  long long stride = (long long)gridDim.x * blockDim.x;
  float inv_s = 1.0f / *sigma, dp = *dotp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int r = (int)(i / cols), c = (int)(i % cols);
    float outer = left ? u[r] * v[c] : v[r] * u[c];
    dw[i] = (dwbar[i] - dp * outer) * inv_s;
  }
}

// ---- batched spectral norm: ONE launch runs the power iteration of every small weight of a network ----------------
// A discriminator call site of resnet_cifar10 / SNDCGAN touches 8-13 spectrally normalised kernels of at most a few
// hundred KB each; run separately that is ~7 tiny launches per kernel (two GEMVs with their finishing passes, two
// normalisations, the scale, the copy of u) = hundreds of launches per training cycle.  Here one CTA per weight does the
// whole of arch_ops.py:503-531 out of L2: t = W^T u (or W u), v = normalize(t), s = W v (or v W), u' = normalize(s),
// sigma = <u', s>, wbar = W / sigma, u <- u', plus the copy of u' the backward needs.  Fixed summation order.
__device__ __forceinline__ void sn_coldot(float* out, const float* __restrict__ w, const float* x, int rows, int cols, float* red) {
  // out[c] = sum_r w[r, c] * x[r]; thread (c, l): column c of a 32..1024-wide block, row lane l
  for (int cb = 0; cb < cols; cb += 1024) {
    const int cw = min(1024, cols - cb);
    int cpad = 32;
    while (cpad < cw) cpad <<= 1;
    const int lanes = 1024 / cpad;
    const int c = threadIdx.x % cpad, l = threadIdx.x / cpad;
    float acc = 0.f;
    if (c < cw)
      for (int r = l; r < rows; r += lanes) acc = fmaf(w[(size_t)r * cols + cb + c], x[r], acc);
    red[threadIdx.x] = acc;
    __syncthreads();
    if (l == 0 && c < cw) {
      float sum = 0.f;
      for (int j = 0; j < lanes; ++j) sum += red[j * cpad + c];
      out[cb + c] = sum;
    }
    __syncthreads();
  }
}
__device__ __forceinline__ void sn_rowdot(float* out, const float* __restrict__ w, const float* x, int rows, int cols) {
  // out[r] = sum_c w[r, c] * x[c]; one warp per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < rows; r += 32) {
    float acc = 0.f;
    for (int c = lane; c < cols; c += 32) acc = fmaf(w[(size_t)r * cols + c], x[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[r] = acc;
  }
  __syncthreads();
}
__device__ __forceinline__ float sn_normalize(float* v, int n, float eps, float* sh, float* dot_with_raw) {
  // v <- v * rsqrt(max(sum v^2, eps)); returns via *dot_with_raw the dot product of the normalised and the raw vector
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i] * v[i];
  s = block_sum(s, sh);
  const float scale = 1.0f / sqrtf(fmaxf(s, eps));
  float d = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float raw = v[i], o = raw * scale;
    v[i] = o;
    d += o * raw;
  }
  d = block_sum(d, sh);
  if (dot_with_raw) *dot_with_raw = d;
  return scale;
}

__global__ void __launch_bounds__(1024, 1)
sn_batched_kernel(const cgan_sn_item* __restrict__ items, float eps, float* wbar_base, float* v_base, float* sigma_base,
                  float* u_used_base) {
  extern __shared__ float sn_smem[];
  __shared__ float sh[32];
  const cgan_sn_item it = items[blockIdx.x];
  const int rows = it.rows, cols = it.cols;
  const int nu = it.left ? rows : cols, nv = it.left ? cols : rows;
  float* red = sn_smem;                 // 1024 floats
  float* uv = red + 1024;               // u (nu floats)
  float* vv = uv + nu;                  // v (nv floats)
  for (int i = threadIdx.x; i < nu; i += blockDim.x) uv[i] = it.u[i];
  __syncthreads();
  if (it.left) sn_coldot(vv, it.w, uv, rows, cols, red); else sn_rowdot(vv, it.w, uv, rows, cols);
  sn_normalize(vv, nv, eps, sh, nullptr);
  __syncthreads();
  if (it.left) sn_rowdot(uv, it.w, vv, rows, cols); else sn_coldot(uv, it.w, vv, rows, cols, red);
  float sigma;
  sn_normalize(uv, nu, eps, sh, &sigma);
  __syncthreads();
  float* v_out = v_base + it.v_off;
  float* u_used = u_used_base + it.u_off;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) v_out[i] = vv[i];
  for (int i = threadIdx.x; i < nu; i += blockDim.x) {
    const float un = uv[i];
    it.u[i] = un;
    u_used[i] = un;
  }
  if (threadIdx.x == 0) sigma_base[blockIdx.x] = sigma;
  float* wbar = wbar_base + it.wbar_off;
  const size_t n = (size_t)rows * cols;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) wbar[i] = it.w[i] / sigma;    // "w / norm_value" (arch_ops.py:531)
}

inline bool v4_ok(const void* a, const void* b, const void* c, const void* d, const void* e, const void* f) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(f)) & 15) == 0;
}
inline int ew_grid(cgan_ctx* ctx, long long n) {
  long long b = (n + 255) / 256;
  long long cap = (long long)ctx->num_sms * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int cgan_colsum(cgan_ctx* ctx, float* out, const float* x, int groups, int64_t rows_per_group, int c) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, out && x && groups > 0 && rows_per_group > 0 && c > 0, "bad argument");
  return colreduce<SumF, 1>(ctx, SumF{x}, groups, rows_per_group, c, 1.0f, out, nullptr);
}

int cgan_bn_moments(cgan_ctx* ctx, float* stats2c, const float* x, int64_t rows, int c) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, stats2c && x && rows > 0 && c > 0, "bad argument");
  return colreduce<MomentsF, 2>(ctx, MomentsF{x}, 1, rows, c, 1.0f / (float)rows, stats2c, stats2c + c);
}

int cgan_bn_finalize(cgan_ctx* ctx, float* mean_var2c, const float* stats2c, int c, float* moving_mean,
                     float* moving_var, float decay) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, mean_var2c && stats2c && c > 0, "bad argument");
  bn_finalize_kernel<<<cdiv(c, 256), 256, 0, ctx->stream>>>(mean_var2c, stats2c, c, moving_mean, moving_var, decay);
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}

int cgan_bn_accumulate(cgan_ctx* ctx, float* mean_var2c, const float* batch, int c, float* accu_mean, float* accu_var,
                       float* accu_counter, const float* update_accus_dev) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, mean_var2c && batch && accu_mean && accu_var && accu_counter && update_accus_dev, "null pointer");
  CGAN_REQUIRE(ctx, c > 0, "channels must be positive");
  bn_accumulate_kernel<<<cdiv(c, 256), 256, 0, ctx->stream>>>(mean_var2c, batch, c, accu_mean, accu_var, accu_counter,
                                                              update_accus_dev);
  CGAN_LAUNCHED(ctx);
  bn_accu_counter_kernel<<<1, 1, 0, ctx->stream>>>(accu_counter, update_accus_dev);
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}

int cgan_bn_apply(cgan_ctx* ctx, float* y, const float* x, int64_t rows, int c, int64_t rows_per_sample,
                  const float* mean_var2c, float eps, const float* gamma, const float* beta, int cond, int act) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, y && x && mean_var2c && rows > 0 && c > 0, "bad argument");
  CGAN_REQUIRE(ctx, !cond || (rows_per_sample > 0 && rows % rows_per_sample == 0), "rows_per_sample must divide rows");
  CGAN_REQUIRE(ctx, (act & ~(1 | CGAN_ACT_ROUND_TF32)) == 0, "act must be 0 / 1 (ReLU), optionally | CGAN_ACT_ROUND_TF32");
  long long total = (long long)rows * c;
  const long long rps = rows_per_sample > 0 ? rows_per_sample : 1;
  if (c % 4 == 0 && total / 4 < (1ll << 31) && rps < (1ll << 31) && v4_ok(y, x, mean_var2c, gamma, beta, nullptr)) {
    bn_apply_v4_kernel<<<ew_grid(ctx, total / 4), 256, 0, ctx->stream>>>(
        reinterpret_cast<float4*>(y), reinterpret_cast<const float4*>(x), (unsigned)(total / 4), (unsigned)(c / 4), (unsigned)rps,
        mean_var2c, eps, gamma, beta, cond, act);
  } else {
    bn_apply_kernel<<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(y, x, total, c, rps, mean_var2c, eps, gamma, beta, cond, act);
  }
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}

int cgan_bn_bwd_reduce(cgan_ctx* ctx, float* sums2c, float* dgamma, float* dbeta, const float* dy, const float* x,
                       int64_t rows, int c, int64_t rows_per_sample, const float* mean_var2c, float eps,
                       const float* gamma, int cond) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, sums2c && dy && x && mean_var2c && rows > 0 && c > 0, "bad argument");
  CGAN_REQUIRE(ctx, !cond || (rows_per_sample > 0 && rows % rows_per_sample == 0), "rows_per_sample must divide rows");
  int groups = cond ? (int)(rows / rows_per_sample) : 1;
  long long rpg = cond ? rows_per_sample : rows;
  // per-group A = sum dy*xhat, B = sum dy; kept in caller-visible dgamma/dbeta or in workspace tail
  float *A = dgamma, *Bv = dbeta;
  void* ws = nullptr;
  if (!A || !Bv) {
    // scratch behind the colreduce partials: reserve generously first so the pointer stays valid
    size_t need = (size_t)groups * c * 2 * sizeof(float);
    int rc = cgan_ws(ctx, need + (size_t)64 * 1024 * 1024, &ws);
    if (rc) return rc;
    float* tail = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ctx->ws_bytes - need);
    if (!A) A = tail;
    if (!Bv) Bv = tail + (size_t)groups * c;
  }
  int rc = colreduce<BnBwdF, 2>(ctx, BnBwdF{dy, x, mean_var2c, eps}, groups, rpg, c, 1.0f, A, Bv);
  if (rc) return rc;
  bn_bwd_sums_kernel<<<cdiv(c, 256), 256, 0, ctx->stream>>>(sums2c, A, Bv, gamma, groups, c, cond);
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}

int cgan_bn_bwd_apply(cgan_ctx* ctx, float* dx, const float* dy, const float* x, int64_t rows, int c,
                      int64_t rows_per_sample, const float* mean_var2c, float eps, const float* gamma, int cond,
                      const float* sums2c, float inv_count, int round_tf32) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, dx && dy && x && mean_var2c && sums2c && rows > 0 && c > 0, "bad argument");
  long long total = (long long)rows * c;
  const long long rps = rows_per_sample > 0 ? rows_per_sample : 1;
  if (c % 4 == 0 && total / 4 < (1ll << 31) && rps < (1ll << 31) && v4_ok(dx, dy, x, mean_var2c, gamma, sums2c)) {
    bn_bwd_apply_v4_kernel<<<ew_grid(ctx, total / 4), 256, 0, ctx->stream>>>(
        reinterpret_cast<float4*>(dx), reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(x),
        (unsigned)(total / 4), (unsigned)(c / 4), (unsigned)rps, mean_var2c, eps, gamma, cond, sums2c, inv_count, round_tf32 ? 1 : 0);
  } else {
    bn_bwd_apply_kernel<<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(dx, dy, x, total, c, rps, mean_var2c, eps, gamma, cond,
                                                                      sums2c, inv_count, round_tf32 ? 1 : 0);
  }
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}

// internal: out[c] = sum_r w[r,c]*a[r]
static int gemv_cols(cgan_ctx* ctx, float* out, const float* w, const float* a, int rows, int cols) {
  return colreduce<WeightedSumF, 1>(ctx, WeightedSumF{w, a}, 1, rows, cols, 1.0f, out, nullptr);
}
static int gemv_rows(cgan_ctx* ctx, float* out, const float* w, const float* b, int rows, int cols) {
  gemv_rows_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, ctx->stream>>>(out, w, b, rows, cols);
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}

int cgan_scale_by_dev(cgan_ctx*, float*, const float*, const float*, float, int, int64_t);

int cgan_spectral_norm(cgan_ctx* ctx, const float* w, int rows, int cols, int left, float eps, float* u, float* v,
                       float* sigma, float* wbar) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, w && u && v && sigma && rows > 0 && cols > 0, "bad argument");
  // scratch vector t (max(rows, cols)) lives in the workspace tail, clear of the colreduce partials
  size_t tbytes = (size_t)(rows > cols ? rows : cols) * sizeof(float);
  void* ws = nullptr;
  int rc = cgan_ws(ctx, tbytes + (size_t)64 * 1024 * 1024, &ws);
  if (rc) return rc;
  float* t = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ctx->ws_bytes - ((tbytes + 255) / 256) * 256);
  if (left) {
    // v = normalize(W^T u); u' = normalize(W v); sigma = u'^T W v       (arch_ops.py:505-509, 525)
    rc = gemv_cols(ctx, t, w, u, rows, cols);
    if (rc) return rc;
    l2_normalize_kernel<<<1, 1024, 0, ctx->stream>>>(v, t, cols, eps, nullptr);
    CGAN_LAUNCHED(ctx);
    rc = gemv_rows(ctx, t, w, v, rows, cols);
    if (rc) return rc;
    l2_normalize_kernel<<<1, 1024, 0, ctx->stream>>>(u, t, rows, eps, sigma);
    CGAN_LAUNCHED(ctx);
  } else {
    // v = normalize(u W^T); u' = normalize(v W); sigma = v W u'^T       (arch_ops.py:511-513, 527)
    rc = gemv_rows(ctx, t, w, u, rows, cols);
    if (rc) return rc;
    l2_normalize_kernel<<<1, 1024, 0, ctx->stream>>>(v, t, rows, eps, nullptr);
    CGAN_LAUNCHED(ctx);
    rc = gemv_cols(ctx, t, w, v, rows, cols);
    if (rc) return rc;
    l2_normalize_kernel<<<1, 1024, 0, ctx->stream>>>(u, t, cols, eps, sigma);
    CGAN_LAUNCHED(ctx);
  }
  if (wbar) return cgan_scale_by_dev(ctx, wbar, w, sigma, 1.0f, 1, (int64_t)rows * cols);
  return CGAN_OK;
}

int cgan_spectral_norm_batched(cgan_ctx* ctx, const cgan_sn_item* items_dev, int n, int max_rows_plus_cols, float eps,
                               float* wbar_base, float* v_base, float* sigma_base, float* u_used_base) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, items_dev && n > 0 && wbar_base && v_base && sigma_base && u_used_base, "bad argument");
  const size_t smem = (size_t)(1024 + max_rows_plus_cols) * sizeof(float);
  CGAN_REQUIRE(ctx, smem <= 200 * 1024, "rows + cols too large for the batched kernel (use cgan_spectral_norm)");
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    CGAN_CUDA(ctx, cudaFuncSetAttribute(sn_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  sn_batched_kernel<<<n, 1024, smem, ctx->stream>>>(items_dev, eps, wbar_base, v_base, sigma_base, u_used_base);
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}

int cgan_spectral_norm_bwd(cgan_ctx* ctx, float* dw, const float* dwbar, const float* wbar, int rows, int cols,
                           int left, const float* u, const float* v, const float* sigma) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, dw && dwbar && wbar && u && v && sigma && rows > 0 && cols > 0, "bad argument");
  void* ws = nullptr;
  int rc = cgan_ws(ctx, (size_t)64 * 1024 * 1024, &ws);
  if (rc) return rc;
  float* dotp = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ctx->ws_bytes - 256);
  rc = cgan_dot(ctx, dotp, dwbar, wbar, (int64_t)rows * cols);
  if (rc) return rc;
  long long total = (long long)rows * cols;
  sn_bwd_kernel<<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(dw, dwbar, rows, cols, left, u, v, sigma, dotp);
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}
