// tcgen05 filter-gradient kernel (math_mode 1):
//
//   dW[tap][ci][co] = sum_pixels  X[pixel + off(tap)][ci] * dY[pixel][co]
//
// A GEMM whose reduction dimension is the PIXEL index, so both operands are "MN-major" as they lie in HBM (NHWC:
// channels contiguous): D[ci, co] += A[ci, pix] * B[pix, co].  Per k-block of 32 pixels, TMA brings
//   * 4 boxes  [32 pixels x 32 ci]  of X  (shifted by the tap offset; SAME padding = TMA zero fill), and
//   * N/32 boxes [32 pixels x 32 co] of dY
// each box = 32 rows x 128 B.  For 32-bit MN-major operands the tensor core only accepts the SWIZZLE_128B_BASE32B
// layout (cute Layout_MN_SW128_32B_Atom: Swizzle<2,5,2>, 128 B x 4-row atoms, 32 B swizzle granularity), which TMA
// produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; channel groups sit LBO = 4096 B apart, 4-pixel row groups
// SBO = 512 B apart.  Four tcgen05.mma (M=128, N<=256, K=8 pixels) consume a k-block.  The epilogue warps round
// both operands to nearest TF32 in shared memory while the tensor core works on earlier stages.
// The pixel range is split across CTAs (split-K); partial tiles go to a workspace and are summed in a fixed order
// (deterministic), replacing TF's Conv2DBackpropFilter.
// A conv over a zero-inserted 2x-upsampled input (resnet_ops.py:35-56) is handled through four strided TMA views of dY
// (one per sub-pixel phase); every tap belongs to exactly one phase.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int WG_MAX_STAGES = 4;
constexpr int WG_P = 32;                 // pixels per k-block
constexpr int WG_BOX = WG_P * 128;       // 4 KB: 32 pixel rows x 32 channels fp32
constexpr int WG_A_BYTES = 4 * WG_BOX;   // 128 input channels
constexpr int WG_RWARPS = 8;         // warps 2..9: operand rounding, then the epilogue (two warps per TMEM lane quarter)
constexpr int WG_THREADS = 64 + 32 * WG_RWARPS;
constexpr int WG_MAX_TAPS = 16;

struct WgParams {
  int ntaps;
  int off_h[WG_MAX_TAPS], off_w[WG_MAX_TAPS], amap[WG_MAX_TAPS], bmap[WG_MAX_TAPS], wtap[WG_MAX_TAPS];
  int bw, bh, bni, tiles_w, tiles_h;     // 32-pixel box geometry
  int kblocks, kb_per_split;
  int ci_tiles, co_tiles, bn;
  int cin, cout, taps_total;
  int stages, tmem_cols;                 // pipeline depth chosen so that two CTAs share an SM; TMEM columns = pow2 >= mt*bn
  int mt;                                // (tap, ci-tile) units per CTA that share one dY tile (mt accumulators in TMEM)
  int round_a, round_b;                  // round the X / dY tiles to nearest TF32 in shared memory (operand not pre-rounded)
  float* partial;                        // [split][taps_total][cin][cout]
};

struct BMaps { CUtensorMap m[4]; };

// MN-major descriptor for 32-bit operands: start>>4 | LBO>>4 (stride between 32-channel groups) | SBO>>4 (stride
// between 4-pixel row groups) | version 1 | layout SWIZZLE_128B_BASE32B (=1)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(WG_BOX >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 2)
wgrad_tc_kernel(const __grid_constant__ BMaps tm_x, const __grid_constant__ BMaps tm_dy, const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = (p.bn / 32) * WG_BOX;
  const int a_bytes = p.mt * WG_A_BYTES;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* ready_bar = full_bar + p.stages;
  uint64_t* empty_bar = ready_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work units (tap, ci tile); this CTA owns units [u0, u0 + nu) — they all contract against the same dY tile, which is
  // fetched from L2 once per k-block for all of them (the kernel is bound by L2->SM bytes per MMA)
  int t = blockIdx.x;
  const int co_t = t % p.co_tiles;
  const int u0 = (t / p.co_tiles) * p.mt;
  const int nu = min(p.mt, p.ntaps * p.ci_tiles - u0);
  const int tap = u0 / p.ci_tiles;               // unit 0's tap: selects the dY view (equal for all units, host-checked)
  const int split = blockIdx.y;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
  const int num_kb = kb1 - kb0;                 // >= 1 by construction
  const int co0 = co_t * p.bn;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x.m[p.amap[tap]]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_dy.m[p.bmap[tap]]) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&ready_bar[s], WG_RWARPS);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mb = &tm_dy.m[p.bmap[tap]];
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        int r = kb;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int tn = r / p.tiles_h;
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bni;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_expect_tx(&full_bar[stage], (uint32_t)(nu * WG_A_BYTES + b_bytes));
        for (int i = 0; i < nu; ++i) {
          const int u = u0 + i, utap = u / p.ci_tiles, ci0 = (u % p.ci_tiles) * 128;
          const CUtensorMap* ma = &tm_x.m[p.amap[utap]];
#pragma unroll
          for (int g = 0; g < 4; ++g)
            tma_load_4d(sa + i * WG_A_BYTES + g * WG_BOX, ma, &full_bar[stage], ci0 + g * 32, w0 + p.off_w[utap],
                        h0 + p.off_h[utap], n0);
        }
        for (int g = 0; g < p.bn / 32; ++g) tma_load_4d(sb + g * WG_BOX, mb, &full_bar[stage], co0 + g * 32, w0, h0, n0);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // D=F32, A=B=TF32, both MN-major (bits 15, 16), N>>3, M=128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.bn >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait((p.round_a | p.round_b) ? &ready_bar[stage] : &full_bar[stage], phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
        const uint32_t b_addr = a_addr + a_bytes;
        for (int i = 0; i < nu; ++i) {
#pragma unroll
          for (int k = 0; k < WG_P / 8; ++k)          // 8 pixel rows (1024 B) per MMA
            umma_tf32(tmem_base + (uint32_t)(i * p.bn), make_desc_mn(a_addr + i * WG_A_BYTES + k * 1024),
                      make_desc_mn(b_addr + k * 1024), idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    const int q = threadIdx.x - 64;
    if (p.round_a | p.round_b) {
      int stage = 0;
      uint32_t phase = 0;
      // only the operand(s) that are not pre-rounded are swept: [0, a_bytes) is X, [a_bytes, stage_bytes) is dY
      const int i0 = p.round_a ? 0 : a_bytes / 16;
      const int n4 = (p.round_b ? stage_bytes : a_bytes) / 16;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t s4 = smem_u32(smem + stage * stage_bytes);
        // the ranges are multiples of 4 KB (32-pixel boxes of 128 B rows): 256 threads x 16 B per sweep
#pragma unroll 4
        for (int i = i0 + q; i < n4; i += 32 * WG_RWARPS) {
          float4 v = lds128(s4 + i * 16);
          v.x = rna_tf32(v.x); v.y = rna_tf32(v.y); v.z = rna_tf32(v.z); v.w = rna_tf32(v.w);
          sts128(s4 + i * 16, v);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;           // ci within the tile
    for (int i = 0; i < nu; ++i) {
    const int u = u0 + i, utap = u / p.ci_tiles, ci0 = (u % p.ci_tiles) * 128;
    float* orow = p.partial + (((long long)split * p.taps_total + p.wtap[utap]) * p.cin + ci0 + row) * p.cout + co0;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * p.bn);
    const bool row_ok = ci0 + row < p.cin;       // the last ci tile may hang over Cin (TMA zero-filled those channels)
    // two warps share a lane quarter: the first takes the lower half of the 32-column chunks, the second the rest
    const int nchunks = p.bn / 32, csplit = (nchunks + 1) / 2;
    const int cbeg = (warp - 2) < 4 ? 0 : csplit * 32, cend = (warp - 2) < 4 ? csplit * 32 : p.bn;
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + (uint32_t)c0, r);       // warp-collective: every lane takes part, stores are predicated
      if (row_ok) {
        if (co0 + c0 + 32 <= p.cout && (p.cout & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(orow + c0 + j) =
                make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        } else {        // Cout zero-padded to the 32-column tile (e.g. the 24-wide attention projections)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (co0 + c0 + j < p.cout) orow[c0 + j] = __uint_as_float(r[j]);
        }
      }
    }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

__global__ void wg_reduce_kernel(float* __restrict__ out, const float* __restrict__ part, long long n, int splits) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(long long)z * n + i];
  out[i] = s;
}

bool box32(int n, int h, int w, int* bw, int* bh, int* bni) {
  int b = w < 32 ? w : 32;
  if (32 % b != 0 || w % b != 0) return false;
  int hh = 32 / b; if (hh > h) hh = h;
  if (h % hh != 0) return false;
  int ni = 32 / (b * hh);
  if (ni < 1 || b * hh * ni != 32 || n % ni != 0) return false;
  *bw = b; *bh = hh; *bni = ni;
  return true;
}

int pick_bn(int ncols) {
  if (ncols <= 256 && ncols % 4 == 0) return (ncols + 31) / 32 * 32;     // zero-padded tile, stores are masked
  if (ncols % 32) return 0;
  if (ncols <= 256) return ncols;
  if (ncols % 256 == 0) return 256;
  if (ncols % 192 == 0) return 192;
  if (ncols % 128 == 0) return 128;
  return 0;
}

}  // namespace

// can the pixel loop of a stride-1 n x h x w grid be cut into 32-pixel TMA boxes?
bool cgan_wgrad_tc_geometry_ok(int n, int h, int w) {
  int bw, bh, bni;
  return box32(n, h, w, &bw, &bh, &bni);
}

bool cgan_wgrad_tc_ok(const cgan_conv_desc* d) {
  if ((d->stride != 1 && d->stride != 2) || d->kh * d->kw > WG_MAX_TAPS) return false;
  if (d->stride == 2) {
    if (d->upsample || (d->h & 1) || (d->w & 1) || d->oh != d->h / 2 || d->ow != d->w / 2) return false;
    if (d->cin % 32 != 0 || d->cin < 64 || pick_bn(d->cout) == 0) return false;
    int bw, bh, bni;
    return box32(d->n, d->oh, d->ow, &bw, &bh, &bni);
  }
  if (d->cin % 32 != 0 || d->cin < 64 || pick_bn(d->cout) == 0) return false;   // Cin tiles of 128, last one zero-padded
  if (d->oh != (d->upsample ? 2 * d->h : d->h) || d->ow != (d->upsample ? 2 * d->w : d->w)) return false;
  int bw, bh, bni;
  return box32(d->n, d->h, d->w, &bw, &bh, &bni);
}

int cgan_wgrad_tc(cgan_ctx* ctx, const cgan_conv_desc* d, const float* x, const float* dy, float* dw, int x_tf32, int dy_tf32) {
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.round_a = x_tf32 ? 0 : 1;
  p.round_b = dy_tf32 ? 0 : 1;
  const int gh = d->stride == 2 ? d->oh : d->h, gw = d->stride == 2 ? d->ow : d->w;     // pixel-loop grid
  if (!box32(d->n, gh, gw, &p.bw, &p.bh, &p.bni)) return cgan_fail(ctx, CGAN_ERR_UNSUPPORTED, "%s: geometry%s", "cgan_wgrad_tc");
  p.tiles_w = gw / p.bw;
  p.tiles_h = gh / p.bh;
  p.kblocks = (int)((long long)d->n * gh * gw / WG_P);
  p.bn = pick_bn(d->cout);
  p.ci_tiles = (d->cin + 127) / 128;
  p.co_tiles = (d->cout + p.bn - 1) / p.bn;
  p.cin = d->cin; p.cout = d->cout; p.taps_total = d->kh * d->kw;
  // taps: the pixel loop runs over a grid (i, j) of gh x gw cells: the input grid (stride 1, incl. zero-inserted
  // inputs) or the output grid (stride 2); tap (kh,kw) pairs X view `amap` at [i+dh, j+dw] with dY view `bmap` at [i, j]
  int nt = 0;
  for (int kh = 0; kh < d->kh; ++kh)
    for (int kw = 0; kw < d->kw; ++kw) {
      p.amap[nt] = 0;
      if (d->stride == 2) {
        // output (i,j) reads input row 2i + kh - pad_t = 2(i + dh) + a: phase view a of X
        int th = kh - d->pad_t, tw = kw - d->pad_l, a = th & 1, b = tw & 1;
        p.off_h[nt] = (th - a) / 2; p.off_w[nt] = (tw - b) / 2; p.amap[nt] = a * 2 + b; p.bmap[nt] = 0;
      } else if (!d->upsample) {
        p.off_h[nt] = kh - d->pad_t; p.off_w[nt] = kw - d->pad_l; p.bmap[nt] = 0;
      } else {
        // output row 2i+a reads virtual row 2i+a+kh-pad_t, real only when even: a = (pad_t - kh) & 1, dh = (a+kh-pad_t)/2
        int a = (d->pad_t - kh) & 1, b = (d->pad_l - kw) & 1;
        p.off_h[nt] = (a + kh - d->pad_t) / 2; p.off_w[nt] = (b + kw - d->pad_l) / 2; p.bmap[nt] = a * 2 + b;
      }
      p.wtap[nt] = kh * d->kw + kw;
      ++nt;
    }
  p.ntaps = nt;
  // two (tap, ci-tile) units per CTA when they read the same dY view: dY is then fetched once per k-block for both
  const int units = p.ci_tiles * nt;
  bool same_b = true;
  for (int i = 1; i < nt; ++i) same_b = same_b && p.bmap[i] == p.bmap[0];
  p.mt = (ctx->tc_mt_max >= 2 && same_b && units >= 2) ? 2 : 1;
  const bool two_ctas = p.mt * p.bn <= 256;      // TMEM: mt*bn accumulator columns per CTA, 512 per SM
  long long tiles = (long long)p.co_tiles * ((units + p.mt - 1) / p.mt);
  // one wave of CTAs, rounded DOWN so the grid never spills a few CTAs into an extra wave
  int splits = (int)(((two_ctas ? 2ll : 1ll) * ctx->num_sms) / tiles);
  int max_splits = p.kblocks / 8 > 0 ? p.kblocks / 8 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kblocks + splits - 1) / splits;
  splits = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;

  long long wn = (long long)p.taps_total * d->cin * d->cout;
  float* partial = dw;
  if (splits > 1) {
    void* ws = nullptr;
    int rc = cgan_ws(ctx, (size_t)splits * wn * sizeof(float), &ws);
    if (rc) return rc;
    partial = reinterpret_cast<float*>(ws);
  }
  p.partial = partial;

  BMaps tm_x, tm_dy;
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_dy, 0, sizeof(tm_dy));
  for (int v = 0; v < 4; ++v) {
    bool ok;
    if (d->stride == 2) {     // X seen through its four stride-2 phases, each of the OUTPUT's spatial size
      int a = v >> 1, b = v & 1;
      ok = make_act_map(&tm_x.m[v], x + ((long long)a * d->w + b) * d->cin, d->cin, gw, gh, d->n, 2ll * d->cin,
                        2ll * d->w * d->cin, (long long)d->h * d->w * d->cin, p.bw, p.bh, p.bni,
                        CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    } else {
      ok = make_act_map(&tm_x.m[v], x, d->cin, d->w, d->h, d->n, d->cin, (long long)d->w * d->cin,
                        (long long)d->h * d->w * d->cin, p.bw, p.bh, p.bni, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    }
    if (!ok) return cgan_fail(ctx, CGAN_ERR_CUDA, "%s: cuTensorMapEncodeTiled(x) failed%s", "cgan_wgrad_tc");
  }
  for (int v = 0; v < 4; ++v) {
    bool ok;
    if (!d->upsample) {
      ok = make_act_map(&tm_dy.m[v], dy, d->cout, gw, gh, d->n, d->cout, (long long)d->ow * d->cout,
                        (long long)d->oh * d->ow * d->cout, p.bw, p.bh, p.bni, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    } else {
      int a = v >> 1, b = v & 1;
      ok = make_act_map(&tm_dy.m[v], dy + ((long long)a * d->ow + b) * d->cout, d->cout, d->w, d->h, d->n, 2ll * d->cout,
                        2ll * d->ow * d->cout, (long long)d->oh * d->ow * d->cout, p.bw, p.bh, p.bni,
                        CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    }
    if (!ok) return cgan_fail(ctx, CGAN_ERR_CUDA, "%s: cuTensorMapEncodeTiled(dy) failed%s", "cgan_wgrad_tc");
  }
  const size_t stage_bytes = (size_t)p.mt * WG_A_BYTES + (size_t)(p.bn / 32) * WG_BOX;
  p.stages = (int)(((two_ctas ? 110 : 220) * 1024) / stage_bytes);
  if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
  if (p.stages < 2) p.stages = 2;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.mt * p.bn) p.tmem_cols *= 2;
  size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    CGAN_CUDA(ctx, cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)tiles, (unsigned)splits);
  wgrad_tc_kernel<<<grid, WG_THREADS, smem, ctx->stream>>>(tm_x, tm_dy, p);
  CGAN_LAUNCHED(ctx);
  if (splits > 1) {
    wg_reduce_kernel<<<cdiv(wn, 256), 256, 0, ctx->stream>>>(dw, partial, wn, splits);
    CGAN_LAUNCHED(ctx);
  }
  return CGAN_OK;
}


// C[i][k1, k2] = sum_pixels A[i][pixel, k1] * B[i][pixel, k2]   (per-image A^T B; attention's d(phi), d(g))
int cgan_wgrad_tc_batched(cgan_ctx* ctx, const float* a, const float* b, float* c, int batch, int h, int w, int k1, int k2) {
  WgParams p;
  memset(&p, 0, sizeof(p));
  if (!box32(1, h, w, &p.bw, &p.bh, &p.bni) || p.bni != 1)
    return cgan_fail(ctx, CGAN_ERR_UNSUPPORTED, "%s: geometry%s", "cgan_wgrad_tc_batched");
  p.tiles_w = w / p.bw;
  p.tiles_h = h / p.bh;
  const int kb_per_image = h * w / WG_P;
  p.kblocks = kb_per_image * batch;
  p.kb_per_split = kb_per_image;            // "split" i = image i: its partial tile IS the result C[i]
  p.bn = pick_bn(k2);
  if (p.bn == 0) return cgan_fail(ctx, CGAN_ERR_UNSUPPORTED, "%s: columns%s", "cgan_wgrad_tc_batched");
  p.ci_tiles = (k1 + 127) / 128;
  p.co_tiles = (k2 + p.bn - 1) / p.bn;
  p.cin = k1; p.cout = k2; p.taps_total = 1;
  p.ntaps = 1;
  p.mt = 1;
  p.round_a = p.round_b = 1;
  p.partial = c;
  BMaps tm_x, tm_dy;
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_dy, 0, sizeof(tm_dy));
  for (int v = 0; v < 4; ++v) {
    if (!make_act_map(&tm_x.m[v], a, k1, w, h, batch, k1, (long long)w * k1, (long long)h * w * k1, p.bw, p.bh, p.bni,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_act_map(&tm_dy.m[v], b, k2, w, h, batch, k2, (long long)w * k2, (long long)h * w * k2, p.bw, p.bh, p.bni,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return cgan_fail(ctx, CGAN_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed%s", "cgan_wgrad_tc_batched");
  }
  const size_t stage_bytes = WG_A_BYTES + (size_t)(p.bn / 32) * WG_BOX;
  p.stages = (int)((110 * 1024) / stage_bytes);
  if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
  if (p.stages < 2) p.stages = 2;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.bn) p.tmem_cols *= 2;
  size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    CGAN_CUDA(ctx, cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (batch > 65535) return cgan_fail(ctx, CGAN_ERR_UNSUPPORTED, "%s: batch exceeds grid.y%s", "cgan_wgrad_tc_batched");
  dim3 grid((unsigned)(p.ci_tiles * p.co_tiles), (unsigned)batch);
  wgrad_tc_kernel<<<grid, WG_THREADS, smem, ctx->stream>>>(tm_x, tm_dy, p);
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}
