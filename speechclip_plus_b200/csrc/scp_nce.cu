// S3 -- masked in-batch InfoNCE (both directions) and the L2-normalise + pack prologue of the all-gather.
// Replaces MaskedContrastiveLoss.forward (avssl/module/losses.py:185-245) and its autograd, and the three feature
// normalisations of KWClip_GeneralTransformer.forward (avssl/model/kwClip.py:857, :905-907, :913-915).
//
// The N x N logit matrix is never written.  Both directions run in ONE launch of the row-streaming tcgen05 engine:
//   rows [0,Np)   : A_i  against all B_j  -> row log-sum-exp   (a2b denominator, losses.py:235)
//   rows [Np,2Np) : B_j  against all A_i  -> column log-sum-exp (b2a denominator, losses.py:239)
// Precision: every fp32 operand x is split as 64x = hi + lo (two fp16 numbers); the operands are laid out as
// [hi|hi|lo] and [hi|lo|hi] along K so that a single accumulator receives hi*hi + hi*lo + lo*hi, i.e. the product is
// exact to ~2^-21 -- the loss that drives training is effectively evaluated in fp32 on the tensor cores.
// Backward: the same sweep re-computes the logits, forms G = mask*e^S*(1/Z^r_i + 1/Z^c_j) - 2 I (scaled), writes it as
// a hi/lo fp16 pair, and a second engine launch computes dA = G B and dB = G^T A (K = N).
#include "scp_stream_gemm.cuh"

namespace scp {

using tc::GemmMaps;
using tc::Sched;
using tc::WorkInfo;

constexpr float kLog2eN = 1.4426950408889634f;
constexpr float kFeatScale = 64.0f;            // operands are stored as 64*x
constexpr float kProdScale = 4096.0f;          // (64*a)(64*b)
constexpr float kGScale = 4096.0f;             // G~ = 4096 * Ghat, |Ghat| <= 2
constexpr float kNegBigN = -1.0e30f;
constexpr int kNceBN = 128;

__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

// warp <-> sample row.  Writes the stacked split operands and the positive-pair dot product.
//   x3 (2*Np, 3D): rows [0,Np) = [Ahi Ahi Alo], rows [Np,2Np) = [Bhi Bhi Blo]
//   y3 (2*Np, 3D): rows [0,Np) = [Bhi Blo Bhi], rows [Np,2Np) = [Ahi Alo Ahi]
__global__ void nce_prep_kernel(const float* __restrict__ A, const float* __restrict__ B, int64_t N, int64_t Np, int D,
                                __half* __restrict__ x3, __half* __restrict__ y3, float* __restrict__ pos) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= Np) return;
  const int64_t K3 = 3 * (int64_t)D;
  __half* xa = x3 + i * K3;
  __half* xb = x3 + (Np + i) * K3;
  __half* ya = y3 + i * K3;
  __half* yb = y3 + (Np + i) * K3;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float a = i < N ? A[i * D + d] : 0.f;
    const float b = i < N ? B[i * D + d] : 0.f;
    dot = fmaf(a, b, dot);
    __half ah, al, bh, bl;
    split_f16(a * kFeatScale, ah, al);
    split_f16(b * kFeatScale, bh, bl);
    xa[d] = ah; xa[D + d] = ah; xa[2 * D + d] = al;
    xb[d] = bh; xb[D + d] = bh; xb[2 * D + d] = bl;
    ya[d] = bh; ya[D + d] = bl; ya[2 * D + d] = bh;
    yb[d] = ah; yb[D + d] = al; yb[2 * D + d] = ah;
  }
  dot = warp_sum(dot);
  if (lane == 0 && i < N) pos[i] = dot;
}

// transposed split operands for the backward GEMM (K = sample index):
//   t3 (2*D, 3*Np): rows [0,D) = [Bhi^T Blo^T Bhi^T], rows [D,2D) = [Ahi^T Alo^T Ahi^T]
// 32 x 32 tiles through shared memory so that both the reads (along d) and the writes (along i) are coalesced.
__global__ void nce_prep_t_kernel(const float* __restrict__ A, const float* __restrict__ B, int64_t N, int64_t Np, int D,
                                  __half* __restrict__ t3) {
  __shared__ float ta[32][33], tb[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = i0 + r;
    const int d = d0 + threadIdx.x;
    ta[r][threadIdx.x] = (i < N && d < D) ? A[i * D + d] : 0.f;
    tb[r][threadIdx.x] = (i < N && d < D) ? B[i * D + d] : 0.f;
  }
  __syncthreads();
  const int64_t ld = 3 * Np;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r;
    const int64_t i = i0 + threadIdx.x;
    if (d < D && i < Np) {
      __half ah, al, bh, bl;
      split_f16(ta[threadIdx.x][r] * kFeatScale, ah, al);
      split_f16(tb[threadIdx.x][r] * kFeatScale, bh, bl);
      __half* rb = t3 + (int64_t)d * ld;
      __half* ra = t3 + ((int64_t)D + d) * ld;
      rb[i] = bh; rb[Np + i] = bl; rb[2 * Np + i] = bh;
      ra[i] = ah; ra[Np + i] = al; ra[2 * Np + i] = ah;
    }
  }
}

struct NceCommon {
  const int64_t* ids;      // (N,) or null
  const float* log_scale;  // device scalar or null
  float fixed_scale;
  float margin;
  int dcl;
  int N;
  __device__ __forceinline__ float scale() const { return log_scale ? expf(__ldg(log_scale)) : fixed_scale; }
};

// mask of losses.py:202-216: different id, or the diagonal unless dcl
__device__ __forceinline__ bool nce_in_denominator(bool diag, bool same_id, int dcl) {
  return diag ? !dcl : !same_id;
}

// ---- forward sweep: online log-sum-exp per stacked row -----------------------------------------------------
struct NceFwdEpi {
  struct Params {
    NceCommon c;
    float* partials;  // (2*Lp, 2*n_groups, 2): running max, sum
    int Np, n_groups;
    int Lp, row_begin, n_local;  // the stacked X rows are the samples [row_begin, row_begin + n_local) in both directions
  };
  static constexpr int kSmemBytes = 0;
  const Params& p;
  int r, own, dir, group;
  bool own_valid;
  int64_t own_id;
  float k, run_max, sum;
  __device__ __forceinline__ NceFwdEpi(const Params& p_, const WorkInfo& w, const tc::EpiCtx& ctx)
      : p(p_), r(w.m_tile * tc::kTileM + ctx.row_in_tile), group(w.n_group * 2 + ctx.half) {
    dir = r >= p.Lp;
    const int loc = r - dir * p.Lp;
    own = p.row_begin + loc;
    own_valid = loc < p.n_local;
    own_id = (p.c.ids && own_valid) ? __ldg(p.c.ids + own) : (int64_t)own;
    k = p.c.scale() / kProdScale;
    run_max = kNegBigN;
    sum = 0.f;
  }
  __device__ __forceinline__ void tile_begin(int) {}
  __device__ __forceinline__ void tile_end(int) {}
  __device__ __forceinline__ void chunk(int col0, float (&v)[1][32]) {
    const int j0 = col0 - dir * p.Np;
    float l[32];
    float cmax = kNegBigN;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int j = j0 + i;
      const bool diag = j == own;
      const int64_t jid = (p.c.ids && j < p.c.N) ? __ldg(p.c.ids + j) : (int64_t)j;
      const bool inc = own_valid && j < p.c.N && nce_in_denominator(diag, !diag && jid == own_id, p.c.dcl);
      float x = v[0][i] * k;
      if (diag) x -= p.c.margin;
      l[i] = inc ? x : kNegBigN;
      cmax = fmaxf(cmax, l[i]);
    }
    if (cmax <= kNegBigN) return;  // nothing of this chunk is in the denominator
    if (cmax > run_max) {
      sum *= tc::fast_ex2((run_max - cmax) * kLog2eN);
      run_max = cmax;
    }
    const float shift = run_max * kLog2eN;
#pragma unroll
    for (int i = 0; i < 32; ++i) sum += tc::fast_ex2(fmaf(l[i], kLog2eN, -shift));
  }
  __device__ __forceinline__ void finish() {
    float2 o = make_float2(run_max, sum);
    *reinterpret_cast<float2*>(p.partials + ((int64_t)r * (2 * p.n_groups) + group) * 2) = o;
  }
};

// one block: combine partials, loss = mean over rows (losses.py:234-243)
__global__ void __launch_bounds__(1024)
nce_loss_kernel(const float* __restrict__ partials, int Np, int n_groups, const float* __restrict__ pos, NceCommon c,
                int a2b, int b2a, float* __restrict__ loss, float* __restrict__ lse_row, float* __restrict__ lse_col) {
  __shared__ float s_red[32];
  const float scale = c.scale();
  float acc = 0.f;
  for (int i = threadIdx.x; i < c.N; i += blockDim.x) {
    float lse[2];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      const float2* pp = reinterpret_cast<const float2*>(partials + ((int64_t)(dir * Np + i) * n_groups) * 2);
      float mx = kNegBigN;
      for (int g = 0; g < n_groups; ++g) mx = fmaxf(mx, pp[g].x);
      float s = 0.f;
      for (int g = 0; g < n_groups; ++g) s += pp[g].y * expf(pp[g].x - mx);
      lse[dir] = mx + logf(s);
    }
    lse_row[i] = lse[0];
    lse_col[i] = lse[1];
    const float p = pos[i] * scale - c.margin;
    if (a2b) acc += lse[0] - p;
    if (b2a) acc += lse[1] - p;
  }
  // block reduction
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) *loss = t / ((float)c.N * (float)(a2b + b2a));
  }
}

// ---- sharded forward (one process per GPU): each rank evaluates the denominators of ITS rows / columns only ----------
// stats (3, n_local): [0] = log sum_j mask e^{S_ij} (row i local), [1] = log sum_i mask e^{S_ij} (column j local),
// [2] = <A_i, B_i>.  The ranks all-gather this vector (12 bytes per sample) instead of recomputing the N x N problem.
// warp <-> (direction, local row): the lanes stride over the row's partial slots (a thread per row walked up to ~150 dependent
// loads serially: 11 us at N = 2048 for 256 local rows)
__global__ void __launch_bounds__(256)
nce_local_stats_kernel(const float* __restrict__ partials, int Lp, int n_groups, const float* __restrict__ pos,
                       int row_begin, int n_local, float* __restrict__ stats) {
  const int w = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= 2 * n_local) return;  // whole warp
  const int dir = w / n_local, l = w - dir * n_local;
  const float2* pp = reinterpret_cast<const float2*>(partials + ((int64_t)(dir * Lp + l) * n_groups) * 2);
  float mx = kNegBigN;
  for (int g = lane; g < n_groups; g += 32) mx = fmaxf(mx, pp[g].x);
  mx = warp_max(mx);
  float sm = 0.f;
  for (int g = lane; g < n_groups; g += 32) sm += pp[g].y * expf(pp[g].x - mx);
  sm = warp_sum(sm);
  if (lane == 0) {
    stats[dir * n_local + l] = mx + logf(sm);
    if (dir == 0) stats[2 * n_local + l] = pos[row_begin + l];
  }
}

// one block: the loss (losses.py:234-243) from the gathered (world, 3, n) statistics; also writes the contiguous
// (N,) denominators the backward reads.  Every rank evaluates the same N terms in the same order: identical loss values.
__global__ void __launch_bounds__(1024)
nce_loss_from_stats_kernel(const float* __restrict__ stats, int world, int n, float scale_fixed,
                           const float* __restrict__ log_scale, float margin, int a2b, int b2a, float* __restrict__ loss,
                           float* __restrict__ lse_row, float* __restrict__ lse_col) {
  __shared__ float s_red[32];
  const float scale = log_scale ? expf(__ldg(log_scale)) : scale_fixed;
  const int N = world * n;
  float acc = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int r = i / n, l = i - r * n;
    const float* st = stats + (int64_t)r * 3 * n;
    const float lr = st[l], lc = st[n + l], p = st[2 * n + l] * scale - margin;
    lse_row[i] = lr;
    lse_col[i] = lc;
    if (a2b) acc += lr - p;
    if (b2a) acc += lc - p;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) *loss = t / ((float)N * (float)(a2b + b2a));
  }
}

// ---- backward sweep: G~ rows (hi/lo fp16) for the local samples ---------------------------------------------
struct NceBwdEpi {
  struct Params {
    NceCommon c;
    const float* lse_row;  // (N,)
    const float* lse_col;  // (N,)
    __half* g3;            // (2*Lp, 3*Np): [hi | hi | lo]
    float* dscale_part;    // (Lp, n_groups) sum_j Ghat_ij * raw_ij   (direction 0 rows only)
    int Np, Lp, row_begin, n_local, n_groups;
    int a2b, b2a;
  };
  static constexpr int kSmemBytes = 0;
  const Params& p;
  int r, own, dir, group;
  bool own_valid;
  int64_t own_id;
  float k, own_lse_l2, own_a, col_a, dsum;
  const float* col_lse;
  __device__ __forceinline__ NceBwdEpi(const Params& p_, const WorkInfo& w, const tc::EpiCtx& ctx)
      : p(p_), r(w.m_tile * tc::kTileM + ctx.row_in_tile), group(w.n_group * 2 + ctx.half) {
    dir = r >= p.Lp;
    const int loc = r - dir * p.Lp;
    own_valid = loc < p.n_local;
    own = p.row_begin + loc;
    own_id = (p.c.ids && own_valid) ? __ldg(p.c.ids + own) : (int64_t)own;
    k = p.c.scale() / kProdScale;
    const float* own_arr = dir ? p.lse_col : p.lse_row;
    col_lse = dir ? p.lse_row : p.lse_col;
    own_a = dir ? (float)p.b2a : (float)p.a2b;
    col_a = dir ? (float)p.a2b : (float)p.b2a;
    own_lse_l2 = own_valid ? __ldg(own_arr + own) * kLog2eN : 0.f;
    dsum = 0.f;
  }
  __device__ __forceinline__ void tile_begin(int) {}
  __device__ __forceinline__ void tile_end(int) {}
  __device__ __forceinline__ void chunk(int col0, float (&v)[1][32]) {
    const int j0 = col0;  // backward Y is not stacked per direction: see launch (n_upper_off selects the half)
    const int jbase = j0 - dir * p.Np;
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i2 = 0; i2 < 16; ++i2) {
      float g2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 2 * i2 + h;
        const int j = jbase + i;
        const bool diag = j == own;
        const bool jv = j < p.c.N && own_valid;
        const int64_t jid = (p.c.ids && j < p.c.N) ? __ldg(p.c.ids + j) : (int64_t)j;
        const bool inc = jv && nce_in_denominator(diag, !diag && jid == own_id, p.c.dcl);
        const float raw = v[0][i] * k;
        const float x = (diag ? raw - p.c.margin : raw) * kLog2eN;
        const float cl = jv ? __ldg(col_lse + j) * kLog2eN : 0.f;
        float g = 0.f;
        if (inc) g = own_a * tc::fast_ex2(x - own_lse_l2) + col_a * tc::fast_ex2(x - cl);
        if (diag && jv) g -= (own_a + col_a);
        dsum = fmaf(g, raw, dsum);
        g2[h] = g * kGScale;
      }
      const __half2 h2 = __floats2half2_rn(g2[0], g2[1]);
      const float2 back = __half22float2(h2);
      const __half2 l2 = __floats2half2_rn(g2[0] - back.x, g2[1] - back.y);
      hi[i2] = *reinterpret_cast<const uint32_t*>(&h2);
      lo[i2] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    __half* row = p.g3 + (int64_t)r * 3 * p.Np + jbase;
    uint4* d0 = reinterpret_cast<uint4*>(row);
    uint4* d1 = reinterpret_cast<uint4*>(row + p.Np);
    uint4* d2 = reinterpret_cast<uint4*>(row + 2 * p.Np);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 h4 = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
      d0[q] = h4;
      d1[q] = h4;
      d2[q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
    }
  }
  __device__ __forceinline__ void finish() {
    if (!dir) p.dscale_part[(int64_t)r * (2 * p.n_groups) + group] = own_valid ? dsum : 0.f;
  }
};

template <int NX>
struct NceStoreEpi {
  struct Params {
    float* out;    // (k_splits, rows, ld)
    int64_t rows;
    int ld;
  };
  static constexpr int kSmemBytes = 0;
  const Params& p;
  int64_t row;
  int ks;
  __device__ __forceinline__ NceStoreEpi(const Params& p_, const WorkInfo& w, const tc::EpiCtx& ctx)
      : p(p_), row((int64_t)w.m_tile * tc::kTileM + ctx.row_in_tile), ks(w.k_split) {}
  __device__ __forceinline__ void tile_begin(int) {}
  __device__ __forceinline__ void tile_end(int) {}
  __device__ __forceinline__ void chunk(int col0, float (&v)[NX][32]) {
    float4* dst = reinterpret_cast<float4*>(p.out + ((int64_t)ks * p.rows + row) * p.ld + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[0][4 * i], v[0][4 * i + 1], v[0][4 * i + 2], v[0][4 * i + 3]);
  }
  __device__ __forceinline__ void finish() {}
};

// dA / dB = coef * sum_ks out ; d_log_scale = g * sum_i dscale_part / (N ndir)
__global__ void nce_bwd_finalize_kernel(const float* __restrict__ out, int k_splits, int64_t Lp, int D, int64_t n_local,
                                        NceCommon c, int ndir, const float* __restrict__ g_loss,
                                        const float* __restrict__ dscale_part, int n_groups, float* __restrict__ dA,
                                        float* __restrict__ dB, float* __restrict__ d_log_scale) {
  const float g = *g_loss / ((float)c.N * (float)ndir);
  const float coef = g * c.scale() / (kGScale * kFeatScale);
  const int64_t total = n_local * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / D;
    const int d = (int)(e - i * D);
    float sa = 0.f, sb = 0.f;
    for (int ks = 0; ks < k_splits; ++ks) {
      sa += out[((int64_t)ks * 2 * Lp + i) * (2 * D) + d];
      if (dB) sb += out[((int64_t)ks * 2 * Lp + Lp + i) * (2 * D) + D + d];
    }
    dA[e] = sa * coef;
    if (dB) dB[e] = sb * coef;
  }
  if (d_log_scale && blockIdx.x == 0 && threadIdx.x < 32) {
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n_local * n_groups; i += 32) s += dscale_part[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) *d_log_scale = s * g;
  }
}

// ---- L2 normalise + pack --------------------------------------------------------------------------------
struct FeatPtrs {
  const void* p[4];
};
template <typename T>
__global__ void l2norm_pack_kernel(FeatPtrs fp, int n_feats, int64_t n, int D, const int64_t* __restrict__ ids,
                                   float* __restrict__ packed, float* __restrict__ inv_norms) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w < (int64_t)n_feats * n) {
    const int f = (int)(w / n);
    const int64_t i = w - (int64_t)f * n;
    const T* src = reinterpret_cast<const T*>(fp.p[f]) + i * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float x = (float)src[d];
      ss = fmaf(x, x, ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / sqrtf(ss);  // no epsilon, as kwClip.py:857
    float* dst = packed + ((int64_t)f * n + i) * D;
    for (int d = lane; d < D; d += 32) dst[d] = (float)src[d] * inv;
    if (lane == 0 && inv_norms) inv_norms[(int64_t)f * n + i] = inv;
  }
  if (ids) {
    int64_t* id_dst = reinterpret_cast<int64_t*>(packed + (int64_t)n_feats * n * D);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) id_dst[t] = ids[t];
  }
}

// backward of f / ||f||: g_f = (g_n - <g_n, fhat> fhat) * inv_norm      warp <-> row
__global__ void l2norm_bwd_kernel(const float* __restrict__ g_n, const float* __restrict__ f_hat,
                                  const float* __restrict__ inv_norm, int64_t n, int D, float* __restrict__ g_f) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot = fmaf(g_n[i * D + d], f_hat[i * D + d], dot);
  dot = warp_sum(dot);
  const float inv = inv_norm[i];
  for (int d = lane; d < D; d += 32) g_f[i * D + d] = (g_n[i * D + d] - dot * f_hat[i * D + d]) * inv;
}

// ---- host ---------------------------------------------------------------------------------------------------
struct NceWs {
  __half* x3;
  __half* y3;
  __half* t3;
  __half* g3;
  float* pos;
  float* partials;
  float* dscale_part;
  float* out;
  size_t total;
  int n_groups, k_splits, bn_out;
};
static int nce_out_bn(int64_t D) { return D % 256 == 0 ? 256 : (D % 128 == 0 ? 128 : 64); }
static NceWs nce_ws(void* base, int64_t N, int64_t D, int64_t n_local) {
  const int64_t Np = round_up(N, tc::kTileM), Lp = round_up(n_local, tc::kTileM);
  NceWs w{};
  const int m_tiles = (int)(2 * Np / tc::kTileM);
  const int n_tiles = (int)(Np / kNceBN);
  w.n_groups = std::max(1, std::min(n_tiles, kNumSMs / m_tiles));
  w.bn_out = nce_out_bn(D);
  const int out_items = (int)(2 * Lp / tc::kTileM) * (int)(D / w.bn_out);
  const int k_chunks = (int)(3 * Np / tc::kChunkK);
  w.k_splits = std::max(1, std::min(k_chunks, kNumSMs / std::max(out_items, 1)));
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  w.x3 = static_cast<__half*>(take((size_t)2 * Np * 3 * D * 2));
  w.y3 = static_cast<__half*>(take((size_t)2 * Np * 3 * D * 2));
  w.t3 = static_cast<__half*>(take((size_t)2 * D * 3 * Np * 2));
  w.g3 = static_cast<__half*>(take((size_t)2 * Lp * 3 * Np * 2));
  w.pos = static_cast<float*>(take((size_t)Np * 4));
  w.partials = static_cast<float*>(take((size_t)2 * Np * w.n_groups * 2 * 8));
  w.dscale_part = static_cast<float*>(take((size_t)Lp * kNumSMs * 2 * 4));
  // split-K partials: k_splits * out_items <= 148 bounds the size independently of n_local
  w.out = static_cast<float*>(take(std::max((size_t)2 * Np * 2 * D * 4, (size_t)kNumSMs * tc::kTileM * 256 * 2 * 4)));
  w.total = off;
  return w;
}

static int check_nce_shape(int64_t N, int64_t D) {
  SCP_CHECK_ARG(N > 0 && D > 0, "nce: non-positive shape");
  if (D % 64 != 0) return fail(SCP_ERR_UNSUPPORTED, "nce: D must be a multiple of 64, got %lld", (long long)D);
  if (N > (1 << 20)) return fail(SCP_ERR_UNSUPPORTED, "nce: N too large");
  return SCP_OK;
}

static Sched nce_sweep_sched(int64_t rows_half_p, int64_t Np, int64_t D, int n_groups) {
  Sched sc{};
  sc.m_tiles = (int)(2 * rows_half_p / tc::kTileM);
  sc.m_half = (int)(rows_half_p / tc::kTileM);
  sc.n_tiles = (int)(Np / kNceBN);
  sc.n_upper_off = sc.n_tiles;
  sc.n_groups = n_groups;
  sc.k_chunks = (int)(3 * D / tc::kChunkK);
  sc.k_splits = 1;
  sc.x_upper_row_off = 0;
  return sc;
}

}  // namespace scp

using namespace scp;

extern "C" size_t scp_nce_workspace_bytes(int64_t N, int64_t D) { return nce_ws(nullptr, N, D, N).total; }

extern "C" int scp_nce_fwd(const float* A, const float* Bm, const int64_t* ids, int64_t N, int64_t D,
                           const float* log_scale, float fixed_scale, float margin, int dcl, int a2b, int b2a,
                           int prepare_bwd, float* loss, float* lse_row, float* lse_col, void* workspace,
                           size_t workspace_bytes, scp_stream_t stream) {
  int rc = check_device_arch();
  if (rc) return rc;
  rc = check_nce_shape(N, D);
  if (rc) return rc;
  SCP_CHECK_ARG(A && Bm && loss && lse_row && lse_col && workspace, "nce_fwd: null pointer");
  SCP_CHECK_ARG(a2b || b2a, "nce_fwd: a2b and b2a both off");
  const NceWs ws = nce_ws(workspace, N, D, N);
  if (workspace_bytes < ws.total) return fail(SCP_ERR_WORKSPACE, "nce_fwd: workspace %zu < %zu", workspace_bytes, ws.total);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t Np = round_up(N, tc::kTileM);
  NceCommon c{ids, log_scale, fixed_scale, margin, dcl, (int)N};

  nce_prep_kernel<<<(unsigned)ceil_div(Np, 8), 256, 0, s>>>(A, Bm, N, Np, (int)D, ws.x3, ws.y3, ws.pos);
  SCP_CUDA_LAUNCH_CHECK("nce_prep");
  if (prepare_bwd) {
    dim3 tb(32, 8), tg((unsigned)(Np / 32), (unsigned)ceil_div(D, 32));
    nce_prep_t_kernel<<<tg, tb, 0, s>>>(A, Bm, N, Np, (int)D, ws.t3);
    SCP_CUDA_LAUNCH_CHECK("nce_prep_t");
  }
  GemmMaps maps{};
  if ((rc = tc::make_tmap_f16(&maps.x[0], ws.x3, 2 * Np, 3 * D, 3 * D, tc::kTileM))) return rc;
  maps.x[1] = maps.x[0];
  if ((rc = tc::make_tmap_f16(&maps.y, ws.y3, 2 * Np, 3 * D, 3 * D, kNceBN))) return rc;
  const Sched sc = nce_sweep_sched(Np, Np, D, ws.n_groups);
  NceFwdEpi::Params ep{c, ws.partials, (int)Np, ws.n_groups, (int)Np, 0, (int)N};
  if ((rc = tc::launch_stream_gemm<kNceBN, 1, 6, NceFwdEpi>(maps, sc, ep, s, "nce_fwd_sweep"))) return rc;
  nce_loss_kernel<<<1, 1024, 0, s>>>(ws.partials, (int)Np, 2 * ws.n_groups, ws.pos, c, a2b, b2a, loss, lse_row, lse_col);
  SCP_CUDA_LAUNCH_CHECK("nce_loss");
  return SCP_OK;
}

extern "C" int scp_nce_fwd_local(const float* A, const float* Bm, const int64_t* ids, int64_t N, int64_t D,
                                 const float* log_scale, float fixed_scale, float margin, int dcl, int64_t row_begin,
                                 int64_t row_end, int prepare_bwd, float* stats_local, void* workspace,
                                 size_t workspace_bytes, scp_stream_t stream) {
  int rc = check_device_arch();
  if (rc) return rc;
  rc = check_nce_shape(N, D);
  if (rc) return rc;
  SCP_CHECK_ARG(A && Bm && stats_local && workspace, "nce_fwd_local: null pointer");
  SCP_CHECK_ARG(0 <= row_begin && row_begin < row_end && row_end <= N, "nce_fwd_local: bad local row range");
  const NceWs ws = nce_ws(workspace, N, D, N);
  if (workspace_bytes < ws.total) return fail(SCP_ERR_WORKSPACE, "nce_fwd_local: workspace %zu < %zu", workspace_bytes, ws.total);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t n_local = row_end - row_begin;
  const int64_t Np = round_up(N, tc::kTileM), Lp = round_up(n_local, tc::kTileM);
  NceCommon c{ids, log_scale, fixed_scale, margin, dcl, (int)N};
  nce_prep_kernel<<<(unsigned)ceil_div(Np, 8), 256, 0, s>>>(A, Bm, N, Np, (int)D, ws.x3, ws.y3, ws.pos);
  SCP_CUDA_LAUNCH_CHECK("nce_prep");
  if (prepare_bwd) {
    dim3 tb(32, 8), tg((unsigned)(Np / 32), (unsigned)ceil_div(D, 32));
    nce_prep_t_kernel<<<tg, tb, 0, s>>>(A, Bm, N, Np, (int)D, ws.t3);
    SCP_CUDA_LAUNCH_CHECK("nce_prep_t");
  }
  // X = the local rows of the stacked split operands, addressed in place (as in scp_nce_bwd)
  const int sweep_m_tiles = (int)(2 * Lp / tc::kTileM);
  const int sweep_groups = std::max(1, std::min((int)(Np / kNceBN), kNumSMs / sweep_m_tiles));
  GemmMaps maps{};
  if ((rc = tc::make_tmap_f16(&maps.x[0], ws.x3 + row_begin * 3 * D, 2 * Np - row_begin, 3 * D, 3 * D, tc::kTileM))) return rc;
  maps.x[1] = maps.x[0];
  if ((rc = tc::make_tmap_f16(&maps.y, ws.y3, 2 * Np, 3 * D, 3 * D, kNceBN))) return rc;
  Sched sc = nce_sweep_sched(Lp, Np, D, sweep_groups);
  sc.x_upper_row_off = (int)(Np - Lp);
  // the partials of a local sweep live at the start of the (larger) split-K buffer of the backward: the forward's own
  // partial buffer is sized for n_groups of the FULL sweep, a local sweep has more groups per row
  float* partials = ws.out;
  NceFwdEpi::Params ep{c, partials, (int)Np, sweep_groups, (int)Lp, (int)row_begin, (int)n_local};
  if ((rc = tc::launch_stream_gemm<kNceBN, 1, 6, NceFwdEpi>(maps, sc, ep, s, "nce_fwd_local_sweep"))) return rc;
  nce_local_stats_kernel<<<(unsigned)ceil_div(2 * n_local, 8), 256, 0, s>>>(partials, (int)Lp, 2 * sweep_groups, ws.pos,
                                                                           (int)row_begin, (int)n_local, stats_local);
  SCP_CUDA_LAUNCH_CHECK("nce_local_stats");
  return SCP_OK;
}

extern "C" int scp_nce_loss_from_stats(const float* stats_all, int world, int64_t n_local, const float* log_scale,
                                       float fixed_scale, float margin, int a2b, int b2a, float* loss, float* lse_row,
                                       float* lse_col, scp_stream_t stream) {
  SCP_CHECK_ARG(stats_all && loss && lse_row && lse_col, "nce_loss_from_stats: null pointer");
  SCP_CHECK_ARG(world >= 1 && n_local >= 1 && (int64_t)world * n_local <= (1 << 20), "nce_loss_from_stats: bad shape");
  SCP_CHECK_ARG(a2b || b2a, "nce_loss_from_stats: a2b and b2a both off");
  nce_loss_from_stats_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      stats_all, world, (int)n_local, fixed_scale, log_scale, margin, a2b, b2a, loss, lse_row, lse_col);
  SCP_CUDA_LAUNCH_CHECK("nce_loss_from_stats");
  return SCP_OK;
}

template <int BN>
static int launch_nce_out(const GemmMaps& maps, const Sched& sc, const NceStoreEpi<1>::Params& ep, cudaStream_t s) {
  constexpr int kStages = BN == 256 ? 4 : 6;
  return tc::launch_stream_gemm<BN, 1, kStages, NceStoreEpi<1>>(maps, sc, ep, s, "nce_gemm_out");
}

extern "C" int scp_nce_bwd(const float* A, const float* Bm, const int64_t* ids, int64_t N, int64_t D,
                           const float* log_scale, float fixed_scale, float margin, int dcl, int a2b, int b2a,
                           const float* lse_row, const float* lse_col, const float* g_loss, int64_t row_begin,
                           int64_t row_end, int fwd_state_valid, float* dA, float* dB, float* d_log_scale,
                           void* workspace, size_t workspace_bytes, scp_stream_t stream) {
  int rc = check_device_arch();
  if (rc) return rc;
  rc = check_nce_shape(N, D);
  if (rc) return rc;
  SCP_CHECK_ARG(A && Bm && lse_row && lse_col && g_loss && dA && workspace, "nce_bwd: null pointer");
  SCP_CHECK_ARG(0 <= row_begin && row_begin < row_end && row_end <= N, "nce_bwd: bad local row range");
  const int64_t n_local = row_end - row_begin;
  const NceWs ws = nce_ws(workspace, N, D, N);  // sized for the worst case n_local = N
  if (workspace_bytes < ws.total) return fail(SCP_ERR_WORKSPACE, "nce_bwd: workspace %zu < %zu", workspace_bytes, ws.total);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t Np = round_up(N, tc::kTileM), Lp = round_up(n_local, tc::kTileM);
  NceCommon c{ids, log_scale, fixed_scale, margin, dcl, (int)N};

  if (!fwd_state_valid) {  // the split operands of the forward call are not in `workspace`: rebuild them
    nce_prep_kernel<<<(unsigned)ceil_div(Np, 8), 256, 0, s>>>(A, Bm, N, Np, (int)D, ws.x3, ws.y3, ws.pos);
    SCP_CUDA_LAUNCH_CHECK("nce_prep");
    dim3 tb(32, 8), tg((unsigned)(Np / 32), (unsigned)ceil_div(D, 32));
    nce_prep_t_kernel<<<tg, tb, 0, s>>>(A, Bm, N, Np, (int)D, ws.t3);
    SCP_CUDA_LAUNCH_CHECK("nce_prep_t");
  }
  const int sweep_m_tiles = (int)(2 * Lp / tc::kTileM);
  const int sweep_groups = std::max(1, std::min((int)(Np / kNceBN), kNumSMs / sweep_m_tiles));
  // ---- sweep: G~ for the local rows.  X = the local rows of the stacked split operands, addressed in place:
  //      lower-half tiles start at row_begin (tensor-map base), upper-half tiles Np - Lp rows further.
  {
    GemmMaps maps{};
    if ((rc = tc::make_tmap_f16(&maps.x[0], ws.x3 + row_begin * 3 * D, 2 * Np - row_begin, 3 * D, 3 * D, tc::kTileM)))
      return rc;
    maps.x[1] = maps.x[0];
    if ((rc = tc::make_tmap_f16(&maps.y, ws.y3, 2 * Np, 3 * D, 3 * D, kNceBN))) return rc;
    Sched sc = nce_sweep_sched(Lp, Np, D, sweep_groups);
    sc.x_upper_row_off = (int)(Np - Lp);
    NceBwdEpi::Params ep{};
    ep.c = c;
    ep.lse_row = lse_row;
    ep.lse_col = lse_col;
    ep.g3 = ws.g3;
    ep.dscale_part = ws.dscale_part;
    ep.Np = (int)Np; ep.Lp = (int)Lp; ep.row_begin = (int)row_begin; ep.n_local = (int)n_local;
    ep.n_groups = sweep_groups;
    ep.a2b = a2b; ep.b2a = b2a;
    if ((rc = tc::launch_stream_gemm<kNceBN, 1, 6, NceBwdEpi>(maps, sc, ep, s, "nce_bwd_sweep"))) return rc;
  }
  // ---- dA = G B, dB = G^T A     (K = 3*Np)
  const int bn = nce_out_bn(D);
  const int out_items = (int)(2 * Lp / tc::kTileM) * (int)(D / bn);
  const int k_chunks = (int)(3 * Np / tc::kChunkK);
  const int k_splits = std::max(1, std::min(k_chunks, kNumSMs / std::max(out_items, 1)));
  {
    GemmMaps maps{};
    if ((rc = tc::make_tmap_f16(&maps.x[0], ws.g3, 2 * Lp, 3 * Np, 3 * Np, tc::kTileM))) return rc;
    maps.x[1] = maps.x[0];
    if ((rc = tc::make_tmap_f16(&maps.y, ws.t3, 2 * D, 3 * Np, 3 * Np, bn))) return rc;
    Sched sc{};
    sc.m_tiles = (int)(2 * Lp / tc::kTileM);
    sc.m_half = (int)(Lp / tc::kTileM);
    sc.n_tiles = (int)(D / bn);
    sc.n_upper_off = sc.n_tiles;
    sc.n_groups = sc.n_tiles;
    sc.k_chunks = k_chunks;
    sc.k_splits = k_splits;
    sc.x_upper_row_off = 0;
    NceStoreEpi<1>::Params ep{ws.out, 2 * Lp, (int)(2 * D)};
    if (bn == 256) rc = launch_nce_out<256>(maps, sc, ep, s);
    else if (bn == 128) rc = launch_nce_out<128>(maps, sc, ep, s);
    else rc = launch_nce_out<64>(maps, sc, ep, s);
    if (rc) return rc;
  }
  const int64_t total = n_local * D;
  nce_bwd_finalize_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 1024), 256, 0, s>>>(
      ws.out, k_splits, Lp, (int)D, n_local, c, a2b + b2a, g_loss, ws.dscale_part, 2 * sweep_groups, dA, dB, d_log_scale);
  SCP_CUDA_LAUNCH_CHECK("nce_bwd_finalize");
  return SCP_OK;
}

extern "C" size_t scp_pack_bytes(int n_feats, int64_t n, int64_t D) {
  return (size_t)n_feats * n * D * 4 + (size_t)n * 8;
}

extern "C" int scp_l2norm_pack(const void* const* feats, int n_feats, int64_t n, int64_t D, int dtype_in,
                               const int64_t* ids, void* packed_out, float* inv_norms, scp_stream_t stream) {
  SCP_CHECK_ARG(feats && packed_out, "l2norm_pack: null pointer");
  SCP_CHECK_ARG(n_feats >= 1 && n_feats <= 4 && n > 0 && D > 0, "l2norm_pack: bad shape");
  FeatPtrs fp{};
  for (int f = 0; f < n_feats; ++f) {
    SCP_CHECK_ARG(feats[f], "l2norm_pack: feats[%d] null", f);
    fp.p[f] = feats[f];
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t warps = (int64_t)n_feats * n;
  const unsigned blocks = (unsigned)std::max<int64_t>(ceil_div(warps, 8), ceil_div(n, 256));
  float* packed = reinterpret_cast<float*>(packed_out);
  if (dtype_in == SCP_F32) l2norm_pack_kernel<float><<<blocks, 256, 0, s>>>(fp, n_feats, n, (int)D, ids, packed, inv_norms);
  else if (dtype_in == SCP_F16) l2norm_pack_kernel<__half><<<blocks, 256, 0, s>>>(fp, n_feats, n, (int)D, ids, packed, inv_norms);
  else if (dtype_in == SCP_BF16) l2norm_pack_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(fp, n_feats, n, (int)D, ids, packed, inv_norms);
  else return fail(SCP_ERR_UNSUPPORTED, "l2norm_pack: dtype %d", dtype_in);
  SCP_CUDA_LAUNCH_CHECK("l2norm_pack");
  return SCP_OK;
}

extern "C" int scp_l2norm_bwd(const float* g_n, const float* f_hat, const float* inv_norm, int64_t n, int64_t D,
                              float* g_f, scp_stream_t stream) {
  SCP_CHECK_ARG(g_n && f_hat && inv_norm && g_f && n > 0 && D > 0, "l2norm_bwd: bad argument");
  l2norm_bwd_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g_n, f_hat, inv_norm,
                                                                                                 n, (int)D, g_f);
  SCP_CUDA_LAUNCH_CHECK("l2norm_bwd");
  return SCP_OK;
}
