// S2 -- keyword vector quantiser (cosine vs CLIP token table, arg-max / softmax / straight-through, fwd + bwd).
// Replaces GeneralBranch.get_keyword_cosine_score + SimpleVectorQuantizer.forward + the lookup matmul of the
// reference (avssl/model/kw_branches.py:158-197, avssl/module/speechclip_c_modules/my_vector_quantizer.py:64-165).
//
// Forward pipeline (the (M,V) fp32 logit matrix is never written to memory):
//   vq_prep_kw          kw fp32 -> unit rows in fp16 (A operand) + 1/||kw||
//   sweep 1 (tcgen05)   S = khat * Ehat^T tile by tile; per row: running softmax statistics at temperature 1 and tau,
//                       the maximum of every 32-column chunk (fp16-product precision) and ONE (M,V) fp16 matrix:
//                       e^c of every logit (scp_vq_fwd: workspace scratch for the column sums) or the soft-max
//                       numerators P'' = exp((c-1)/tau + 10) (scp_vq_fwd_save: the caller's buffer, kept for the backward)
//   vq_select           per row: chunks whose maximum is within the fp16 error bound of the row maximum are re-scored
//                       EXACTLY (fp64 accumulation over the fp32 table) -> arg-max is bit-exact w.r.t. an exact cosine,
//                       first index wins ties; combines the split statistics; gathers keywords = E[idx]; code histogram
//   vq_colsum           avg_probs[v] = mean_m e^c[m,v] / Z_m from that matrix (one pass; HBM-bound from e^c, SFU-bound from
//                       P'').  SCP_VQ_COLSUM=0 selects the older second tensor-core sweep instead (transposed product
//                       Ehat * khat^T with thread-serial column sums, no scratch)
//   vq_metrics          code_perplexity, prob_perplexity, diversity_loss, ent_per_t
// Backward, saved numerators (scp_vq_bwd_saved; default of the Python wrapper for a fixed / scheduled temperature):
//   vq_bwd_prep         g_keywords -> unit rows fp16 + scale + centring constant
//   sweep T (tcgen05)   ONE accumulator T = ghat Ehat^T (resident ghat); Q'' = P'' (T r_v - s0) from the forward's P'';
//                       writes fp16 Q'' and the row sums of Q'', P''
//   gemm_out (tcgen05)  U = Q'' * Ehat, W = P'' * Ehat  (K = V, split-K; P'' read in place from the forward's buffer)
//   vq_bwd_finalize     g_khat = (U - s W)/(tau sum P''), projection through the normalisation
// Backward, recompute (scp_vq_bwd; learnable temperature, tau < 0.1, single-tile problems):
//   sweep 3 (tcgen05)   two accumulators sharing Ehat tiles: S1 = khat Ehat^T, S2 = ghat Ehat^T;
//                       P = softmax_tau row, Q = P * (T - s0); writes fp16 P~, Q~ and the row sums
//   gemm_out, vq_bwd_finalize as above (+ optional d/dtau)
// Backward, opt-in low-memory mode (SCP_VQ_BWD_PIPE=1, D = 128 / 256 / 512; scp_vq_pipe.cuh): ONE producer/consumer launch --
//   producer CTA pairs compute c and T with one N = 256 MMA against a resident [khat | ghat] tile and emit P~, Q~ tiles into
//   a small ring that stays in L2; consumer CTA pairs accumulate U, W in TMEM.  Slower (487 vs 385 us), 8x less workspace.
#include <cfloat>

#include "scp_stream_gemm.cuh"

namespace scp {

using tc::GemmMaps;
using tc::Sched;
using tc::WorkInfo;

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNegBig = -1.0e30f;          // stands in for -inf (keeps 0 * x finite)
constexpr float kRescueMargin = 2.2e-3f;     // 2 x (2^-10 fp16 product bound) + slack, see DESIGN.md
constexpr float kPScale = 16384.0f;          // P~ = P * 2^14 keeps softmax rows in the fp16 normal range
constexpr int kVqBN = 256;                   // sweep 1 column tile
constexpr int kVqPad = 256;                  // vocabulary padding

struct MaskedCols {
  int n;
  int col[SCP_MAX_MASKED];
};
__device__ __forceinline__ bool is_masked(const MaskedCols& mc, int c) {
  bool m = false;
#pragma unroll
  for (int i = 0; i < SCP_MAX_MASKED; ++i) m |= (i < mc.n && mc.col[i] == c);
  return m;
}
// bit i set <=> column col0 + i is masked or beyond the vocabulary (rare path: kept small, not unrolled over columns)
__device__ __forceinline__ uint32_t chunk_mask_bits(const MaskedCols& mc, int col0, int V) {
  uint32_t bits = col0 + 32 > V ? (col0 >= V ? 0xffffffffu : ~((1u << (V - col0)) - 1u)) : 0u;
#pragma unroll
  for (int i = 0; i < SCP_MAX_MASKED; ++i)
    if (i < mc.n && (mc.col[i] >> 5) == (col0 >> 5)) bits |= 1u << (mc.col[i] & 31);
  return bits;
}
__device__ __forceinline__ bool chunk_has_mask(const MaskedCols& mc, int col0) {
  bool m = false;
#pragma unroll
  for (int i = 0; i < SCP_MAX_MASKED; ++i) m |= (i < mc.n && (mc.col[i] >> 5) == (col0 >> 5));
  return m;
}

// =====================================================================================================================
// preparation kernels
// =====================================================================================================================
// warp <-> table row: unit-normalise into fp16, record the norm.  rows >= V are zero padding.
__global__ void vq_table_normalize_kernel(const float* __restrict__ table, int64_t V, int64_t Vp, int D,
                                          __half* __restrict__ hat, float* __restrict__ norm,
                                          unsigned int* __restrict__ norm_max_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= Vp) return;
  __half* dst = hat + v * D;
  if (v >= V) {
    for (int d = lane; d < D; d += 32) dst[d] = __float2half(0.f);
    if (lane == 0) norm[v] = 1.f;
    return;
  }
  const float* src = table + v * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float x = src[d];
    ss = fmaf(x, x, ss);
  }
  ss = warp_sum(ss);
  const float n = fmaxf(sqrtf(ss), 1e-8f);  // F.cosine_similarity eps
  const float inv = 1.0f / n;
  for (int d = lane; d < D; d += 32) dst[d] = __float2half_rn(src[d] * inv);
  if (lane == 0) {
    norm[v] = n;
    atomicMax(norm_max_bits, __float_as_uint(n));  // positive floats order like their bit patterns
  }
}

// (R, C) fp16 -> (C, ldo) fp16 transpose through a padded smem tile
__global__ void transpose_f16_kernel(const __half* __restrict__ in, int64_t R, int C, __half* __restrict__ out,
                                     int64_t ldo) {
  __shared__ __half tile[32][34];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[r * C + c] : __float2half(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t r = r0 + threadIdx.x;
    if (c < C && r < ldo) out[(int64_t)c * ldo + r] = tile[threadIdx.x][i];
  }
}

// column means of the fp32 table: grid (ceil(D/32), row_splits), atomics into a zeroed buffer
__global__ void vq_table_mean_kernel(const float* __restrict__ table, int64_t V, int D, float* __restrict__ mean) {
  const int d = blockIdx.x * 32 + threadIdx.x;
  if (d >= D) return;
  float s = 0.f;
  for (int64_t v = (int64_t)blockIdx.y * blockDim.y + threadIdx.y; v < V; v += (int64_t)gridDim.y * blockDim.y)
    s += table[v * D + d];
  atomicAdd(&mean[d], s / (float)V);
}

// warp <-> keyword row: unit-normalise into fp16 (zero rows for m >= M), 1/max(||kw||,1e-8) -> row_stats[m][3].
// Also clears the code histogram and presets the padding entries of the sweep-2 normaliser vector (two memsets /
// helper launches folded into this one: the forward is a chain of short kernels and every launch costs ~3 us).
__global__ void vq_prep_kw_kernel(const float* __restrict__ kw, int64_t M, int64_t Mp, int D,
                                  __half* __restrict__ kw_hat, float* __restrict__ row_stats,
                                  float* __restrict__ code_hist, int64_t Vp, float* __restrict__ lse1_l2, int64_t Mp2,
                                  unsigned int* __restrict__ ticket) {
  const int lane = threadIdx.x & 31;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  if (gtid == 0) *ticket = 0u;  // the metrics kernel's last-block ticket
  for (int64_t v = gtid; v < Vp; v += gsz) code_hist[v] = 0.f;
  for (int64_t i = M + gtid; i < Mp2; i += gsz) lse1_l2[i] = -1.0e30f;  // padding columns contribute exp2(-1e30) = 0
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= Mp) return;
  __half* dst = kw_hat + m * D;
  if (m >= M) {
    for (int d = lane; d < D; d += 32) dst[d] = __float2half(0.f);
    return;
  }
  const float* src = kw + m * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float x = src[d];
    ss = fmaf(x, x, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
  for (int d = lane; d < D; d += 32) dst[d] = __float2half_rn(src[d] * inv);
  if (lane == 0) row_stats[m * 4 + 3] = inv;
}

// =====================================================================================================================
// sweep 1: per-row statistics of S = khat * Ehat^T
// =====================================================================================================================
// x^N by repeated squaring on a packed pair, N known at compile time (no branches in the epilogue's inner loop)
template <int N>
__device__ __forceinline__ tc::f32x2 ipow2(tc::f32x2 x) {
  if constexpr (N == 1) {
    return x;
  } else if constexpr (N % 2 == 0) {
    const tc::f32x2 t = ipow2<N / 2>(x);
    return tc::mul2(t, t);
  } else {
    return tc::mul2(x, ipow2<N - 1>(x));
  }
}

// SAVE: the fp16 scratch holds the soft-max numerators at temperature tau INSTEAD of e^c,
//     P''[m,v] = exp((c - 1)/tau + 10)      (= e^{c/tau} at tau = 0.1; in [e^-10, e^10] for every tau: fp16 range),
// and is owned by the caller: the column sums and the arg-max filter recover e^c = P''^tau e^{1 - 10 tau} from it
// (relative error tau x the fp16 rounding), and the backward pass (scp_vq_bwd_saved) needs neither the k . E^T product
// nor an exponential -- SweepTEpi reads P'' back and the output GEMM consumes the same buffer.  One (M,V) fp16 matrix is
// written per forward either way.
template <bool SAVE>
struct Sweep1EpiT {
  struct Params {
    float* chunk_max;   // (Mp, n_chunks): maximum of every 32-column chunk (scanned by the arg-max kernel)
    float* group_max;   // (Mp, n_chunks, 4): maximum of every 8-column group (read for candidate chunks only)
    float* partials;    // (Mp, n_slots, 4): sum e^c, sum c e^c, running max, sum e^{(c-max)/tau};  n_slots = parts*n_groups
    const float* tau;   // device scalar
    __half* e16;        // nullable (Mp, ldE) fp16: e^c of every logit (SAVE: P'', zeros for masked / padding columns),
                        // consumed by vq_colsum_kernel (avg_probs), the arg-max filter and, SAVE, the backward pass
    int64_t ldE;
    int n_chunks;
    int n_groups;
    int V;
    int dbg;            // timing ablations (env SCP_VQ_S1_DBG, results invalid): 1 = no chunk/group maxima stores,
                        // 2 = no e^c / P'' stores (pow10 path)
    MaskedCols mc;
  };
  static constexpr int kSmemBytes = 0;
  const Params& p;
  int64_t row;
  int slot, n_slots;
  bool pow10;   // 1/tau == 10 (every shipped recipe: "fixed=0.1"): e^{c/tau} = (e^c)^10 by repeated squaring
  float k_tau;  // log2(e)/tau
  // running sums as packed pairs (even / odd columns), two independent sets to shorten the dependency chains:
  // sum e^c, sum c e^c, sum e^{(c - run_max)/tau}
  tc::f32x2 acc_e[2], acc_ce[2], acc_et[2];
  float run_max;

  __device__ __forceinline__ Sweep1EpiT(const Params& p_, const WorkInfo& w, const tc::EpiCtx& ctx)
      : p(p_), row((int64_t)w.m_tile * tc::kTileM + ctx.row_in_tile), slot(w.n_group * ctx.parts + ctx.half),
        n_slots(ctx.parts * p_.n_groups) {
    const float inv_tau = 1.0f / __ldg(p.tau);
    k_tau = kLog2e * inv_tau;
    pow10 = fabsf(inv_tau - 10.0f) <= 1e-4f;
    const tc::f32x2 z = tc::pack2(0.f, 0.f);
    acc_e[0] = acc_e[1] = acc_ce[0] = acc_ce[1] = acc_et[0] = acc_et[1] = z;
    run_max = kNegBig;
  }
  __device__ __forceinline__ void tile_begin(int) {}
  __device__ __forceinline__ void tile_end(int) {}
  template <int N>
  __device__ __forceinline__ void pow_chunk(const float (&c)[32], uint32_t (&h)[16]) {
    const tc::f32x2 kl = tc::pack2(kLog2e, kLog2e);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const int a = (i >> 1) & 1;
      const tc::f32x2 cc = tc::pack2(c[i], c[i + 1]);
      const tc::f32x2 e1 = tc::ex2_2(tc::mul2(cc, kl));
      acc_e[a] = tc::add2(acc_e[a], e1);
      acc_ce[a] = tc::fma2(cc, e1, acc_ce[a]);
      const tc::f32x2 et = ipow2<N>(e1);
      acc_et[a] = tc::add2(acc_et[a], et);
      h[i >> 1] = tc::cvt_f16x2(SAVE ? et : e1);
    }
  }
  // e^c of the chunk's 32 columns as fp16: 64 contiguous bytes of this thread's row, two full-sector stores.  e^c lies in
  // [1/e, e], so fp16 keeps 11 significant bits (relative error <= 4.9e-4 per element, unbiased; avg_probs averages M of
  // them).  Masked / out-of-range columns hold 0 or are never written: vq_colsum_kernel selects them away by index.
  __device__ __forceinline__ void store_e(int col0, const uint32_t (&h)[16]) {
    __half* dst = p.e16 + row * p.ldE + col0;
    tc::stg256(dst, h);
    tc::stg256(dst + 16, h + 8);
  }
  __device__ __forceinline__ void chunk(int col0, float (&v)[1][32]) {
    float(&c)[32] = v[0];
    if (col0 + 32 > p.V || chunk_has_mask(p.mc, col0)) {
      const uint32_t bits = chunk_mask_bits(p.mc, col0, p.V);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if ((bits >> i) & 1u) c[i] = kNegBig;
    }
    // maxima of the four 8-column groups (3-input max, FMNMX3: 4 instructions per group), one 16-byte store; the
    // exact arg-max kernel re-scores only the groups that can hide the true maximum
    float gm[4];
#pragma unroll
    for (int g = 0; g < 4; ++g)
      gm[g] = tc::fmax3(tc::fmax3(c[8 * g], c[8 * g + 1], c[8 * g + 2]), tc::fmax3(c[8 * g + 3], c[8 * g + 4], c[8 * g + 5]),
                        fmaxf(c[8 * g + 6], c[8 * g + 7]));
    const float cmax = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
    if (!(p.dbg & 1)) {
      p.chunk_max[row * p.n_chunks + (col0 >> 5)] = cmax;
      // with the e^c scratch the arg-max kernel filters the groups of a candidate chunk from those 64 bytes instead: the
      // 16-byte group store and the two 32-byte e^c stores together saturated the store path (sweep 1: 92 -> 123 us)
      if (!p.e16)
        *reinterpret_cast<float4*>(p.group_max + (row * p.n_chunks + (col0 >> 5)) * 4) = make_float4(gm[0], gm[1], gm[2], gm[3]);
    }
    if (cmax <= kNegBig) {  // fully masked / padding chunk: contributes nothing (and has no finite maximum)
      if (SAVE && p.e16) {  // the output GEMM of the backward pass reads every column of the buffer
        const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        store_e(col0, z);
      }
      return;
    }
    if (pow10) {
      // |c| <= 1 and 1/tau = 10: e^{c/tau} = (e^c)^10 cannot overflow, so no running maximum is needed and the
      // second exponential becomes four packed multiplies -- ONE SFU op per logit instead of two.  Only this one
      // specialisation is compiled: every extra variant is another unrolled copy of the loop in the instruction cache.
      run_max = 0.f;  // the partial sum is relative to a shift of 0
      uint32_t h[16];
      pow_chunk<10>(c, h);
      if (p.e16 && !(p.dbg & 2)) store_e(col0, h);
      return;
    }
    if (cmax > run_max) {  // rescale the temperature-tau sum to the new running maximum
      const float f = tc::fast_ex2((run_max - cmax) * k_tau);
      const tc::f32x2 ff = tc::pack2(f, f);
      acc_et[0] = tc::mul2(acc_et[0], ff);
      acc_et[1] = tc::mul2(acc_et[1], ff);
      run_max = cmax;
    }
    const float shift = -run_max * k_tau;
    const tc::f32x2 kl = tc::pack2(kLog2e, kLog2e), kt = tc::pack2(k_tau, k_tau), sh = tc::pack2(shift, shift);
    const float shift_p = 10.0f * kLog2e - k_tau;  // P'' = 2^(c k_tau + shift_p) = exp((c - 1)/tau + 10)
    const tc::f32x2 shp = tc::pack2(shift_p, shift_p);
    uint32_t h[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const int a = (i >> 1) & 1;
      const tc::f32x2 cc = tc::pack2(c[i], c[i + 1]);
      const tc::f32x2 e1 = tc::ex2_2(tc::mul2(cc, kl));  // |c| <= 1: no shift needed at temperature 1
      acc_e[a] = tc::add2(acc_e[a], e1);
      acc_ce[a] = tc::fma2(cc, e1, acc_ce[a]);
      acc_et[a] = tc::add2(acc_et[a], tc::ex2_2(tc::fma2(cc, kt, sh)));
      h[i >> 1] = tc::cvt_f16x2(SAVE ? tc::ex2_2(tc::fma2(cc, kt, shp)) : e1);
    }
    if (p.e16) store_e(col0, h);
  }
  __device__ __forceinline__ void finish() {
    float4 o = make_float4(tc::hsum2(tc::add2(acc_e[0], acc_e[1])), tc::hsum2(tc::add2(acc_ce[0], acc_ce[1])), run_max,
                           tc::hsum2(tc::add2(acc_et[0], acc_et[1])));
    *reinterpret_cast<float4*>(p.partials + (row * n_slots + slot) * 4) = o;
  }
};
using Sweep1Epi = Sweep1EpiT<false>;
using Sweep1SaveEpi = Sweep1EpiT<true>;

// =====================================================================================================================
// exact arg-max + statistics combine + keyword gather      (block of 128 threads <-> row)
// =====================================================================================================================
struct Best {
  double val;
  int idx;
};
__device__ __forceinline__ Best better(const Best& a, const Best& b) {
  // larger value wins; equal values: smaller index (torch.max returns the first maximum)
  if (b.idx < 0) return a;
  if (a.idx < 0) return b;
  if (b.val > a.val || (b.val == a.val && b.idx < a.idx)) return b;
  return a;
}


// Exact re-scoring of one 8-column group by a warp: the 32 lanes split the D axis (coalesced 16-byte loads, the loads
// of all eight columns in flight together).
// Three precision levels keep both HBM and the fp64 pipe almost idle:
//   0. fp16-operand scores of the 8 columns from the UNIT fp16 table (the rows sweep 1 has just streamed: L2 hits, half
//      the bytes of the fp32 table).  They carry the same <= 2^-10 bound as the tensor-core logits, so only columns
//      within kRescueMargin of the row maximum (normally exactly one per row) can be the exact arg-max;
//   1. fp32 scores of those columns from the fp32 table (16 products per lane + butterfly: error <= ~1.3e-6 * ||kw||);
//   2. only the columns within 4e-6 * ||kw|| of the best fp32 score of the group are evaluated in fp64 (exact products
//      of fp32 numbers, fp64 accumulation) -- the group's true maximum is always among them.
// Before level 0 existed every candidate chunk cost 32 fp32 rows (64 KB) of mostly-cold HBM reads: 200 MB per call.
// kreg / khreg hold this lane's slice of the keyword row (fp32) and of its unit fp16 copy.
template <int NV>  // float4 vectors per lane: D <= 128*NV
__device__ __forceinline__ Best rescore_chunk(const float* __restrict__ table, const float* __restrict__ table_norm,
                                              const __half* __restrict__ table_hat, int V, int D, int group8,
                                              const float (&kreg)[NV][4], const uint4 (&khreg)[(NV + 1) / 2],
                                              float margin2, float thr16, const MaskedCols& mc, int lane) {
  Best best{0.0, -1};
  const int nvec = D >> 2;
  constexpr int NH = (NV + 1) / 2;  // uint4 (8 halfs) per lane
  const int nvech = D >> 3;
  // ---- level 0: fp16-operand scores, all 8 columns' loads in flight together
  uint4 h[8][NH];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int v = group8 * 8 + c;  // < Vp: padding rows of the unit table are zero
    const uint4* e = reinterpret_cast<const uint4*>(table_hat + (int64_t)v * D);
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      const int q = lane + 32 * j;
      h[c][j] = q < nvech ? __ldg(e + q) : make_uint4(0, 0, 0, 0);
    }
  }
  unsigned cand = 0;  // bit c: column c of this warp survives level 0 (warp-uniform)
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int v = group8 * 8 + c;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      const uint32_t a[4] = {h[c][j].x, h[c][j].y, h[c][j].z, h[c][j].w};
      const uint32_t b[4] = {khreg[j].x, khreg[j].y, khreg[j].z, khreg[j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&a[i]));
        const float2 y = __half22float2(*reinterpret_cast<const __half2*>(&b[i]));
        dot = fmaf(x.x, y.x, dot);
        dot = fmaf(x.y, y.y, dot);
      }
    }
    dot = warp_sum(dot);
    if (v < V && !is_masked(mc, v) && dot >= thr16) cand |= 1u << c;
  }
  if (cand == 0) return best;
  // ---- level 1: fp32 scores of the surviving columns
  // (levels 1 and 2 visit one or two columns: rolled loops keep the kernel small enough for the instruction cache)
  float sc[8];
  float gmax = -INFINITY;
#pragma unroll
  for (int c = 0; c < 8; ++c) sc[c] = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    if (cand & (1u << c)) {  // warp-uniform
      const int v = group8 * 8 + c;
      const float4* e = reinterpret_cast<const float4*>(table + (int64_t)v * D);
      float4 x[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int q = lane + 32 * j;
        x[j] = q < nvec ? __ldg(e + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        dot = fmaf(kreg[j][0], x[j].x, dot);
        dot = fmaf(kreg[j][1], x[j].y, dot);
        dot = fmaf(kreg[j][2], x[j].z, dot);
        dot = fmaf(kreg[j][3], x[j].w, dot);
      }
      dot = warp_sum(dot);
      const float sv = dot / __ldg(table_norm + v);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i == c) sc[i] = sv;
      gmax = fmaxf(gmax, sv);
    }
  }
  // ---- level 2: fp64 for the near-ties of the group
  const float thr2 = gmax - margin2;
  unsigned near = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (sc[c] >= thr2 && sc[c] > -INFINITY) near |= 1u << c;
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    if (near & (1u << c)) {  // warp-uniform; normally true for exactly one column
      const int v = group8 * 8 + c;
      const float4* e = reinterpret_cast<const float4*>(table + (int64_t)v * D);
      double dot = 0.0, nn = 0.0;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int q = lane + 32 * j;
        const float4 xv = q < nvec ? __ldg(e + q) : make_float4(0.f, 0.f, 0.f, 0.f);  // L1/L2 hit
        const double x0 = xv.x, x1 = xv.y, x2 = xv.z, x3 = xv.w;
        dot = fma((double)kreg[j][0], x0, dot); nn = fma(x0, x0, nn);
        dot = fma((double)kreg[j][1], x1, dot); nn = fma(x1, x1, nn);
        dot = fma((double)kreg[j][2], x2, dot); nn = fma(x2, x2, nn);
        dot = fma((double)kreg[j][3], x3, dot); nn = fma(x3, x3, nn);
      }
      dot = warp_sum(dot);
      nn = warp_sum(nn);
      const double score = dot / fmax(sqrt(nn), 1e-8);  // the common factor 1/||kw|| does not change the order
      best = better(best, Best{score, v});
    }
  }
  return best;  // identical in every lane of the warp
}

// warp <-> keyword row (four rows per 128-thread block).  A block-per-row version had a ~12 us serial critical path
// per row (scan, barrier, candidate list, re-scoring, single-thread statistics) and only ~600 rows resident at a time:
// 3.5 waves, 70 us.  One warp per row needs no block barriers or shared memory and keeps all M rows of the benchmark
// shape resident in a single wave.
template <int NV>
__global__ void __launch_bounds__(128, NV <= 4 ? 4 : 2)  // <= 128 registers for D <= 512: all 2048 rows in one wave
vq_select_kernel(const float* __restrict__ kw, const float* __restrict__ table, const float* __restrict__ table_norm,
                 const __half* __restrict__ table_hat, const __half* __restrict__ kw_hat, int64_t M, int V, int D,
                 const float* __restrict__ chunk_max, const float* __restrict__ group_max, int n_chunks,
                 const float* __restrict__ partials, int n_groups,
                 const float* __restrict__ tau_ptr, MaskedCols mc, int64_t* __restrict__ idx_out,
                 float* __restrict__ keywords, float* __restrict__ row_stats, float* __restrict__ code_hist,
                 float* __restrict__ lse1_l2, int phases /* bit 0: row statistics, bit 1: arg-max + gather */,
                 const __half* __restrict__ e16 /* nullable: (Mp, ldE) fp16 e^c, replaces group_max */, int64_t ldE,
                 int e16_is_p /* the scratch holds P'' = exp((c - 1)/tau + 10) instead of e^c (scp_vq_fwd_save) */) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (m >= M) return;  // whole warp
  // ---- statistics of the row: combine the per-group partials of sweep 1 (lane <-> group, fixed tree order)
  if (phases & 1) {
    const float tau = *tau_ptr;
    float z1 = 0.f, c1 = 0.f, gmax = kNegBig;
    for (int g = lane; g < n_groups; g += 32) {
      const float4 q = *reinterpret_cast<const float4*>(partials + (m * n_groups + g) * 4);
      z1 += q.x; c1 += q.y; gmax = fmaxf(gmax, q.z);
    }
    z1 = warp_sum(z1); c1 = warp_sum(c1); gmax = warp_max(gmax);
    float zt = 0.f;
    for (int g = lane; g < n_groups; g += 32) {
      const float4 q = *reinterpret_cast<const float4*>(partials + (m * n_groups + g) * 4);
      zt += q.w * expf((q.z - gmax) / tau);
    }
    zt = warp_sum(zt);
    if (lane == 0) {
      const float lse1 = logf(z1);
      int n_valid = V;
      for (int i = 0; i < mc.n; ++i) n_valid -= (mc.col[i] >= 0 && mc.col[i] < V);
      // -sum p log(p + 1e-9) = H - n_valid*1e-9 + O(1e-8): at temperature 1 every p >= 1/(V e^2) >> 1e-9
      float* rs = row_stats + m * 4;
      rs[0] = lse1;
      rs[1] = gmax / tau + logf(zt);
      rs[2] = lse1 - c1 / z1 - (float)n_valid * 1e-9f;
      lse1_l2[m] = -lse1 * kLog2e;  // sweep 2 adds it with one packed FMA
    }
  }
  if (!(phases & 2)) return;
  // ---- this lane's slice of the keyword row (fp32) and of its unit fp16 copy (the tensor-core operand of sweep 1)
  float kreg[NV][4];
  const int nvec = D >> 2;
  float kss = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int q = lane + 32 * j;
    const float4 k4 = q < nvec ? __ldg(reinterpret_cast<const float4*>(kw + m * D) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    kreg[j][0] = k4.x; kreg[j][1] = k4.y; kreg[j][2] = k4.z; kreg[j][3] = k4.w;
    kss += k4.x * k4.x + k4.y * k4.y + k4.z * k4.z + k4.w * k4.w;
  }
  kss = warp_sum(kss);
  const float margin2 = 4e-6f * sqrtf(kss);  // 2 x the fp32 dot-product error bound, see rescore_chunk
  uint4 khreg[(NV + 1) / 2];
#pragma unroll
  for (int j = 0; j < (NV + 1) / 2; ++j) {
    const int q = lane + 32 * j;
    khreg[j] = q < (D >> 3) ? __ldg(reinterpret_cast<const uint4*>(kw_hat + m * D) + q) : make_uint4(0, 0, 0, 0);
  }
  Best best{0.0, -1};
  if (kss > 0.f) {  // a zero keyword ties every cosine at 0: the first unmasked column wins (handled below)
    // approximate (fp16-product) row maximum.  The row's chunk maxima are read ONCE from global memory with 16-byte
    // loads and parked in this warp's slice of shared memory; the candidate pass then reads them back from there.
    // (Two dependent passes of scalar global loads were the kernel's critical path; a fully unrolled register-resident
    // scan instead inlined the re-scoring code 64 times and thrashed the instruction cache: 47 % stall_no_inst.)
    extern __shared__ float s_scan[];
    float* mine = s_scan + (size_t)(threadIdx.x >> 5) * n_chunks;
    const float4* cm4 = reinterpret_cast<const float4*>(chunk_max + m * n_chunks);  // n_chunks % 8 == 0 (Vp % 256 == 0)
    const int n4 = n_chunks >> 2;
    float mx = kNegBig;
#pragma unroll 4
    for (int q = lane; q < n4; q += 32) {
      const float4 t = __ldg(cm4 + q);
      *reinterpret_cast<float4*>(mine + 4 * q) = t;
      mx = fmaxf(mx, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
    }
    mx = warp_max(mx);  // (the shuffles also order the shared-memory writes before the reads below)
    __syncwarp();
    const float thr = mx - kRescueMargin;
    // group filter from the e^c scratch: fp16(e^c) >= e^thr (1 - 1e-3) holds for every column with c >= thr (fp16 rounding
    // 4.9e-4 + ex2.approx 2^-22), i.e. the filter only ever admits MORE groups than the exact test c >= thr
    const float thr_e = expf(thr) * (1.0f - 1.0e-3f);
    // the same filter on P'' (monotone in c; relative error of the stored value: fp16 rounding 4.9e-4 + the repeated-squaring
    // chain ~3e-6, i.e. LESS than 1e-3 / tau in c).  Below the fp16 normal range the rounding is absolute (6e-8): a threshold
    // that low (row maximum below about -0.8) admits every group of the candidate chunk instead.
    const float thr_p = expf((thr - 1.0f) / *tau_ptr + 10.0f) * (1.0f - 1.0e-3f);
    const bool p_all = e16_is_p && thr_p < 1.3e-4f;
    // every chunk whose maximum could hide the true arg-max is re-scored exactly (usually one or two per row), and
    // inside it only the 8-column groups whose own maximum qualifies
#pragma unroll 1
    for (int c0 = 0; c0 < n_chunks; c0 += 32) {
      const float v = c0 + lane < n_chunks ? mine[c0 + lane] : kNegBig;
      unsigned todo = __ballot_sync(0xffffffffu, v >= thr && v > kNegBig);
#pragma unroll 1
      while (todo) {
        const int chunk = c0 + __ffs(todo) - 1;
        todo &= todo - 1;
        unsigned gq;  // bit g: group g of this chunk can hold the arg-max (warp-uniform)
        if (e16) {
          const float e = __half2float(e16[m * ldE + (int64_t)chunk * 32 + lane]);
          const unsigned cols = __ballot_sync(0xffffffffu, p_all || e >= (e16_is_p ? thr_p : thr_e));
          gq = ((cols & 0xffu) ? 1u : 0u) | ((cols & 0xff00u) ? 2u : 0u) | ((cols & 0xff0000u) ? 4u : 0u) |
               ((cols & 0xff000000u) ? 8u : 0u);
        } else {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(group_max + (m * n_chunks + chunk) * 4));
          gq = (g4.x >= thr ? 1u : 0u) | (g4.y >= thr ? 2u : 0u) | (g4.z >= thr ? 4u : 0u) | (g4.w >= thr ? 8u : 0u);
        }
#pragma unroll 1
        for (int g = 0; g < 4; ++g)
          if ((gq >> g) & 1u)
            best = better(best, rescore_chunk<NV>(table, table_norm, table_hat, V, D, chunk * 4 + g, kreg, khreg,
                                                  margin2, thr, mc, lane));
      }
    }
  }
  int k = best.idx;
  if (k < 0) {  // no finite candidate (zero keyword / all columns masked): first unmasked column, like torch's argmax
    k = 0;
    while (k < V - 1 && is_masked(mc, k)) ++k;
  }
  if (lane == 0) {
    idx_out[m] = k;
    atomicAdd(&code_hist[k], 1.0f);
  }
  // keywords = E[k]   (value of subword_prob @ E, kw_branches.py:195)
  const float4* src = reinterpret_cast<const float4*>(table + (int64_t)k * D);
  float4* dst = reinterpret_cast<float4*>(keywords + m * D);
  for (int d4 = lane; d4 < D / 4; d4 += 32) dst[d4] = __ldg(src + d4);
}

// =====================================================================================================================
// sweep 2: avg_probs[v] = (1/M) sum_m exp(c[m,v] - lse1[m])       X = Ehat (rows v), Y = khat (rows m)
// =====================================================================================================================
struct Sweep2Epi {
  struct Params {
    const float* lse1_l2;  // (Mp2,) MINUS lse at temperature 1 times log2(e); -1e30 for padding
    float* avg_probs;      // (Vp,)
    float inv_m;
    int V;
    MaskedCols mc;
  };
  // [128 floats: half-1 partial sums][per warp: the 128 normalisers of the columns it consumes in the current tile]
  static constexpr int kWarpVec = 128 * 4;
  static constexpr int kSmemBytes = tc::kTileM * 4 + tc::kEpiWarps * kWarpVec;
  const Params& p;
  int v, half, row_in_tile, lane, nt_stride, nt_end;
  tc::f32x2 acc2[2];
  float* s_acc;
  uint32_t wsm;  // shared address of this warp's normaliser vector
  float4 pre;    // normalisers of the NEXT tile, fetched one tile ahead (a per-chunk __ldg sat on the critical path:
                 // ncu attributed 18 % of all stall samples to its long-scoreboard wait)
  __device__ __forceinline__ const float4* vec_src(int nt) const {
    return reinterpret_cast<const float4*>(p.lse1_l2 + (int64_t)nt * 256 + half * 128) + lane;
  }
  __device__ __forceinline__ Sweep2Epi(const Params& p_, const WorkInfo& w, const tc::EpiCtx& ctx)
      : p(p_), v(w.m_tile * tc::kTileM + ctx.row_in_tile), half(ctx.half), row_in_tile(ctx.row_in_tile),
        lane(ctx.tid & 31), nt_stride(w.nt_stride), nt_end(w.nt_end), s_acc(reinterpret_cast<float*>(ctx.smem)) {
    acc2[0] = acc2[1] = tc::pack2(0.f, 0.f);
    wsm = tc::smem_u32(ctx.smem + tc::kTileM * 4 + (ctx.tid >> 5) * kWarpVec);
    pre = w.nt_first < w.nt_end ? __ldg(vec_src(w.nt_first)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void tile_begin(int nt) {
    __syncwarp();  // the previous tile's reads of the vector are complete
    tc::sts128f(wsm + lane * 16, pre);
    __syncwarp();
    const int next = nt + nt_stride;
    if (next < nt_end) pre = __ldg(vec_src(next));
  }
  __device__ __forceinline__ void tile_end(int) {}
  __device__ __forceinline__ void chunk(int col0, float (&c)[1][32]) {
    const uint32_t src = wsm + (uint32_t)(col0 & 127) * 4;  // NEGATED lse * log2(e) of the 32 columns (broadcast reads)
    const tc::f32x2 kl = tc::pack2(kLog2e, kLog2e);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 l = tc::lds128f(src + i * 16);
      acc2[0] = tc::add2(acc2[0], tc::ex2_2(tc::fma2(tc::pack2(c[0][4 * i + 0], c[0][4 * i + 1]), kl, tc::pack2(l.x, l.y))));
      acc2[1] = tc::add2(acc2[1], tc::ex2_2(tc::fma2(tc::pack2(c[0][4 * i + 2], c[0][4 * i + 3]), kl, tc::pack2(l.z, l.w))));
    }
  }
  __device__ __forceinline__ void finish() {
    // the two warps that own a row add their halves in a fixed order
    const float acc = tc::hsum2(tc::add2(acc2[0], acc2[1]));
    if (half == 1) s_acc[row_in_tile] = acc;
    tc::named_bar_sync(tc::kEpiBarrierId, tc::kEpiThreads);
    if (half == 0) p.avg_probs[v] = (v < p.V && !is_masked(p.mc, v)) ? (acc + s_acc[row_in_tile]) * p.inv_m : 0.f;
  }
};

// =====================================================================================================================
// column sums of the normalised logits: avg_probs[v] = 1/M sum_m e^{c[m,v]} / Z_m      (HBM-bound, one pass)
// =====================================================================================================================
// e16 (Mp, ldE) fp16 = e^c from sweep 1, lse1_l2[m] = -log2(Z_m) from the statistics phase of vq_select.  Block <-> strip
// of 64 columns x all M rows: warp w takes the rows m = w (mod 8), lane l the columns 2l, 2l+1 of the strip (one 128-byte
// row segment per warp load, eight loads in flight per lane); the eight warps are combined through shared memory in a
// fixed order, so the result is deterministic.  Vp/64 = 772 strips at the full vocabulary = 5.2 blocks per SM: the SMs that
// draw six blocks set the time of this pipe-bound pass, so the rows are split over blockIdx.y as well (kColsumRowSplit row
// ranges -> 3088 short blocks, dynamically scheduled) and vq_metrics_kernel adds the row-range partials in a fixed order.
constexpr int kColsumCols = 64;
constexpr int kColsumRowSplit = 4;   // row ranges of the column-sum pass (M >= 1024), 1 below
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// PMODE: the scratch holds P'' = exp((c - 1)/tau + 10) (scp_vq_fwd_save):  e^c / Z_m = 2^(tau lg2 P'' + (1 - 10 tau) log2 e - log2 Z_m)
// -- two SFU ops per element instead of a multiply, which makes the pass SFU-bound (72 us against 39 us at M = 2048,
// V = 49408; an FMA-pipe polynomial for the 2^x half was issue-bound and slower: 102 us); the backward pass saves more than
// that.  P'' = 0 (masked / padding columns) gives 2^-inf = 0.
template <bool PMODE>
__global__ void __launch_bounds__(256)
vq_colsum_kernel(const __half* __restrict__ e16, int64_t ldE, const float* __restrict__ lse1_l2, int64_t M_all, int V,
                 float inv_m, MaskedCols mc, const float* __restrict__ tau_ptr, float* __restrict__ avg_probs,
                 int64_t rows_per_block /* blockIdx.y selects a row range; its sums go to avg_probs + blockIdx.y * ldE */) {
  __shared__ float2 s_part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t m_begin = (int64_t)blockIdx.y * rows_per_block;
  const int64_t M = m_begin + rows_per_block < M_all ? m_begin + rows_per_block : M_all;  // end of this block's rows
  avg_probs += (int64_t)blockIdx.y * ldE;
  const int64_t col = (int64_t)blockIdx.x * kColsumCols + 2 * lane;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(e16 + col);
  const int64_t ld32 = ldE >> 1;
  const float tau = PMODE ? __ldg(tau_ptr) : 1.0f;
  const float shift = PMODE ? (1.0f - 10.0f * tau) * kLog2e : 0.f;
  float a0 = 0.f, a1 = 0.f;
  auto add = [&](uint32_t hbits, float w) {
    const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&hbits));
    if (PMODE) {
      const float c = w + shift;
      a0 += tc::fast_ex2(fmaf(tau, fast_lg2(e.x), c));
      a1 += tc::fast_ex2(fmaf(tau, fast_lg2(e.y), c));
    } else {
      const float wj = tc::fast_ex2(w);
      a0 = fmaf(e.x, wj, a0);
      a1 = fmaf(e.y, wj, a1);
    }
  };
  int64_t m = m_begin + warp;
  for (; m + 56 < M; m += 64) {
    uint32_t h[8];
    float w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] = __ldcs(src + (m + 8 * j) * ld32);   // streamed once: evict first
      w[j] = __ldg(lse1_l2 + m + 8 * j);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) add(h[j], w[j]);
  }
  for (; m < M; m += 8) add(__ldcs(src + m * ld32), __ldg(lse1_l2 + m));
  s_part[warp][lane] = make_float2(a0, a1);
  __syncthreads();
  if (warp == 0) {
    float2 t = s_part[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      t.x += s_part[w][lane].x;
      t.y += s_part[w][lane].y;
    }
    const int v0 = (int)col, v1 = v0 + 1;
    float2 o;
    o.x = (v0 < V && !is_masked(mc, v0)) ? t.x * inv_m : 0.f;
    o.y = (v1 < V && !is_masked(mc, v1)) ? t.y * inv_m : 0.f;
    *reinterpret_cast<float2*>(avg_probs + col) = o;
  }
}

// =====================================================================================================================
// metrics: code_perplexity, prob_perplexity, diversity_loss, ent_per_t      (one block)
// =====================================================================================================================
constexpr int kMetricBlocks = 128;

// Stage 1 (all blocks): per-block partial sums of h log(h + 1e-7) over the code histogram and the average softmax.
// Stage 2 (the block that draws the last ticket): fixed-order sum of the partials, perplexities, diversity loss,
// ent_per_t.  `ticket` must be zero on entry (vq_prep_kw / a memset clears it) and is left at zero.
__global__ void __launch_bounds__(256)
vq_metrics_kernel(const float* __restrict__ code_hist, float* __restrict__ avg_probs,
                  const float* __restrict__ row_stats, int64_t M, int K, int V,
                  float* __restrict__ partial /* (kMetricBlocks, 2) | ticket | (kMetricBlocks,) */,
                  unsigned int* __restrict__ ticket, float* __restrict__ metrics,
                  const float* __restrict__ avg_part = nullptr /* (n_part, Vp): row-range partials of the column sums */,
                  int n_part = 0, int Vp = 0) {
  __shared__ float s_red[3][8];
  __shared__ bool s_last;
  const float inv_m = 1.0f / (float)M;
  float* partial_sum = partial + kMetricBlocks * 2 + 1;  // third partial (sum of avg_probs), behind the ticket word
  float hc = 0.f, hp = 0.f, sp = 0.f;
  for (int v = blockIdx.x * 256 + threadIdx.x; v < (avg_part ? Vp : V); v += kMetricBlocks * 256) {
    float a = 0.f;
    if (avg_part) {  // avg_probs = sum of the row-range partials, fixed order (also the zero padding of [V, Vp))
      for (int y = 0; y < n_part; ++y) a += avg_part[(int64_t)y * Vp + v];
      avg_probs[v] = a;
    } else if (avg_probs) {
      a = avg_probs[v];
    }
    if (v >= V) continue;
    const float h = code_hist[v] * inv_m;  // my_vector_quantizer.py:94-99
    hc += h * logf(h + 1e-7f);
    if (avg_probs) {  // :119-121
      hp += a * logf(a + 1e-7f);
      sp += a;
    }
  }
  hc = warp_sum(hc);
  hp = warp_sum(hp);
  sp = warp_sum(sp);
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = hc;
    s_red[1][threadIdx.x >> 5] = hp;
    s_red[2][threadIdx.x >> 5] = sp;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_red[threadIdx.x][i];
    __stcg(threadIdx.x < 2 ? &partial[blockIdx.x * 2 + threadIdx.x] : &partial_sum[blockIdx.x], t);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == (unsigned)kMetricBlocks - 1u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    float hc2 = 0.f, hp2 = 0.f, sp2 = 0.f;
    for (int i = threadIdx.x; i < kMetricBlocks; i += 32) {
      hc2 += __ldcg(&partial[2 * i]);
      hp2 += __ldcg(&partial[2 * i + 1]);
      sp2 += __ldcg(&partial_sum[i]);
    }
    hc2 = warp_sum(hc2);
    hp2 = warp_sum(hp2);
    sp2 = warp_sum(sp2);
    if (threadIdx.x == 0) {
      metrics[0] = expf(-hc2);
      // avg_probs is a mean of softmax rows, so its entries sum to one exactly; the column-sum path carries a common
      // factor S = 1 + O(1e-5) (fp16 rounding of e^c against the unrounded normaliser).  The entropy of a/S follows in
      // closed form: -sum (a/S) log(a/S) = (-sum a log a)/S + log S  -- an identity when S = 1.
      const float pp = !avg_probs ? nanf("") : (sp2 > 0.f ? expf(-hp2 / sp2 + logf(sp2)) : expf(-hp2));
      metrics[1] = pp;
      metrics[2] = ((float)V - pp) / (float)V;  // diversity_loss, :155-158
      *ticket = 0u;
    }
  }
  // ent_per_t[i] = mean_b entropy[b*K + i]     (:104-116)
  const int64_t Bsz = M / K;
  for (int i = threadIdx.x >> 5; i < K; i += (blockDim.x >> 5)) {
    float sum = 0.f;
    for (int64_t b = threadIdx.x & 31; b < Bsz; b += 32) sum += row_stats[(b * K + i) * 4 + 2];
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) metrics[3 + i] = sum / (float)Bsz;
  }
}

// =====================================================================================================================
// backward
// =====================================================================================================================
// warp <-> row: ghat = g/||g|| (fp16), gscale = ||g||, s0 = <ghat, mean(E)> / norm_ref
__global__ void vq_bwd_prep_kernel(const float* __restrict__ g, int64_t M, int64_t Mp, int D,
                                   const float* __restrict__ table_mean /* (D+1): [D] = norm_ref */,
                                   __half* __restrict__ g_hat, float* __restrict__ g_aux /* (Mp,2): scale, s0 */,
                                   unsigned int* __restrict__ flags /* nullable: pipeline hand-off counters */,
                                   int n_flags) {
  const int lane = threadIdx.x & 31;
  if (flags && blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_flags; i += blockDim.x) flags[i] = 0u;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= Mp) return;
  __half* dst = g_hat + m * D;
  if (m >= M) {
    for (int d = lane; d < D; d += 32) dst[d] = __float2half(0.f);
    if (lane == 0) { g_aux[m * 2] = 0.f; g_aux[m * 2 + 1] = 0.f; }
    return;
  }
  const float* src = g + m * D;
  float ss = 0.f, dm = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float x = src[d];
    ss = fmaf(x, x, ss);
    dm = fmaf(x, table_mean[d], dm);
  }
  ss = warp_sum(ss);
  dm = warp_sum(dm);
  const float n = sqrtf(ss);
  const float inv = n > 0.f ? 1.0f / n : 0.f;
  for (int d = lane; d < D; d += 32) dst[d] = __float2half_rn(src[d] * inv);
  if (lane == 0) {
    g_aux[m * 2] = n;
    g_aux[m * 2 + 1] = dm * inv / table_mean[D];
  }
}

struct Sweep3Epi {
  struct Params {
    const float* row_stats;   // (M,4): [1] = lse at tau
    const float* g_aux;       // (Mp,2): scale, s0
    const float* table_norm;  // (Vp,)
    const float* table_mean;  // [D] = norm_ref
    const float* tau;
    __half* pq;               // (2*Mp, Vp): rows [0,Mp) = Q~, rows [Mp,2Mp) = P~
    float* partials;          // (Mp, 2*n_groups, 4): sum Q~, sum P~, sum Q~ c, sum P~ c   (all carry the 2^14 scale)
    int64_t M, Mp, Vp;
    int n_groups, V, D;
    int want_tau;             // accumulate the two c-weighted sums (only the learnable-temperature gradient needs them)
    MaskedCols mc;
  };
  // Each warp owns a private 2 KB staging area (32 rows x 32 columns, used for Q~ and then for P~): the accumulator
  // layout gives every lane one ROW, but a store instruction in which 32 lanes touch 32 different rows costs 32 LSU
  // transactions; transposing through shared memory turns it into 64-byte row segments (8 rows per instruction, full
  // sectors).  One tile-sized buffer instead of two frees 16 KB for a fifth ring stage (the sweep is feed-bound).
  static constexpr int kWarpStage = 32 * 64;
  static constexpr int kWarpVec = 64 * 4;  // table norms of the 64 columns a warp consumes per tile
  static constexpr int kSmemBytes = tc::kEpiWarps * (kWarpStage + kWarpVec);
  const Params& p;
  int64_t row, row0;
  int slot, lane, half, nt_stride, nt_end;
  uint32_t stage;  // shared address of this warp's staging area
  uint32_t wsm;    // shared address of this warp's column-norm vector
  float4 pre;      // column norms of the NEXT tile, fetched one tile ahead (see Sweep2Epi)
  float k_tau, bias, s0, inv_norm_ref;  // bias = log2(kPScale) - lse_tau*log2(e): P~ = 2^(c*k_tau + bias)
  tc::f32x2 acc_q, acc_p, acc_qc, acc_pc;
  __device__ __forceinline__ const float4* vec_src(int nt) const {
    return reinterpret_cast<const float4*>(p.table_norm + (int64_t)nt * 128 + half * 64) + (lane & 15);
  }
  __device__ __forceinline__ Sweep3Epi(const Params& p_, const WorkInfo& w, const tc::EpiCtx& ctx)
      : p(p_), row((int64_t)w.m_tile * tc::kTileM + ctx.row_in_tile), slot(w.n_group * 2 + ctx.half),
        half(ctx.half), nt_stride(w.nt_stride), nt_end(w.nt_end) {
    lane = ctx.tid & 31;
    row0 = row - lane;
    stage = tc::smem_u32(ctx.smem + (ctx.tid >> 5) * kWarpStage);
    wsm = tc::smem_u32(ctx.smem + tc::kEpiWarps * kWarpStage + (ctx.tid >> 5) * kWarpVec);
    pre = w.nt_first < w.nt_end ? __ldg(vec_src(w.nt_first)) : make_float4(0.f, 0.f, 0.f, 0.f);
    k_tau = kLog2e / __ldg(p.tau);
    const bool valid = row < p.M;
    bias = valid ? 14.0f - p.row_stats[row * 4 + 1] * kLog2e : -1.0e30f;  // 2^14 = kPScale; padding rows: P = 0
    s0 = p.g_aux[row * 2 + 1];
    inv_norm_ref = 1.0f / p.table_mean[p.D];
    acc_q = acc_p = acc_qc = acc_pc = tc::pack2(0.f, 0.f);
  }
  __device__ __forceinline__ void tile_begin(int nt) {
    __syncwarp();  // the previous tile's reads of the vector are complete
    // staged as r_v = ||e_v|| / norm_ref, so that T' = (ghat . ehat_v) * r_v - s0 is one FMA per logit
    if (lane < 16)
      tc::sts128f(wsm + lane * 16, make_float4(pre.x * inv_norm_ref, pre.y * inv_norm_ref, pre.z * inv_norm_ref,
                                               pre.w * inv_norm_ref));
    __syncwarp();
    const int next = nt + nt_stride;
    if (next < nt_end) pre = __ldg(vec_src(next));
  }
  __device__ __forceinline__ void tile_end(int) {}
  __device__ __forceinline__ void chunk(int col0, float (&v)[2][32]) {
    float(&c)[32] = v[0];
    float(&t)[32] = v[1];
    if (col0 + 32 > p.V || chunk_has_mask(p.mc, col0)) {
      const uint32_t bits = chunk_mask_bits(p.mc, col0, p.V);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if ((bits >> i) & 1u) c[i] = kNegBig;
    }
    const uint32_t n_src = wsm + (uint32_t)(col0 & 63) * 4;  // ||e_v|| of the 32 columns (broadcast reads)
    // staging layout: row r at r*64 B, its four 16-byte units XOR-swizzled with (r >> 1) & 3 (conflict-free both ways)
    const uint32_t qs = stage + lane * 64;
    const int sw = (lane >> 1) & 3;
    const tc::f32x2 kt2 = tc::pack2(k_tau, k_tau), b2 = tc::pack2(bias, bias), ns0 = tc::pack2(-s0, -s0);
    uint4 pk_q[4], pk_p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t wq[4], wp[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 rv = tc::lds128f(n_src + (2 * i + h) * 16);
        const tc::f32x2 rr[2] = {tc::pack2(rv.x, rv.y), tc::pack2(rv.z, rv.w)};
#pragma unroll
        for (int j = 0; j < 2; ++j) {  // packed pairs of adjacent columns
          const int e = 8 * i + 4 * h + 2 * j;
          const tc::f32x2 cc = tc::pack2(c[e], c[e + 1]);
          const tc::f32x2 pj = tc::ex2_2(tc::fma2(cc, kt2, b2));               // P~ = 2^14 softmax_tau
          const tc::f32x2 tj = tc::fma2(tc::pack2(t[e], t[e + 1]), rr[j], ns0);  // (g . e_v)/(|g| norm_ref) - s0
          const tc::f32x2 qj = tc::mul2(pj, tj);
          acc_p = tc::add2(acc_p, pj);
          acc_q = tc::add2(acc_q, qj);
          if (p.want_tau) {  // masked columns: P~ = 0 exactly and c = -1e30 is finite, so the products vanish
            acc_pc = tc::fma2(pj, cc, acc_pc);
            acc_qc = tc::fma2(qj, cc, acc_qc);
          }
          float a, b;
          __half2 hh;
          tc::unpack2(qj, a, b);
          hh = __floats2half2_rn(a, b); wq[2 * h + j] = *reinterpret_cast<uint32_t*>(&hh);
          tc::unpack2(pj, a, b);
          hh = __floats2half2_rn(a, b); wp[2 * h + j] = *reinterpret_cast<uint32_t*>(&hh);
        }
      }
      pk_q[i] = make_uint4(wq[0], wq[1], wq[2], wq[3]);
      pk_p[i] = make_uint4(wp[0], wp[1], wp[2], wp[3]);
    }
    // transpose + store, Q~ first and then P~ through the same staging tile
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __syncwarp();  // the previous read-out of the staging area is complete
#pragma unroll
      for (int i = 0; i < 4; ++i) tc::sts128(qs + ((i ^ sw) << 4), which == 0 ? pk_q[i] : pk_p[i]);
      __syncwarp();
      // read-out: instruction k covers rows 8k..8k+7, four lanes per row -> 64-byte contiguous global segments
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = 8 * k + (lane >> 2);
        const int u = lane & 3;
        const uint4 v4 = tc::lds128(stage + r * 64 + ((u ^ ((r >> 1) & 3)) << 4));
        const int64_t grow = (which == 0 ? 0 : p.Mp) + row0 + r;
        *reinterpret_cast<uint4*>(p.pq + grow * p.Vp + col0 + u * 8) = v4;
      }
    }
  }
  __device__ __forceinline__ void finish() {
    *reinterpret_cast<float4*>(p.partials + (row * (2 * p.n_groups) + slot) * 4) =
        make_float4(tc::hsum2(acc_q), tc::hsum2(acc_p), tc::hsum2(acc_qc), tc::hsum2(acc_pc));
  }
};

// Saved-numerator backward sweep (scp_vq_bwd_saved): X = ghat (resident), Y = Ehat, ONE N = 256 accumulator T = ghat . ehat_v.
// Per logit the epilogue reads T from TMEM and the forward's fp16 P'' from global memory and writes
//     Q''[m,v] = P''[m,v] (T r_v - s0)       r_v = ||e_v|| / norm_ref
// as fp16 plus the row sums of Q'' and P''.  No exponential, one TMEM value and 2 + 2 bytes per logit.
// The thread's row segment of P'' for one tile (4 chunks x 64 B) is kept in registers and refilled one tile ahead: the
// load of chunk k of the NEXT tile is issued as soon as chunk k of this tile has been consumed (`kUnrollTile`: k is a
// compile-time constant after unrolling, so the buffer stays in registers).
struct SweepTEpi {
  struct Params {
    const __half* p16;        // (Mp, ld) fp16 P'' written by scp_vq_fwd_save
    __half* q16;              // (Mp, ld) fp16 Q''
    int64_t ld;               // = Vp
    const float* g_aux;       // (Mp,2): |g|, s0
    const float* table_norm;  // (Vp,)
    const float* table_mean;  // [D] = norm_ref
    float* partials;          // (Mp, 2*n_groups, 4): sum Q'', sum P'', 0, 0
    int n_groups, D;
    int dbg;                  // timing ablations (env SCP_VQ_ST_DBG, results invalid): 1 = no Q'' stores, 2 = no P'' loads
  };
  static constexpr bool kUnrollTile = true;
  static constexpr int kChunks = kVqBN / 32 / 2;   // chunks per warp and tile (two warps share a lane quadrant)
  static constexpr int kWarpVec = 32 * 4;          // r_v of the chunk being consumed (lane <-> column, read back broadcast)
  static constexpr int kSmemBytes = tc::kEpiWarps * kWarpVec;
  const Params& p;
  int64_t row;
  int slot, n_slots, lane, half, cur_nt, nt_stride, nt_end;
  uint32_t wsm;
  float pre;  // r_v of this lane's column of the NEXT chunk
  float s0, inv_norm_ref;
  uint32_t pbuf[kChunks][16];
  tc::f32x2 acc_q[2], acc_p[2];
  __device__ __forceinline__ float vec_load(int nt, int k) const {
    return __ldg(p.table_norm + (int64_t)nt * kVqBN + (half * kChunks + k) * 32 + lane) * inv_norm_ref;
  }
  __device__ __forceinline__ const __half* p_src(int nt, int k) const {
    return p.p16 + row * p.ld + (int64_t)nt * kVqBN + (half * kChunks + k) * 32;
  }
  __device__ __forceinline__ void fetch(int nt, int k) {
    if (p.dbg & 2) return;
    const __half* src = p_src(nt, k);
    tc::ldg256_stream(src, pbuf[k]);
    tc::ldg256_stream(src + 16, pbuf[k] + 8);
  }
  __device__ __forceinline__ SweepTEpi(const Params& p_, const WorkInfo& w, const tc::EpiCtx& ctx)
      : p(p_), row((int64_t)w.m_tile * tc::kTileM + ctx.row_in_tile), slot(w.n_group * ctx.parts + ctx.half),
        n_slots(ctx.parts * p_.n_groups), half(ctx.half), cur_nt(w.nt_first), nt_stride(w.nt_stride), nt_end(w.nt_end) {
    lane = ctx.tid & 31;
    wsm = tc::smem_u32(ctx.smem + (ctx.tid >> 5) * kWarpVec);
    const bool any = w.nt_first < w.nt_end;
    inv_norm_ref = 1.0f / p.table_mean[p.D];
    pre = any ? vec_load(w.nt_first, 0) : 0.f;
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      if (any) fetch(w.nt_first, k);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pbuf[k][i] = 0u;
      }
    }
    s0 = p.g_aux[row * 2 + 1];
    acc_q[0] = acc_q[1] = acc_p[0] = acc_p[1] = tc::pack2(0.f, 0.f);
  }
  __device__ __forceinline__ void tile_begin(int nt) { cur_nt = nt; }
  __device__ __forceinline__ void tile_end(int) {}
  __device__ __forceinline__ void chunk_k(int k, int col0, float (&v)[1][32]) {
    float(&t)[32] = v[0];
    // r_v of the chunk's 32 columns: parked by lane <-> column, read back as broadcast vectors; the next chunk's value is
    // fetched while this one is consumed
    __syncwarp();  // the previous chunk's reads of the vector are complete
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(wsm + (uint32_t)lane * 4), "f"(pre) : "memory");
    __syncwarp();
    {
      const int nk = k + 1 < kChunks ? k + 1 : 0;
      const int nnt = k + 1 < kChunks ? cur_nt : cur_nt + nt_stride;
      if (nnt < nt_end) pre = vec_load(nnt, nk);
    }
    const tc::f32x2 ns0 = tc::pack2(-s0, -s0);
    uint32_t hq[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 rv = tc::lds128f(wsm + i * 16);
      const tc::f32x2 rr[2] = {tc::pack2(rv.x, rv.y), tc::pack2(rv.z, rv.w)};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int e = 4 * i + 2 * j;
        const tc::f32x2 tj = tc::fma2(tc::pack2(t[e], t[e + 1]), rr[j], ns0);  // (g . e_v)/(|g| norm_ref) - s0
        const tc::f32x2 pj = tc::cvt_f32x2_f16x2(pbuf[k][e >> 1]);
        const tc::f32x2 qj = tc::mul2(pj, tj);
        acc_p[j] = tc::add2(acc_p[j], pj);
        acc_q[j] = tc::add2(acc_q[j], qj);
        hq[e >> 1] = tc::cvt_f16x2(qj);
      }
    }
    if (!(p.dbg & 1)) {
      __half* dst = p.q16 + row * p.ld + col0;
      tc::stg256(dst, hq);
      tc::stg256(dst + 16, hq + 8);
    }
    const int next = cur_nt + nt_stride;
    if (next < nt_end) fetch(next, k);  // this row's chunk k of the next tile: in flight for a whole tile
  }
  __device__ __forceinline__ void finish() {
    *reinterpret_cast<float4*>(p.partials + (row * n_slots + slot) * 4) =
        make_float4(tc::hsum2(tc::add2(acc_q[0], acc_q[1])), tc::hsum2(tc::add2(acc_p[0], acc_p[1])), 0.f, 0.f);
  }
};
// bring-up switch: SCP_VQ_SAVED=0 makes scp_vq_bwd_saved ignore the saved numerators (A/B against the recompute path)
static bool vq_saved_enabled() {
  static const bool on = [] { const char* e = getenv("SCP_VQ_SAVED"); return !(e && e[0] == '0'); }();
  return on;
}

// plain accumulator store: out[k_split][x][row][col]  (fp32)
template <int NX>
struct StoreEpi {
  struct Params {
    float* out;
    int64_t rows;  // rows per x-slab (Mp)
    int ld;        // columns (D)
  };
  static constexpr int kSmemBytes = 0;
  const Params& p;
  int64_t row;
  int ks;
  __device__ __forceinline__ StoreEpi(const Params& p_, const WorkInfo& w, const tc::EpiCtx& ctx)
      : p(p_), row((int64_t)w.m_tile * tc::kTileM + ctx.row_in_tile), ks(w.k_split) {}
  __device__ __forceinline__ void tile_begin(int) {}
  __device__ __forceinline__ void tile_end(int) {}
  __device__ __forceinline__ void chunk(int col0, float (&v)[NX][32]) {
#pragma unroll
    for (int x = 0; x < NX; ++x) {
      float4* dst = reinterpret_cast<float4*>(p.out + (((int64_t)ks * NX + x) * p.rows + row) * p.ld + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[x][4 * i], v[x][4 * i + 1], v[x][4 * i + 2], v[x][4 * i + 3]);
    }
  }
  __device__ __forceinline__ void finish() {}
};

// block (D/4 threads) <-> row, thread <-> 4 channels:  g_khat = (U - s W) * scale / tau ;
// g_kw = (g_khat - <g_khat,khat> khat) / ||kw||.  (A warp-per-row version left most of the 33 MB of split-K partials
// to 256 resident warps and took 28 us; one block per row keeps every SM streaming.)
__global__ void __launch_bounds__(256)
vq_bwd_finalize_kernel(const float* __restrict__ uw /* (k_splits, 2, Mp, D) */, int k_splits, int64_t M, int64_t Mp,
                       int D, const float* __restrict__ partials, int n_groups, const float* __restrict__ g_aux,
                       const float* __restrict__ kw, const float* __restrict__ row_stats,
                       const float* __restrict__ table_mean, const float* __restrict__ tau_ptr,
                       float* __restrict__ g_kw, float* __restrict__ g_tau) {
  __shared__ float s_red[8];
  const int64_t m = blockIdx.x;
  const int d0 = threadIdx.x * 4;
  const float tau = *tau_ptr;
  float sq = 0.f, sp = 0.f, sqc = 0.f, spc = 0.f;
  for (int g = 0; g < n_groups; ++g) {  // same address in every thread: broadcast loads
    const float4 q = *reinterpret_cast<const float4*>(partials + (m * n_groups + g) * 4);
    sq += q.x; sp += q.y; sqc += q.z; spc += q.w;
  }
  const float s_adj = sp > 0.f ? sq / sp : 0.f;
  // U, W carry the row's own normaliser sum P (2^14 on the recompute path, sum_v P'' on the saved-numerator path)
  const float inv_sp = sp > 0.f ? 1.0f / sp : 0.f;
  const float scale = g_aux[m * 2] * table_mean[D] * inv_sp / tau;
  const float inv_norm = row_stats[m * 4 + 3];
  float4 u = make_float4(0.f, 0.f, 0.f, 0.f), w = u;
  for (int ks = 0; ks < k_splits; ++ks) {
    const float4 a = *reinterpret_cast<const float4*>(uw + (((int64_t)ks * 2 + 0) * Mp + m) * D + d0);
    const float4 b = *reinterpret_cast<const float4*>(uw + (((int64_t)ks * 2 + 1) * Mp + m) * D + d0);
    u.x += a.x; u.y += a.y; u.z += a.z; u.w += a.w;
    w.x += b.x; w.y += b.y; w.z += b.z; w.w += b.w;
  }
  const float4 k4 = *reinterpret_cast<const float4*>(kw + m * D + d0);
  const float gk[4] = {(u.x - s_adj * w.x) * scale, (u.y - s_adj * w.y) * scale, (u.z - s_adj * w.z) * scale,
                       (u.w - s_adj * w.w) * scale};
  const float kh[4] = {k4.x * inv_norm, k4.y * inv_norm, k4.z * inv_norm, k4.w * inv_norm};
  float proj = gk[0] * kh[0] + gk[1] * kh[1] + gk[2] * kh[2] + gk[3] * kh[3];
  proj = warp_sum(proj);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = proj;
  __syncthreads();
  proj = 0.f;
  for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) proj += s_red[i];  // fixed order
  *reinterpret_cast<float4*>(g_kw + m * D + d0) =
      make_float4((gk[0] - proj * kh[0]) * inv_norm, (gk[1] - proj * kh[1]) * inv_norm,
                  (gk[2] - proj * kh[2]) * inv_norm, (gk[3] - proj * kh[3]) * inv_norm);
  if (g_tau && threadIdx.x == 0) {
    // d/dtau = -(1/tau^2) sum_v P (T - s) c      (T in true units = T' * |g| * norm_ref)
    const float contrib = -(g_aux[m * 2] * table_mean[D]) * (sqc - s_adj * spc) * inv_sp / (tau * tau);
    atomicAdd(g_tau, contrib);
  }
}


// =====================================================================================================================
// dense-input form of SimpleVectorQuantizer.forward (my_vector_quantizer.py:64-165): the caller already holds the (M,V)
// score matrix.  HBM/L2-bound helper kernels; not on the fused hot path.
// =====================================================================================================================
__device__ __forceinline__ float block_max_256(float v, float* s_red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = s_red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) t = fmaxf(t, s_red[i]);
  return t;
}
__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += s_red[i];
  return t;
}

// block (256 threads) <-> row.  Masks x in place (:78-79), arg-max with first-index ties (:82), softmax statistics.
__global__ void __launch_bounds__(256)
vq_dense_row_kernel(float* __restrict__ x, int64_t M, int V, int64_t ldx, MaskedCols mc, const float* __restrict__ tau_ptr,
                    int training, int64_t* __restrict__ idx_out, float* __restrict__ row_stats,
                    float* __restrict__ code_hist, float* __restrict__ subword_prob) {
  __shared__ float s_red[8];
  __shared__ int s_idx[8];
  const int64_t m = blockIdx.x;
  float* row = x + m * ldx;
  const int tid = threadIdx.x;
  if (tid < mc.n && mc.col[tid] >= 0 && mc.col[tid] < V) row[mc.col[tid]] = -INFINITY;
  __syncthreads();
  // pass A: maximum and its first index
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int v = tid; v < V; v += 256) {
    const float c = row[v];
    if (c > mx) { mx = c; mi = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
  }
  if ((tid & 31) == 0) { s_red[tid >> 5] = mx; s_idx[tid >> 5] = mi; }
  __syncthreads();
  mx = s_red[0]; mi = s_idx[0];
#pragma unroll
  for (int i = 1; i < 8; ++i)
    if (s_red[i] > mx || (s_red[i] == mx && s_idx[i] < mi)) { mx = s_red[i]; mi = s_idx[i]; }
  if (mi == 0x7fffffff) mi = 0;  // all -inf / NaN row: torch returns index 0
  const float tau = *tau_ptr;
  // pass B: normalisers at temperature 1 and tau
  float z1 = 0.f, zt = 0.f;
  for (int v = tid; v < V; v += 256) {
    const float c = row[v];
    z1 += expf(c - mx);
    zt += expf((c - mx) / tau);
  }
  z1 = block_sum_256(z1, s_red);
  zt = block_sum_256(zt, s_red);
  const float lse1 = mx + logf(z1);
  const bool soft = (training & 3) == 3;
  // pass C: entropy exactly as the reference (:111) and the value of subword_prob (:130-139)
  float ent = 0.f;
  for (int v = tid; v < V; v += 256) {
    const float p = expf(row[v] - lse1);
    ent -= p * logf(p + 1e-9f);
    // value of subword_prob: the one-hot (hard = True: `hard + p - p.detach()`, and eval), or, training with hard = False
    // (`training` bit 1), softmax(x / tau) itself (:130-131)
    if (subword_prob)
      subword_prob[m * (int64_t)V + v] = soft ? expf((row[v] - mx) / tau) / zt : (v == mi ? 1.f : 0.f);
  }
  ent = block_sum_256(ent, s_red);
  if (tid == 0) {
    idx_out[m] = mi;
    atomicAdd(&code_hist[mi], 1.0f);
    float* rs = row_stats + m * 4;
    rs[0] = lse1;
    rs[1] = mx / tau + logf(zt);
    rs[2] = ent;
    rs[3] = 0.f;
  }
}

// thread <-> column: avg_probs[v] = mean_m exp(x[m,v] - lse1[m])   (deterministic column sums)
__global__ void vq_dense_colsum_kernel(const float* __restrict__ x, int64_t M, int V, int64_t ldx,
                                       const float* __restrict__ row_stats, float* __restrict__ avg_probs) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  float s = 0.f;
  for (int64_t m = 0; m < M; ++m) s += expf(x[m * ldx + v] - row_stats[m * 4]);
  avg_probs[v] = s / (float)M;
}

// block <-> row: g_x = p_tau (g_p - <p_tau, g_p>) / tau ; g_tau partial
__global__ void __launch_bounds__(256)
vq_dense_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g_p, int V, int64_t ldx, int64_t ldg,
                    const float* __restrict__ row_stats, const float* __restrict__ tau_ptr, float* __restrict__ g_x,
                    float* __restrict__ g_tau) {
  __shared__ float s_red[8];
  const int64_t m = blockIdx.x;
  const float tau = *tau_ptr;
  const float lse_t = row_stats[m * 4 + 1];
  const float* row = x + m * ldx;
  const float* gp = g_p + m * ldg;
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) s += expf(row[v] / tau - lse_t) * gp[v];
  s = block_sum_256(s, s_red);
  float gt = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) {
    const float c = row[v];
    const float g = expf(c / tau - lse_t) * (gp[v] - s) / tau;
    g_x[m * (int64_t)V + v] = g;
    if (c > -INFINITY) gt -= g * c / tau;
  }
  if (g_tau) {
    gt = block_sum_256(gt, s_red);
    if (threadIdx.x == 0) atomicAdd(g_tau, gt);
  }
}

}  // namespace scp
#include "scp_vq_pipe.cuh"
namespace scp {

// =====================================================================================================================
// host side
// =====================================================================================================================
static MaskedCols make_masked(const int32_t* cols, int n) {
  MaskedCols mc{};
  mc.n = n;
  for (int i = 0; i < SCP_MAX_MASKED; ++i) mc.col[i] = i < n ? cols[i] : -1;
  return mc;
}

struct VqFwdWs {
  float* chunk_max;
  float* group_max;
  float* partials;
  float* lse1_l2;
  float* metric_part;
  float* avg_part;
  __half* e16;   // (Mp, Vp) fp16 e^c written by sweep 1 for the column sums (null in the two-sweep mode)
  size_t total;
  int n_chunks, n_groups;
};
// CTA pairs (cta_group::2) need at least two row tiles; single-tile problems run the one-CTA kernel.
static bool vq_use_pair(int m_tiles) { return m_tiles >= 2; }
static int vq_m_ctas(int m_tiles) { return vq_use_pair(m_tiles) ? 2 * (int)ceil_div(m_tiles, 2) : m_tiles; }

// avg_probs = mean_m softmax(c)[m,:] needs the row normalisers, i.e. a second pass over the logits.  Default: sweep 1
// also writes e^c as fp16 (Mp x Vp x 2 bytes of workspace) and vq_colsum_kernel reduces it in one HBM pass;
// SCP_VQ_COLSUM=0 selects the older second tensor-core sweep (transposed product, no logit scratch).
static bool vq_colsum_enabled() {
  static const bool on = [] { const char* e = getenv("SCP_VQ_COLSUM"); return !(e && e[0] == '0'); }();
  return on;
}

// bring-up switch: SCP_VQ_RESIDENT=0 selects the streaming-X kernels (A/B measurement of the resident-X mode)
static bool vq_resident_enabled() {
  static const bool on = [] { const char* e = getenv("SCP_VQ_RESIDENT"); return !(e && e[0] == '0'); }();
  return on;
}

static int vq_sweep1_groups(int64_t Mp, int64_t Vp) {
  const int m_tiles = (int)(Mp / tc::kTileM);
  const int n_tiles = (int)(Vp / kVqBN);
  int g = kNumSMs / vq_m_ctas(m_tiles);
  if (g < 1) g = 1;
  if (g > n_tiles) g = n_tiles;
  return g;
}
static VqFwdWs vq_fwd_ws(void* base, int64_t M, int64_t V, bool with_scratch = true) {
  const int64_t Mp = round_up(M, tc::kTileM), Vp = scp_vq_padded_vocab(V);
  VqFwdWs w{};
  w.n_chunks = (int)(Vp / 32);
  w.n_groups = vq_sweep1_groups(Mp, Vp);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  w.chunk_max = static_cast<float*>(take((size_t)Mp * w.n_chunks * 4));
  w.group_max = static_cast<float*>(take((size_t)Mp * w.n_chunks * 4 * 4));  // four 8-column group maxima per chunk
  w.partials = static_cast<float*>(take((size_t)Mp * w.n_groups * 4 * 16));  // up to 4 epilogue warps per lane quadrant
  w.lse1_l2 = static_cast<float*>(take((size_t)round_up(M, 256) * 4));
  w.metric_part = static_cast<float*>(take((size_t)(kMetricBlocks * 3 + 1) * 4));  // + the ticket counter
  w.avg_part = static_cast<float*>(take((size_t)kColsumRowSplit * Vp * 4));        // row-range partials of the column sums
  // (scp_vq_fwd_save: the caller's saved_probs buffer is the scratch)
  w.e16 = with_scratch && vq_colsum_enabled() ? static_cast<__half*>(take((size_t)Mp * Vp * 2)) : nullptr;
  w.total = off;
  return w;
}

struct VqBwdWs {
  __half* g_hat;
  float* g_aux;
  __half* pq;
  float* partials;
  float* uw;
  size_t total;
  int n_groups, k_splits, bn_out;
  // producer/consumer pipeline (scp_vq_pipe.cuh)
  int pipe_mode;  // 0 = two-kernel path, 1 = fused pipeline (ring in L2), 2 = the two roles as separate launches
  int NP, ring, uw_slots, sa, sc, MT, NVT;
  unsigned int* flags;  // ready[NP][ring] | done[NP]
  size_t pipe_smem;
};
static int vq_out_bn(int64_t D) { return D % 256 == 0 ? 256 : (D % 128 == 0 ? 128 : 64); }

// SCP_VQ_BWD_PIPE: 0 (default) the sweep-3 + gemm_out kernels; 1 the producer/consumer pipeline of scp_vq_pipe.cuh
// (D = 128 / 256 / 512: it keeps a 128 x D fp16 tile resident), 2 the same two roles as separate launches through a
// full-size scratch.  Measured on B200 (M = 2048, V = 49408, D = 512, ncu launch times): two-kernel path 385 us with
// 872 MB of DRAM traffic and a 460 MB workspace; pipeline 487 us with the (M,V) matrices confined to a 19 MB ring in L2
// and a 56 MB workspace.  Both roles are bound by the bytes a CTA can keep in flight towards L2 (96 KB ring next to the
// 128 KB resident operand), so giving each role half of the SMs halves its rate: the pipeline trades 25 % of time for
// ~8x less memory and is therefore opt-in (profiles/README.md, round 2).
static int vq_bwd_pipe_mode(int64_t D) {
  static const int env = [] { const char* e = getenv("SCP_VQ_BWD_PIPE"); return e ? atoi(e) : 0; }();
  if (!(D == 128 || D == 256 || D == 512)) return 0;
  return env;
}
static int vq_pipe_ring() {
  static const int r = [] { const char* e = getenv("SCP_VQ_BWD_RING"); const int v = e ? atoi(e) : 4; return v < 2 ? 2 : (v > 64 ? 64 : v); }();
  return r;
}

// the saved-numerator path of scp_vq_bwd_saved is taken iff ... (one predicate for the launch and the workspace size)
static bool vq_saved_path_ok(int64_t M, int64_t D, bool want_tau) {
  const int64_t Mp = round_up(M, tc::kTileM);
  return !want_tau && !vq_bwd_pipe_mode(D) && vq_use_pair((int)(Mp / tc::kTileM)) && vq_saved_enabled();
}
// sweep T keeps the ghat tile resident when it fits next to a 5-stage ring (D <= 512) and streams it otherwise
static bool vq_sweep_t_resident(int64_t D) {
  return vq_resident_enabled() &&
         tc::resident_smem_bytes<kVqBN, 1, 5, SweepTEpi, tc::MC_PAIR>((int)(D / tc::kChunkK)) <= tc::kMaxDynSmem;
}
static VqBwdWs vq_bwd_ws(void* base, int64_t M, int64_t V, int64_t D, bool saved = false) {
  const int64_t Mp = round_up(M, tc::kTileM), Vp = scp_vq_padded_vocab(V);
  VqBwdWs w{};
  w.pipe_mode = vq_bwd_pipe_mode(D);
  if (w.pipe_mode) {
    w.MT = (int)(Mp / tc::kTileM);
    w.NVT = (int)(Vp / pipe::kStepV);
    const long long total = (long long)w.MT * w.NVT;
    const int max_pipes = w.pipe_mode == 1 ? kNumSMs / 4 : kNumSMs / 2;
    w.NP = (int)std::max<long long>(1, std::min<long long>(max_pipes, total));
    w.ring = vq_pipe_ring();
    w.uw_slots = 1;
    for (int mt = 0; mt < w.MT; ++mt) {
      const int qa = pipe::pipe_of_step((long long)mt * w.NVT, total, w.NP);
      const int qb = pipe::pipe_of_step((long long)(mt + 1) * w.NVT - 1, total, w.NP);
      w.uw_slots = std::max(w.uw_slots, qb - qa + 1);
    }
    const int KC = (int)(D / tc::kChunkK);
    const int smem_budget = tc::kMaxDynSmem - 1024 - pipe::kBarBytes;
    const int prod_fixed = 1024;  // per-keyword vectors of the producer epilogue
    w.sa = std::min(pipe::kMaxStages, (smem_budget - KC * tc::kXTileBytes - prod_fixed) / tc::kXTileBytes);
    const int c_stage = tc::kXTileBytes + (int)D * 64;
    w.sc = std::min(pipe::kMaxStages, smem_budget / c_stage);
    const size_t prod = (size_t)KC * tc::kXTileBytes + (size_t)w.sa * tc::kXTileBytes + prod_fixed;
    const size_t cons = (size_t)w.sc * c_stage;
    w.pipe_smem = 1024 + pipe::kBarBytes + std::max(prod, cons);
    size_t off = 0;
    auto take = [&](size_t bytes) {
      void* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
      off += (bytes + 255) & ~size_t(255);
      return p;
    };
    w.g_hat = static_cast<__half*>(take((size_t)Mp * D * 2));
    w.g_aux = static_cast<float*>(take((size_t)Mp * 2 * 4));
    w.flags = static_cast<unsigned int*>(take((size_t)w.NP * (w.ring + 1) * 4));
    const size_t slots = w.pipe_mode == 1 ? (size_t)w.NP * w.ring : (size_t)total;
    w.pq = static_cast<__half*>(take(slots * pipe::kSlotHalfs * 2));
    w.partials = static_cast<float*>(take((size_t)Mp * w.uw_slots * 8 * 16));
    w.uw = static_cast<float*>(take((size_t)w.uw_slots * 2 * Mp * D * 4));
    w.total = off;
    return w;
  }
  const int m_tiles = (int)(Mp / tc::kTileM);
  const int n_tiles3 = (int)(Vp / 128);
  w.n_groups = std::max(1, std::min(n_tiles3, kNumSMs / vq_m_ctas(m_tiles)));
  w.bn_out = vq_out_bn(D);
  const int out_items = vq_m_ctas(m_tiles) * (int)(D / w.bn_out);
  const int k_chunks = (int)(Vp / tc::kChunkK);
  w.k_splits = std::max(1, std::min(k_chunks, kNumSMs / out_items));
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  w.g_hat = static_cast<__half*>(take((size_t)Mp * D * 2));
  w.g_aux = static_cast<float*>(take((size_t)Mp * 2 * 4));
  w.pq = static_cast<__half*>(take((size_t)(saved ? 1 : 2) * Mp * Vp * 2));  // saved: Q'' only, P'' is the forward's buffer
  w.partials = static_cast<float*>(take((size_t)Mp * w.n_groups * 2 * 16));
  w.uw = static_cast<float*>(take((size_t)w.k_splits * 2 * Mp * D * 4));
  w.total = off;
  return w;
}

template <bool SAVE>
static int launch_sweep1(const GemmMaps& maps, const Sched& sc, const VqFwdWs& ws, const float* tau, __half* e16,
                         int64_t Vp, int64_t V, const MaskedCols& mc, bool pair, cudaStream_t s) {
  using Epi = Sweep1EpiT<SAVE>;
  typename Epi::Params ep{};
  ep.chunk_max = ws.chunk_max;
  ep.group_max = ws.group_max;
  ep.partials = ws.partials;
  ep.tau = tau;
  ep.e16 = e16;
  ep.ldE = Vp;
  static const int s1_dbg = ablation_env("SCP_VQ_S1_DBG");
  ep.dbg = s1_dbg;
  ep.n_chunks = ws.n_chunks;
  ep.n_groups = ws.n_groups;
  ep.V = (int)V;
  ep.mc = mc;
  // resident keyword tile (128 x D fp16) when it fits next to a 6-stage ring of table half-tiles (D <= 512)
  const bool xres = pair && vq_resident_enabled() &&
                    tc::resident_smem_bytes<kVqBN, 1, 6, Epi, tc::MC_PAIR>(sc.k_chunks) <= tc::kMaxDynSmem;
  if (xres) return tc::launch_stream_gemm<kVqBN, 1, 6, Epi, 2, tc::MC_PAIR, true>(maps, sc, ep, s, "vq_sweep1");
  if (pair) return tc::launch_stream_gemm<kVqBN, 1, 6, Epi, 2, tc::MC_PAIR>(maps, sc, ep, s, "vq_sweep1");
  return tc::launch_stream_gemm<kVqBN, 1, 4, Epi>(maps, sc, ep, s, "vq_sweep1");
}

static int check_vq_shape(int64_t M, int64_t V, int64_t D) {
  SCP_CHECK_ARG(M > 0 && V > 0 && D > 0, "vq: non-positive shape");
  if (D % 64 != 0 || D > 1024) return fail(SCP_ERR_UNSUPPORTED, "vq: D must be a multiple of 64 (<= 1024), got %lld", (long long)D);
  if (V < 2) return fail(SCP_ERR_UNSUPPORTED, "vq: V must be >= 2");
  return SCP_OK;
}

}  // namespace scp

using namespace scp;

extern "C" int64_t scp_vq_padded_vocab(int64_t V) { return round_up(V, kVqPad); }

extern "C" int scp_vq_prepare_table(const float* table, int64_t V, int64_t D, void* table_hat, void* table_hat_t,
                                    float* table_norm, float* table_mean, scp_stream_t stream) {
  int rc = check_device_arch();
  if (rc) return rc;
  rc = check_vq_shape(1, V, D);
  if (rc) return rc;
  SCP_CHECK_ARG(table && table_hat && table_hat_t && table_norm && table_mean, "vq_prepare_table: null pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t Vp = scp_vq_padded_vocab(V);
  // table_mean has D+1 entries: [D] temporarily holds the bit pattern of the maximum norm
  if (cudaMemsetAsync(table_mean, 0, (size_t)(D + 1) * 4, s) != cudaSuccess) return fail(SCP_ERR_CUDA, "memset");
  vq_table_normalize_kernel<<<(unsigned)ceil_div(Vp, 8), 256, 0, s>>>(
      table, V, Vp, (int)D, reinterpret_cast<__half*>(table_hat), table_norm,
      reinterpret_cast<unsigned int*>(table_mean + D));
  SCP_CUDA_LAUNCH_CHECK("vq_table_normalize");
  dim3 tb(32, 8), tg((unsigned)ceil_div(Vp, 32), (unsigned)ceil_div(D, 32));
  transpose_f16_kernel<<<tg, tb, 0, s>>>(reinterpret_cast<const __half*>(table_hat), Vp, (int)D,
                                         reinterpret_cast<__half*>(table_hat_t), Vp);
  SCP_CUDA_LAUNCH_CHECK("transpose_f16");
  dim3 mb(32, 8), mg((unsigned)ceil_div(D, 32), 64);
  vq_table_mean_kernel<<<mg, mb, 0, s>>>(table, V, (int)D, table_mean);
  SCP_CUDA_LAUNCH_CHECK("vq_table_mean");
  return SCP_OK;
}

extern "C" size_t scp_vq_fwd_workspace_bytes(int64_t M, int64_t V, int64_t) { return vq_fwd_ws(nullptr, M, V).total; }
extern "C" size_t scp_vq_fwd_save_workspace_bytes(int64_t M, int64_t V, int64_t) {
  return vq_fwd_ws(nullptr, M, V, false).total;
}

extern "C" size_t scp_vq_saved_probs_bytes(int64_t M, int64_t V) {
  return (size_t)round_up(M, tc::kTileM) * (size_t)scp_vq_padded_vocab(V) * 2;
}

extern "C" int scp_vq_fwd(const float* kw, int64_t M, int64_t K, int64_t V, int64_t D, const void* table_hat,
                          const float* table_norm, const float* table, const int32_t* masked_cols, int n_masked,
                          const float* tau, int64_t* idx, float* keywords, float* row_stats, float* code_hist,
                          float* avg_probs, float* metrics, void* kw_hat, void* workspace, size_t workspace_bytes,
                          scp_stream_t stream) {
  return scp_vq_fwd_save(kw, M, K, V, D, table_hat, table_norm, table, masked_cols, n_masked, tau, idx, keywords, row_stats,
                         code_hist, avg_probs, metrics, kw_hat, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int scp_vq_fwd_save(const float* kw, int64_t M, int64_t K, int64_t V, int64_t D, const void* table_hat,
                               const float* table_norm, const float* table, const int32_t* masked_cols, int n_masked,
                               const float* tau, int64_t* idx, float* keywords, float* row_stats, float* code_hist,
                               float* avg_probs, float* metrics, void* kw_hat, void* saved_probs, void* workspace,
                               size_t workspace_bytes, scp_stream_t stream) {
  int rc = check_device_arch();
  if (rc) return rc;
  rc = check_vq_shape(M, V, D);
  if (rc) return rc;
  SCP_CHECK_ARG(kw && table_hat && table_norm && table && tau && idx && keywords && row_stats && code_hist && metrics &&
                    kw_hat && workspace,
                "vq_fwd: null pointer");
  SCP_CHECK_ARG(K > 0 && M % K == 0, "vq_fwd: M=%lld is not a multiple of K=%lld", (long long)M, (long long)K);
  SCP_CHECK_ARG(n_masked >= 0 && n_masked <= SCP_MAX_MASKED && (n_masked == 0 || masked_cols), "vq_fwd: masked cols");
  const VqFwdWs ws = vq_fwd_ws(workspace, M, V, saved_probs == nullptr);
  if (workspace_bytes < ws.total) return fail(SCP_ERR_WORKSPACE, "vq_fwd: workspace %zu < %zu", workspace_bytes, ws.total);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t Mp = round_up(M, tc::kTileM), Vp = scp_vq_padded_vocab(V);
  const MaskedCols mc = make_masked(masked_cols, n_masked);

  vq_prep_kw_kernel<<<(unsigned)ceil_div(Mp, 8), 256, 0, s>>>(kw, M, Mp, (int)D, reinterpret_cast<__half*>(kw_hat),
                                                              row_stats, code_hist, Vp, ws.lse1_l2, round_up(M, 256),
                                                              reinterpret_cast<unsigned int*>(ws.metric_part + kMetricBlocks * 2));
  SCP_CUDA_LAUNCH_CHECK("vq_prep_kw");

  // ---- sweep 1
  const int s1_parts = 2;  // partial-statistics slots per V group = epilogue warps per lane quadrant of the launched kernel
  {
    GemmMaps maps{};
    if ((rc = tc::make_tmap_f16(&maps.x[0], kw_hat, Mp, D, D, tc::kTileM))) return rc;
    maps.x[1] = maps.x[0];
    const bool pair = vq_use_pair((int)(Mp / tc::kTileM));
    if ((rc = tc::make_tmap_f16(&maps.y, table_hat, Vp, D, D, pair ? kVqBN / 2 : kVqBN))) return rc;
    Sched sc{};
    sc.m_tiles = (int)(Mp / tc::kTileM);
    sc.n_tiles = (int)(Vp / kVqBN);
    sc.n_groups = ws.n_groups;
    sc.k_chunks = (int)(D / tc::kChunkK);
    sc.k_splits = 1;
    sc.m_half = sc.m_tiles;
    sc.n_upper_off = 0;
    // scp_vq_fwd_save: the caller's buffer takes the place of the workspace scratch and receives P'' (always written: the
    // backward pass needs it even when prob_perplexity is not wanted)
    __half* e16 = saved_probs ? static_cast<__half*>(saved_probs) : (avg_probs ? ws.e16 : nullptr);
    rc = saved_probs ? launch_sweep1<true>(maps, sc, ws, tau, e16, Vp, V, mc, pair, s)
                     : launch_sweep1<false>(maps, sc, ws, tau, e16, Vp, V, mc, pair, s);
    if (rc) return rc;
  }
  // the (M,V) fp16 scratch sweep 1 has written: e^c in the workspace, or P'' in the caller's buffer (scp_vq_fwd_save)
  const __half* scratch = saved_probs ? static_cast<const __half*>(saved_probs) : (avg_probs ? ws.e16 : nullptr);
  // ---- exact arg-max, statistics, gather
  const size_t sel_smem = (size_t)4 * ws.n_chunks * sizeof(float);  // one row of chunk maxima per warp
  if (sel_smem > 48 * 1024)
    return fail(SCP_ERR_UNSUPPORTED, "vq_fwd: V=%lld exceeds the arg-max kernel's shared-memory scan (V <= 98304)", (long long)V);
  // The arg-max phase (latency-bound, light on every unit) and sweep 2 (tensor-bound) are independent once the row
  // statistics exist: the statistics run on the caller's stream, the arg-max is forked onto the library's helper stream
  // and joined before the metrics -- it then hides under sweep 2 instead of preceding it.  SCP_VQ_OVERLAP=0 disables.
  static const bool overlap_on = [] { const char* e = getenv("SCP_VQ_OVERLAP"); return !(e && e[0] == '0'); }();
#define SCP_SELECT_LAUNCH(NVV, STREAM, PHASES)                                                                         \
  vq_select_kernel<NVV><<<(unsigned)ceil_div(M, 4), 128, sel_smem, STREAM>>>(                                          \
      kw, table, table_norm, reinterpret_cast<const __half*>(table_hat), reinterpret_cast<const __half*>(kw_hat), M,   \
      (int)V, (int)D, ws.chunk_max, ws.group_max, ws.n_chunks, ws.partials, s1_parts * ws.n_groups, tau, mc, idx, keywords, \
      row_stats, code_hist, ws.lse1_l2, PHASES, scratch, Vp, saved_probs ? 1 : 0)
#define SCP_SELECT(STREAM, PHASES)                                                                                     \
  do {                                                                                                                 \
    if (D <= 128) SCP_SELECT_LAUNCH(1, STREAM, PHASES);                                                                \
    else if (D <= 256) SCP_SELECT_LAUNCH(2, STREAM, PHASES);                                                           \
    else if (D <= 512) SCP_SELECT_LAUNCH(4, STREAM, PHASES);                                                           \
    else if (D <= 768) SCP_SELECT_LAUNCH(6, STREAM, PHASES);                                                           \
    else SCP_SELECT_LAUNCH(8, STREAM, PHASES);                                                                         \
  } while (0)
  if (D > 1024) return fail(SCP_ERR_UNSUPPORTED, "vq_fwd: D > 1024 is not supported by the exact arg-max kernel (got %lld)", (long long)D);
  bool forked = false;
  if (avg_probs && overlap_on) {
    SCP_SELECT(s, 1);  // row statistics: sweep 2 needs the normalisers
    SCP_CUDA_LAUNCH_CHECK("vq_select(stats)");
    cudaStream_t side = fork_to_side(s);
    if (side) {
      SCP_SELECT(side, 2);  // arg-max + gather + histogram, concurrently with sweep 2
      SCP_CUDA_LAUNCH_CHECK("vq_select(argmax)");
      forked = true;
    } else {
      SCP_SELECT(s, 2);
      SCP_CUDA_LAUNCH_CHECK("vq_select(argmax)");
    }
  } else {
    SCP_SELECT(s, 3);
    SCP_CUDA_LAUNCH_CHECK("vq_select");
  }
#undef SCP_SELECT
#undef SCP_SELECT_LAUNCH
  // ---- column sums -- skipped when the caller does not want prob_perplexity
  const int n_part = M >= 1024 ? kColsumRowSplit : 1;
  if (avg_probs && scratch) {
    const int64_t rows_per_block = round_up(ceil_div(M, n_part), 64);
    float* dst = n_part > 1 ? ws.avg_part : avg_probs;
    const dim3 grid((unsigned)(Vp / kColsumCols), (unsigned)n_part);
    if (saved_probs)
      vq_colsum_kernel<true><<<grid, 256, 0, s>>>(scratch, Vp, ws.lse1_l2, M, (int)V, 1.0f / (float)M, mc, tau, dst, rows_per_block);
    else
      vq_colsum_kernel<false><<<grid, 256, 0, s>>>(scratch, Vp, ws.lse1_l2, M, (int)V, 1.0f / (float)M, mc, tau, dst, rows_per_block);
    if (cudaGetLastError() != cudaSuccess) {
      if (forked) join_from_side(s);  // never leave the helper stream un-joined (stream capture would be invalidated)
      return fail(SCP_ERR_CUDA, "vq_colsum launch failed");
    }
    count_launch();
  } else if (avg_probs) {  // two-sweep mode (SCP_VQ_COLSUM=0)
    const int64_t Mp2 = round_up(M, 256);
    GemmMaps maps{};
    if ((rc = tc::make_tmap_f16(&maps.x[0], table_hat, Vp, D, D, tc::kTileM))) return rc;
    maps.x[1] = maps.x[0];
    if ((rc = tc::make_tmap_f16(&maps.y, kw_hat, Mp, D, D, 128))) return rc;  // pair: half of the 256-row Y tile
    Sched sc{};
    sc.m_tiles = (int)(Vp / tc::kTileM);
    sc.n_tiles = (int)(Mp2 / 256);
    sc.n_groups = 1;
    sc.k_chunks = (int)(D / tc::kChunkK);
    sc.k_splits = 1;
    sc.m_half = sc.m_tiles;
    sc.n_upper_off = 0;
    Sweep2Epi::Params ep{};
    ep.lse1_l2 = ws.lse1_l2;
    ep.avg_probs = avg_probs;
    ep.inv_m = 1.0f / (float)M;
    ep.V = (int)V;
    ep.mc = mc;
    const bool xres = vq_resident_enabled() &&
                      tc::resident_smem_bytes<256, 1, 5, Sweep2Epi, tc::MC_PAIR>(sc.k_chunks) <= tc::kMaxDynSmem;
    if (xres) rc = tc::launch_stream_gemm<256, 1, 5, Sweep2Epi, 2, tc::MC_PAIR, true>(maps, sc, ep, s, "vq_sweep2");
    else rc = tc::launch_stream_gemm<256, 1, 6, Sweep2Epi, 2, tc::MC_PAIR>(maps, sc, ep, s, "vq_sweep2");
    if (rc) {
      if (forked) join_from_side(s);  // never leave the helper stream un-joined (stream capture would be invalidated)
      return rc;
    }
  }
  if (forked && (rc = join_from_side(s))) return rc;  // the metrics need the code histogram of the arg-max phase
  const bool parts = avg_probs && scratch && n_part > 1;
  vq_metrics_kernel<<<kMetricBlocks, 256, 0, s>>>(code_hist, avg_probs, row_stats, M, (int)K, (int)V, ws.metric_part,
                                                  reinterpret_cast<unsigned int*>(ws.metric_part + kMetricBlocks * 2),
                                                  metrics, parts ? ws.avg_part : nullptr, parts ? n_part : 0, (int)Vp);
  SCP_CUDA_LAUNCH_CHECK("vq_metrics");
  return SCP_OK;
}

extern "C" size_t scp_vq_bwd_workspace_bytes(int64_t M, int64_t V, int64_t D) {
  return vq_bwd_ws(nullptr, M, V, D).total;
}
extern "C" int scp_vq_bwd_saved_available(int64_t M, int64_t V, int64_t D) {
  (void)V;
  return vq_saved_path_ok(M, D, false) ? 1 : 0;
}
extern "C" size_t scp_vq_bwd_saved_workspace_bytes(int64_t M, int64_t V, int64_t D, int want_tau) {
  return vq_bwd_ws(nullptr, M, V, D, vq_saved_path_ok(M, D, want_tau != 0)).total;
}

template <int BN>
static int launch_gemm_out(const GemmMaps& maps, const Sched& sc, const StoreEpi<2>::Params& ep, bool pair, cudaStream_t s) {
  constexpr int kStages = BN == 256 ? 3 : 4;       // one CTA: 32 KB of X + BN*128 B of Y per stage
  constexpr int kPairStages = BN == 256 ? 4 : 5;   // pair: 32 KB of X + BN*64 B of Y per stage
  if (pair) return tc::launch_stream_gemm<BN, 2, kPairStages, StoreEpi<2>, 2, tc::MC_PAIR>(maps, sc, ep, s, "vq_gemm_out");
  return tc::launch_stream_gemm<BN, 2, kStages, StoreEpi<2>>(maps, sc, ep, s, "vq_gemm_out");
}

extern "C" int scp_vq_bwd(const float* g_keywords, const float* kw, int64_t M, int64_t V, int64_t D,
                          const void* kw_hat, const void* table_hat, const void* table_hat_t,
                          const float* table_norm, const float* table_mean, const float* row_stats,
                          const int32_t* masked_cols, int n_masked, const float* tau, float* g_kw, float* g_tau,
                          void* workspace, size_t workspace_bytes, scp_stream_t stream) {
  return scp_vq_bwd_saved(g_keywords, kw, M, V, D, kw_hat, table_hat, table_hat_t, table_norm, table_mean, row_stats,
                          masked_cols, n_masked, tau, nullptr, g_kw, g_tau, workspace, workspace_bytes, stream);
}

extern "C" int scp_vq_bwd_saved(const float* g_keywords, const float* kw, int64_t M, int64_t V, int64_t D,
                                const void* kw_hat, const void* table_hat, const void* table_hat_t,
                                const float* table_norm, const float* table_mean, const float* row_stats,
                                const int32_t* masked_cols, int n_masked, const float* tau, const void* saved_probs,
                                float* g_kw, float* g_tau, void* workspace, size_t workspace_bytes, scp_stream_t stream) {
  int rc = check_device_arch();
  if (rc) return rc;
  rc = check_vq_shape(M, V, D);
  if (rc) return rc;
  SCP_CHECK_ARG(g_keywords && kw && kw_hat && table_hat && table_hat_t && table_norm && table_mean && row_stats &&
                    tau && g_kw && workspace,
                "vq_bwd: null pointer");
  SCP_CHECK_ARG(n_masked >= 0 && n_masked <= SCP_MAX_MASKED && (n_masked == 0 || masked_cols), "vq_bwd: masked cols");
  const bool use_saved = saved_probs && vq_saved_path_ok(M, D, g_tau != nullptr);
  const VqBwdWs ws = vq_bwd_ws(workspace, M, V, D, use_saved);
  if (workspace_bytes < ws.total) return fail(SCP_ERR_WORKSPACE, "vq_bwd: workspace %zu < %zu", workspace_bytes, ws.total);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t Mp = round_up(M, tc::kTileM), Vp = scp_vq_padded_vocab(V);
  const MaskedCols mc = make_masked(masked_cols, n_masked);

  vq_bwd_prep_kernel<<<(unsigned)ceil_div(Mp, 8), 256, 0, s>>>(g_keywords, M, Mp, (int)D, table_mean, ws.g_hat, ws.g_aux,
                                                               ws.pipe_mode ? ws.flags : nullptr, ws.NP * (ws.ring + 1));
  SCP_CUDA_LAUNCH_CHECK("vq_bwd_prep");
  if (ws.pipe_mode) {
    pipe::PipeMaps maps{};
    if ((rc = tc::make_tmap_f16(&maps.tab_k, table_hat, Vp, D, D, 128))) return rc;
    if ((rc = tc::make_tmap_f16(&maps.kw, kw_hat, Mp, D, D, 128))) return rc;
    if ((rc = tc::make_tmap_f16(&maps.gh, ws.g_hat, Mp, D, D, 128))) return rc;
    const int64_t slots = ws.pipe_mode == 1 ? (int64_t)ws.NP * ws.ring : (int64_t)ws.MT * ws.NVT;
    if ((rc = tc::make_tmap_f16_blocked(&maps.scr, ws.pq, slots * 512, 2, 128, 64, 2))) return rc;
    {
      const int n_mma = D > 256 ? 2 : 1;
      const int nb = (int)(D / n_mma / 2 / 64);  // 64-column table blocks one CTA stages per MMA
      if ((rc = tc::make_tmap_f16_blocked(&maps.tab_mn, table_hat, Vp, D / 64, D, 64, nb))) return rc;
    }
    pipe::PipeParams pp{};
    pp.MT = ws.MT; pp.NVT = ws.NVT; pp.NP = ws.NP; pp.KC = (int)(D / tc::kChunkK); pp.D = (int)D; pp.V = (int)V;
    pp.M = M; pp.Mp = Mp;
    pp.fused = ws.pipe_mode == 1;
    pp.ring = ws.ring; pp.sa = ws.sa; pp.sc = ws.sc; pp.uw_slots = ws.uw_slots;
    static const int pipe_dbg = ablation_env("SCP_PIPE_DEBUG");
    pp.debug = pipe_dbg;
    pp.scratch = ws.pq;
    pp.ready = ws.flags; pp.done = ws.flags + (size_t)ws.NP * ws.ring;
    pp.row_stats = row_stats; pp.g_aux = ws.g_aux; pp.table_norm = table_norm; pp.table_mean = table_mean; pp.tau = tau;
    pp.sums = ws.partials; pp.uw = ws.uw; pp.mc = mc;
    auto kern = g_tau ? pipe::vq_bwd_pipe_kernel<true> : pipe::vq_bwd_pipe_kernel<false>;
    static thread_local bool configured[2] = {false, false};
    if (!configured[g_tau ? 1 : 0]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kMaxDynSmem);
      if (e != cudaSuccess) return fail(SCP_ERR_CUDA, "vq_bwd_pipe: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      configured[g_tau ? 1 : 0] = true;
    }
    if (ws.pipe_smem > (size_t)tc::kMaxDynSmem) return fail(SCP_ERR_UNSUPPORTED, "vq_bwd_pipe: %zu B of shared memory", ws.pipe_smem);
    auto launch = [&](int role) -> int {
      pp.role = role;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(pp.fused ? 4 * ws.NP : 2 * ws.NP));
      cfg.blockDim = dim3(pipe::kPipeThreads);
      cfg.dynamicSmemBytes = ws.pipe_smem;
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps, pp);
      if (e != cudaSuccess) return fail(SCP_ERR_CUDA, "vq_bwd_pipe: launch failed: %s", cudaGetErrorString(e));
      count_launch();
      return SCP_OK;
    };
    if (pp.fused) {
      if ((rc = launch(0))) return rc;
    } else {
      if ((rc = launch(0))) return rc;
      if ((rc = launch(1))) return rc;
    }
    if (g_tau && cudaMemsetAsync(g_tau, 0, 4, s) != cudaSuccess) return fail(SCP_ERR_CUDA, "memset g_tau");
    pipe::vq_bwd_pipe_finalize_kernel<<<(unsigned)M, (unsigned)(D / 4), 0, s>>>(
        ws.uw, ws.uw_slots, ws.NVT, ws.MT, ws.NP, M, Mp, (int)D, ws.partials, ws.g_aux, kw, row_stats, table_mean, tau, g_kw,
        g_tau);
    SCP_CUDA_LAUNCH_CHECK("vq_bwd_pipe_finalize");
    return SCP_OK;
  }
  // ---- saved-numerator path: T' = ghat . Ehat^T only (one N = 256 accumulator, resident ghat), Q'' = P''(T' - s0) from the
  //      forward's P''; the output GEMM reads P'' straight from the forward's buffer.  4 M V D FLOP less than the recompute
  //      path below, no exponentials, half the scratch.  A learnable temperature (g_tau) needs the logits: recompute path.
  const __half* p_src = nullptr;  // W operand of the output GEMM (default: second half of the scratch)
  int fin_groups = 2 * ws.n_groups;
  if (use_saved) {
    GemmMaps maps{};
    if ((rc = tc::make_tmap_f16(&maps.x[0], ws.g_hat, Mp, D, D, tc::kTileM))) return rc;
    maps.x[1] = maps.x[0];
    if ((rc = tc::make_tmap_f16(&maps.y, table_hat, Vp, D, D, kVqBN / 2))) return rc;
    Sched sc{};
    sc.m_tiles = (int)(Mp / tc::kTileM);
    sc.n_tiles = (int)(Vp / kVqBN);
    sc.n_groups = vq_sweep1_groups(Mp, Vp);
    sc.k_chunks = (int)(D / tc::kChunkK);
    sc.k_splits = 1;
    sc.m_half = sc.m_tiles;
    sc.n_upper_off = 0;
    SweepTEpi::Params ep{};
    ep.p16 = static_cast<const __half*>(saved_probs);
    ep.q16 = ws.pq;
    ep.ld = Vp;
    ep.g_aux = ws.g_aux;
    ep.table_norm = table_norm;
    ep.table_mean = table_mean;
    ep.partials = ws.partials;
    ep.n_groups = sc.n_groups;
    ep.D = (int)D;
    static const int st_dbg = ablation_env("SCP_VQ_ST_DBG");
    ep.dbg = st_dbg;
    rc = vq_sweep_t_resident(D) ? tc::launch_stream_gemm<kVqBN, 1, 5, SweepTEpi, 2, tc::MC_PAIR, 1>(maps, sc, ep, s, "vq_sweep_t")
                                : tc::launch_stream_gemm<kVqBN, 1, 5, SweepTEpi, 2, tc::MC_PAIR>(maps, sc, ep, s, "vq_sweep_t");
    if (rc) return rc;
    p_src = static_cast<const __half*>(saved_probs);
    fin_groups = 2 * sc.n_groups;
  }
  // ---- sweep 3: P~, Q~ and row sums
  if (!use_saved) {
    GemmMaps maps{};
    const bool pair = vq_use_pair((int)(Mp / tc::kTileM));
    if ((rc = tc::make_tmap_f16(&maps.x[0], kw_hat, Mp, D, D, tc::kTileM))) return rc;
    if ((rc = tc::make_tmap_f16(&maps.x[1], ws.g_hat, Mp, D, D, tc::kTileM))) return rc;
    if ((rc = tc::make_tmap_f16(&maps.y, table_hat, Vp, D, D, pair ? 64 : 128))) return rc;
    Sched sc{};
    sc.m_tiles = (int)(Mp / tc::kTileM);
    sc.n_tiles = (int)(Vp / 128);
    sc.n_groups = ws.n_groups;
    sc.k_chunks = (int)(D / tc::kChunkK);
    sc.k_splits = 1;
    sc.m_half = sc.m_tiles;
    sc.n_upper_off = 0;
    Sweep3Epi::Params ep{};
    ep.row_stats = row_stats;
    ep.g_aux = ws.g_aux;
    ep.table_norm = table_norm;
    ep.table_mean = table_mean;
    ep.tau = tau;
    ep.pq = ws.pq;
    ep.partials = ws.partials;
    ep.M = M; ep.Mp = Mp; ep.Vp = Vp;
    ep.n_groups = ws.n_groups;
    ep.V = (int)V;
    ep.D = (int)D;
    ep.want_tau = g_tau != nullptr;
    ep.mc = mc;
    // optional (SCP_VQ_S3_RES=1): keep the keyword operand resident, stream only ghat and the table (24 KB stages, 3 of them)
    static const bool s3_res = [] { const char* e = getenv("SCP_VQ_S3_RES"); return e && e[0] == '1'; }();
    if (pair && s3_res && tc::resident_smem_bytes<128, 2, 3, Sweep3Epi, tc::MC_PAIR, 1>(sc.k_chunks) <= tc::kMaxDynSmem)
      rc = tc::launch_stream_gemm<128, 2, 3, Sweep3Epi, 2, tc::MC_PAIR, 1>(maps, sc, ep, s, "vq_sweep3");
    else if (pair) rc = tc::launch_stream_gemm<128, 2, 5, Sweep3Epi, 2, tc::MC_PAIR>(maps, sc, ep, s, "vq_sweep3");
    else rc = tc::launch_stream_gemm<128, 2, 3, Sweep3Epi>(maps, sc, ep, s, "vq_sweep3");
    if (rc) return rc;
  }
  // ---- U = Q~ Ehat, W = P~ Ehat   (K = Vp, split-K partials)
  {
    GemmMaps maps{};
    if ((rc = tc::make_tmap_f16(&maps.x[0], ws.pq, Mp, Vp, Vp, tc::kTileM))) return rc;
    if ((rc = tc::make_tmap_f16(&maps.x[1], p_src ? p_src : ws.pq + Mp * Vp, Mp, Vp, Vp, tc::kTileM))) return rc;
    const bool pair = vq_use_pair((int)(Mp / tc::kTileM));
    if ((rc = tc::make_tmap_f16(&maps.y, table_hat_t, D, Vp, Vp, pair ? ws.bn_out / 2 : ws.bn_out))) return rc;
    Sched sc{};
    sc.m_tiles = (int)(Mp / tc::kTileM);
    sc.n_tiles = (int)(D / ws.bn_out);
    sc.n_groups = sc.n_tiles;
    sc.k_chunks = (int)(Vp / tc::kChunkK);
    sc.k_splits = ws.k_splits;
    sc.m_half = sc.m_tiles;
    sc.n_upper_off = 0;
    StoreEpi<2>::Params ep{};
    ep.out = ws.uw;
    ep.rows = Mp;
    ep.ld = (int)D;
    if (ws.bn_out == 256) rc = launch_gemm_out<256>(maps, sc, ep, pair, s);
    else if (ws.bn_out == 128) rc = launch_gemm_out<128>(maps, sc, ep, pair, s);
    else rc = launch_gemm_out<64>(maps, sc, ep, pair, s);
    if (rc) return rc;
  }
  if (g_tau && cudaMemsetAsync(g_tau, 0, 4, s) != cudaSuccess) return fail(SCP_ERR_CUDA, "memset g_tau");
  vq_bwd_finalize_kernel<<<(unsigned)M, (unsigned)(D / 4), 0, s>>>(ws.uw, ws.k_splits, M, Mp, (int)D, ws.partials,
                                                                  fin_groups, ws.g_aux, kw, row_stats, table_mean,
                                                                  tau, g_kw, g_tau);
  SCP_CUDA_LAUNCH_CHECK("vq_bwd_finalize");
  return SCP_OK;
}

extern "C" size_t scp_vq_dense_workspace_bytes(int64_t, int64_t) { return (size_t)(kMetricBlocks * 3 + 1) * 4; }

extern "C" int scp_vq_dense_fwd(float* x, int64_t M, int64_t K, int64_t V, int64_t ldx, const int32_t* masked_cols,
                                int n_masked, const float* tau, int training, int64_t* idx, float* row_stats,
                                float* code_hist, float* avg_probs, float* metrics, float* subword_prob,
                                void* workspace, size_t workspace_bytes, scp_stream_t stream) {
  SCP_CHECK_ARG(x && tau && idx && row_stats && code_hist && metrics && workspace, "vq_dense_fwd: null pointer");
  if (workspace_bytes < scp_vq_dense_workspace_bytes(M, V))
    return fail(SCP_ERR_WORKSPACE, "vq_dense_fwd: workspace %zu < %zu", workspace_bytes, scp_vq_dense_workspace_bytes(M, V));
  SCP_CHECK_ARG(M > 0 && V > 1 && K > 0 && M % K == 0 && ldx >= V, "vq_dense_fwd: bad shape");
  SCP_CHECK_ARG(V < (1ll << 31), "vq_dense_fwd: V too large");
  SCP_CHECK_ARG(n_masked >= 0 && n_masked <= SCP_MAX_MASKED && (n_masked == 0 || masked_cols), "vq_dense_fwd: masked cols");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const MaskedCols mc = make_masked(masked_cols, n_masked);
  if (cudaMemsetAsync(code_hist, 0, (size_t)V * 4, s) != cudaSuccess) return fail(SCP_ERR_CUDA, "memset code_hist");
  vq_dense_row_kernel<<<(unsigned)M, 256, 0, s>>>(x, M, (int)V, ldx, mc, tau, training, idx, row_stats, code_hist,
                                                  subword_prob);
  SCP_CUDA_LAUNCH_CHECK("vq_dense_row");
  if (avg_probs) {
    vq_dense_colsum_kernel<<<(unsigned)ceil_div(V, 128), 128, 0, s>>>(x, M, (int)V, ldx, row_stats, avg_probs);
    SCP_CUDA_LAUNCH_CHECK("vq_dense_colsum");
  }
  float* part = reinterpret_cast<float*>(workspace);
  if (cudaMemsetAsync(part + kMetricBlocks * 2, 0, 4, s) != cudaSuccess) return fail(SCP_ERR_CUDA, "memset ticket");
  vq_metrics_kernel<<<kMetricBlocks, 256, 0, s>>>(code_hist, avg_probs, row_stats, M, (int)K, (int)V, part,
                                                  reinterpret_cast<unsigned int*>(part + kMetricBlocks * 2), metrics);
  SCP_CUDA_LAUNCH_CHECK("vq_metrics");
  return SCP_OK;
}

extern "C" int scp_vq_dense_bwd(const float* x_masked, const float* g_p, int64_t M, int64_t V, int64_t ldx, int64_t ldg,
                                const float* row_stats, const float* tau, float* g_x, float* g_tau,
                                scp_stream_t stream) {
  SCP_CHECK_ARG(x_masked && g_p && row_stats && tau && g_x, "vq_dense_bwd: null pointer");
  SCP_CHECK_ARG(M > 0 && V > 1 && ldx >= V && ldg >= V, "vq_dense_bwd: bad shape");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (g_tau && cudaMemsetAsync(g_tau, 0, 4, s) != cudaSuccess) return fail(SCP_ERR_CUDA, "memset g_tau");
  vq_dense_bwd_kernel<<<(unsigned)M, 256, 0, s>>>(x_masked, g_p, (int)V, ldx, ldg, row_stats, tau, g_x, g_tau);
  SCP_CUDA_LAUNCH_CHECK("vq_dense_bwd");
  return SCP_OK;
}
