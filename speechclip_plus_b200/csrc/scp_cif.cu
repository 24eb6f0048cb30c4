// N4 -- CIF down-sampler: CIF.integrate_and_fire of the reference (avssl/module/cif.py:157-311), the step that turns the
// (B,S,C) frame sequence into the dynamic-K keyword sequence the "+" branches feed to the VQ.
//
// Reference: cumsum + floor -> per-source (left_idx, right_idx, weights), then 2 + extra_weights.max().item()
// scatter_add_ passes over (B,S,C) temporaries, and (inference) a tail pass; three host synchronisations.
// Here:
//   cif_plan      warp <-> utterance: SEQUENTIAL fp32 cumsum of alpha (the firing positions are floor(csum/threshold): the
//                 summation order decides borderline fires, so it is fixed), fired length per utterance
//   [host reads max(feat_len) once -- the output shape depends on it, exactly as in the reference]
//   cif_fire_fwd  block <-> (utterance, 512-channel slab), thread <-> 4 channels: per-source descriptors in shared memory, then ONE sequential
//                 pass over the sources; every thread keeps the accumulator of the output row being integrated in
//                 registers and writes each of the T+1 output rows exactly once (no zero-fill, no atomics, deterministic).
//                 Bytes: B*S*C*4 read + B*(T+1)*C*4 written -- HBM-bound.
//   cif_tail      inference tail handling (:246-296): up-scale / extend / erase, no host-side `if extend_mask.any()`
//   cif_fire_bwd  warp <-> source: d_input row + the two channel dots that feed d_alpha; cif_alpha_grad: reverse cumsum.
// A source s spreads alpha_s over the output rows left_idx_s .. right_idx_s:
//   right_w = fire ? csum_s - right_idx_s*thr : 0 ; left_w = alpha_s - right_w - extra*thr ; extra = max(fire_num-1,0) rows
//   in between receive thr each; indices are clipped to T (the tail row).
#include "scp_common.cuh"

namespace scp {

struct CifSrc {
  int left, right, extra;
  float left_w, right_w;
};

__device__ __forceinline__ CifSrc cif_source(const float* __restrict__ csum, const float* __restrict__ alpha, int s,
                                             float thr, int T) {
  CifSrc d;
  const float cs = csum[s];
  int r = (int)floorf(cs / thr);
  r = r < 0 ? 0 : (r > T ? T : r);
  int l = 0;
  if (s > 0) {
    l = (int)floorf(csum[s - 1] / thr);
    l = l < 0 ? 0 : (l > T ? T : l);
  }
  const int fire = r - l;
  d.left = l;
  d.right = r;
  d.extra = fire > 1 ? fire - 1 : 0;
  d.right_w = fire > 0 ? cs - (float)r * thr : 0.f;
  d.left_w = alpha[s] - d.right_w - (float)d.extra * thr;
  return d;
}

// warp <-> utterance.  dynamic smem: 4 warps * S floats
__global__ void __launch_bounds__(128)
cif_plan_kernel(const float* __restrict__ alpha, int64_t B, int S, float thr, int max_len, float* __restrict__ csum,
                int64_t* __restrict__ feat_len) {
  extern __shared__ float s_row[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * 4 + warp;
  if (b >= B) return;
  float* row = s_row + (size_t)warp * S;
  for (int s = lane; s < S; s += 32) row[s] = alpha[b * S + s];
  __syncwarp();
  if (lane == 0) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) {  // sequential on purpose
      acc += row[s];
      row[s] = acc;
    }
    int64_t n = (int64_t)floorf(acc / thr);  // cif.py:183-188
    n = n < 1 ? 1 : (n > max_len ? max_len : n);
    feat_len[b] = n;
  }
  __syncwarp();
  for (int s = lane; s < S; s += 32) csum[b * S + s] = row[s];
}

// block (128 threads) <-> (utterance b, slab of 512 channels); thread <-> 4 channels.
// The kernel is instruction-issue bound, not latency bound (two channels per thread = twice the warps ran 1.6x SLOWER),
// so the per-source work is kept minimal: descriptors are stored as (left,right) + two weights, and a source that does
// not fire -- the common case, ~95 % of the frames -- costs two shared loads, one compare and the four FMAs.
__global__ void __launch_bounds__(128)
cif_fire_fwd_kernel(const float* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ csum,
                    const int64_t* __restrict__ feat_len, int64_t B, int S, int C, float thr, int T,
                    float* __restrict__ out /* (B, T+1, C) */, uint8_t* __restrict__ fire_mask /* (B,S) nullable */,
                    float* __restrict__ tail_w /* (B,) nullable */) {
  extern __shared__ unsigned char s_raw[];
  int2* s_lr = reinterpret_cast<int2*>(s_raw);                 // (left, right) per source
  float2* s_w = reinterpret_cast<float2*>(s_lr + S);           // (left_w, right_w) per source
  const int64_t b = blockIdx.x;
  const float* cs = csum + b * S;
  const float* al = alpha + b * S;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const CifSrc d = cif_source(cs, al, s, thr, T);
    s_lr[s] = make_int2(d.left, d.right);
    s_w[s] = make_float2(d.left_w, d.right_w);
    if (fire_mask && blockIdx.y == 0) fire_mask[b * S + s] = d.right > d.left ? 1 : 0;
  }
  __syncthreads();
  if (tail_w && blockIdx.y == 0 && threadIdx.x == 0) {
    // contribution of every source to the row just past the fired length (cif.py:251-258), fixed order
    const int fl = (int)feat_len[b];
    float tw = 0.f;
    for (int s = 0; s < S; ++s) {
      if (s_lr[s].y == fl) tw += s_w[s].y;
      if (s_lr[s].x == fl) tw += s_w[s].x;
    }
    tail_w[b] = tw;
  }
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c0 >= C) return;
  const float* xb = x + b * (int64_t)S * C + c0;
  float* ob = out + b * (int64_t)(T + 1) * C + c0;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int t_cur = 0;
  auto move_to = [&](int t) {  // targets are visited in non-decreasing order
    *reinterpret_cast<float4*>(ob + (int64_t)t_cur * C) = acc;
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = t_cur + 1; z < t; ++z) *reinterpret_cast<float4*>(ob + (int64_t)z * C) = acc;
    t_cur = t;
  };
  auto add = [&](float w, const float4& v) {
    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
  };
  constexpr int kAhead = 8;  // source rows in flight per thread: the integration is sequential, the loads are not
  float4 buf[kAhead];
#pragma unroll
  for (int i = 0; i < kAhead; ++i)
    buf[i] = i < S ? *reinterpret_cast<const float4*>(xb + (int64_t)i * C) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s0 = 0; s0 < S; s0 += kAhead) {
#pragma unroll
    for (int i = 0; i < kAhead; ++i) {
      const int s = s0 + i;
      if (s < S) {
        const float4 v = buf[i];
        if (s + kAhead < S) buf[i] = *reinterpret_cast<const float4*>(xb + (int64_t)(s + kAhead) * C);
        const int2 lr = s_lr[s];
        const float2 w = s_w[s];
        if (lr.x != t_cur) move_to(lr.x);
        add(w.x, v);
        if (lr.y != lr.x) {  // the source fires: thr on every row strictly between, the remainder on the last
          for (int t = lr.x + 1; t < lr.y; ++t) {  // extra = right - left - 1 rows (clipped indices never exceed T)
            move_to(t);
            add(thr, v);
          }
          move_to(lr.y);
          add(w.y, v);
        }
      }
    }
  }
  move_to(T + 1);  // flushes the last row and zero-fills up to the tail row
}

// inference tail handling (cif.py:246-296): thread <-> 4 channels of one output row
//   extend[b] = tail_w[b] >= firing_thr ; row feat_len[b] is scaled by thr / tail_w[b] when extended ;
//   feat_len_new = min(feat_len + extend, max_len) ; rows >= feat_len_new are erased
__global__ void cif_tail_kernel(float* __restrict__ out, int64_t B, int T1, int C, const int64_t* __restrict__ feat_len,
                                const float* __restrict__ tail_w, float thr, float firing_thr, int max_len,
                                int64_t* __restrict__ feat_len_new) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int vec = C >> 2;
  if (i >= B * T1 * vec) return;
  const int64_t row = i / vec;
  const int64_t b = row / T1;
  const int t = (int)(row - b * T1);
  const int fl = (int)feat_len[b];
  const float tw = tail_w[b];
  const bool extend = tw >= firing_thr;
  int fn = fl + (extend ? 1 : 0);
  if (fn > max_len) fn = max_len;
  if (t == 0 && i % vec == 0) feat_len_new[b] = fn;
  float4* p = reinterpret_cast<float4*>(out + row * C) + (i % vec);
  if (t >= fn) {
    *p = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (extend && t == fl) {
    const float sc = thr / tw;
    float4 v = *p;
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    *p = v;
  }
}

// gradient row of the (B,T+1,C) integration buffer seen through slicing / tail handling:
//   G[b,t,:] = g_out[b,t,:] * (t == scale_row[b] ? scale[b] : 1) for t < min(T_out, keep_len[b]), else 0
struct CifGradView {
  const float* g_out;
  const int64_t* keep_len;   // nullable: rows >= keep_len[b] were erased in the forward pass
  const int64_t* scale_row;  // nullable: row that was up-scaled
  const float* tail_w;       // with scale_row: scale = thr / tail_w[b] when tail_w[b] >= firing_thr
  float thr, firing_thr;
  int T_out;
};
__device__ __forceinline__ float cif_row_factor(const CifGradView& gv, int64_t b, int t) {
  if (t >= gv.T_out) return 0.f;
  if (gv.keep_len && t >= (int)gv.keep_len[b]) return 0.f;
  if (gv.scale_row && t == (int)gv.scale_row[b] && gv.tail_w[b] >= gv.firing_thr) return gv.thr / gv.tail_w[b];
  return 1.f;
}

// warp <-> source (8 sources per block): d_input[b,s,:] and the channel dots Lg = <G[left], x>, Rg = fire * <G[right], x>
__global__ void __launch_bounds__(256)
cif_fire_bwd_kernel(const float* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ csum,
                    int64_t B, int S, int C, float thr, int T, CifGradView gv, float* __restrict__ g_x,
                    float* __restrict__ d_alpha_direct /* (B,S): Lg */, float* __restrict__ d_csum /* (B,S): fire*(Rg-Lg) */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  const int s = blockIdx.y * 8 + warp;
  if (s >= S) return;
  const CifSrc d = cif_source(csum + b * S, alpha + b * S, s, thr, T);
  const float* xr = x + (b * S + s) * (int64_t)C;
  const float* gb = gv.g_out + b * (int64_t)gv.T_out * C;
  const float fl = cif_row_factor(gv, b, d.left), fr = d.right > d.left ? cif_row_factor(gv, b, d.right) : 0.f;
  float lg = 0.f, rg = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(xr + c);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (fl != 0.f) {
      const float4 g = *reinterpret_cast<const float4*>(gb + (int64_t)d.left * C + c);
      const float w = d.left_w * fl;
      o.x = w * g.x; o.y = w * g.y; o.z = w * g.z; o.w = w * g.w;
      lg += fl * (g.x * xv.x + g.y * xv.y + g.z * xv.z + g.w * xv.w);
    }
    for (int e = 1; e <= d.extra; ++e) {
      const int t = min(d.left + e, T);
      const float f = cif_row_factor(gv, b, t);
      if (f != 0.f) {
        const float4 g = *reinterpret_cast<const float4*>(gb + (int64_t)t * C + c);
        const float w = thr * f;
        o.x = fmaf(w, g.x, o.x); o.y = fmaf(w, g.y, o.y); o.z = fmaf(w, g.z, o.z); o.w = fmaf(w, g.w, o.w);
      }
    }
    if (fr != 0.f) {
      const float4 g = *reinterpret_cast<const float4*>(gb + (int64_t)d.right * C + c);
      const float w = d.right_w * fr;
      o.x = fmaf(w, g.x, o.x); o.y = fmaf(w, g.y, o.y); o.z = fmaf(w, g.z, o.z); o.w = fmaf(w, g.w, o.w);
      rg += fr * (g.x * xv.x + g.y * xv.y + g.z * xv.z + g.w * xv.w);
    }
    *reinterpret_cast<float4*>(g_x + (b * S + s) * (int64_t)C + c) = o;
  }
  lg = warp_sum(lg);
  rg = warp_sum(rg);
  if (lane == 0) {
    d_alpha_direct[b * S + s] = lg;
    d_csum[b * S + s] = d.right > d.left ? rg - lg : 0.f;
  }
}

// g_alpha[b,j] = Lg[b,j] + sum_{s >= j} d_csum[b,s]   (reverse cumulative sum, sequential: deterministic)
__global__ void __launch_bounds__(128)
cif_alpha_grad_kernel(const float* __restrict__ d_alpha_direct, const float* __restrict__ d_csum, int64_t B, int S,
                      float* __restrict__ g_alpha) {
  extern __shared__ float s_row[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * 4 + warp;
  if (b >= B) return;
  float* row = s_row + (size_t)warp * S;
  for (int s = lane; s < S; s += 32) row[s] = d_csum[b * S + s];
  __syncwarp();
  if (lane == 0) {
    float acc = 0.f;
    for (int s = S - 1; s >= 0; --s) {
      acc += row[s];
      row[s] = acc;
    }
  }
  __syncwarp();
  for (int s = lane; s < S; s += 32) g_alpha[b * S + s] = d_alpha_direct[b * S + s] + row[s];
}

static int check_cif(int64_t B, int64_t S, int64_t C, float thr) {
  SCP_CHECK_ARG(B > 0 && S > 0 && C > 0 && C % 4 == 0, "cif: bad shape (C must be a multiple of 4)");
  SCP_CHECK_ARG(S <= 2048, "cif: S=%lld > 2048 source frames", (long long)S);
  SCP_CHECK_ARG(thr > 0.f, "cif: threshold must be positive");
  SCP_CHECK_ARG(B <= 2147483647ll, "cif: B too large");
  return SCP_OK;
}

}  // namespace scp

using namespace scp;

extern "C" int scp_cif_plan(const float* alpha, int64_t B, int64_t S, float threshold, int max_len, float* csum,
                            int64_t* feat_len, scp_stream_t stream) {
  int rc = check_cif(B, S, 4, threshold);
  if (rc) return rc;
  SCP_CHECK_ARG(alpha && csum && feat_len && max_len >= 1, "cif_plan: bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cif_plan_kernel<<<(unsigned)ceil_div(B, 4), 128, (size_t)4 * S * sizeof(float), s>>>(alpha, B, (int)S, threshold, max_len,
                                                                                    csum, feat_len);
  SCP_CUDA_LAUNCH_CHECK("cif_plan");
  return SCP_OK;
}

extern "C" int scp_cif_fire_fwd(const float* x, const float* alpha, const float* csum, const int64_t* feat_len, int64_t B,
                                int64_t S, int64_t C, float threshold, int64_t T, float* out, uint8_t* fire_mask,
                                float* tail_w, scp_stream_t stream) {
  int rc = check_cif(B, S, C, threshold);
  if (rc) return rc;
  SCP_CHECK_ARG(x && alpha && csum && feat_len && out && T >= 1, "cif_fire_fwd: bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid((unsigned)B, (unsigned)ceil_div(C, 512));
  cif_fire_fwd_kernel<<<grid, 128, (size_t)S * (sizeof(int2) + sizeof(float2)), s>>>(x, alpha, csum, feat_len, B, (int)S, (int)C, threshold,
                                                                  (int)T, out, fire_mask, tail_w);
  SCP_CUDA_LAUNCH_CHECK("cif_fire_fwd");
  return SCP_OK;
}

extern "C" int scp_cif_tail(float* out, int64_t B, int64_t T1, int64_t C, const int64_t* feat_len, const float* tail_w,
                            float threshold, float firing_threshold, int max_len, int64_t* feat_len_new,
                            scp_stream_t stream) {
  SCP_CHECK_ARG(out && feat_len && tail_w && feat_len_new && B > 0 && T1 > 0 && C > 0 && C % 4 == 0, "cif_tail: bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cif_tail_kernel<<<(unsigned)ceil_div(B * T1 * (C / 4), 256), 256, 0, s>>>(out, B, (int)T1, (int)C, feat_len, tail_w,
                                                                          threshold, firing_threshold, max_len, feat_len_new);
  SCP_CUDA_LAUNCH_CHECK("cif_tail");
  return SCP_OK;
}

extern "C" int scp_cif_fire_bwd(const float* g_out, int64_t T_out, const float* x, const float* alpha, const float* csum,
                                int64_t B, int64_t S, int64_t C, float threshold, int64_t T, const int64_t* keep_len,
                                const int64_t* scale_row, const float* tail_w, float firing_threshold, float* g_x,
                                float* g_alpha, void* workspace, size_t workspace_bytes, scp_stream_t stream) {
  int rc = check_cif(B, S, C, threshold);
  if (rc) return rc;
  SCP_CHECK_ARG(g_out && x && alpha && csum && g_x && g_alpha && workspace && T_out >= 1 && T_out <= T + 1,
                "cif_fire_bwd: bad argument");
  SCP_CHECK_ARG(!scale_row || tail_w, "cif_fire_bwd: scale_row needs tail_w");
  const size_t need = (size_t)2 * B * S * sizeof(float);
  if (workspace_bytes < need) return fail(SCP_ERR_WORKSPACE, "cif_fire_bwd: workspace %zu < %zu", workspace_bytes, need);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* lg = reinterpret_cast<float*>(workspace);
  float* dcs = lg + B * S;
  CifGradView gv{g_out, keep_len, scale_row, tail_w, threshold, firing_threshold, (int)T_out};
  const dim3 grid((unsigned)B, (unsigned)ceil_div(S, 8));
  cif_fire_bwd_kernel<<<grid, 256, 0, s>>>(x, alpha, csum, B, (int)S, (int)C, threshold, (int)T, gv, g_x, lg, dcs);
  SCP_CUDA_LAUNCH_CHECK("cif_fire_bwd");
  cif_alpha_grad_kernel<<<(unsigned)ceil_div(B, 4), 128, (size_t)4 * S * sizeof(float), s>>>(lg, dcs, B, (int)S, g_alpha);
  SCP_CUDA_LAUNCH_CHECK("cif_alpha_grad");
  return SCP_OK;
}
