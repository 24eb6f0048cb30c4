// S1 -- upstream-feature fusion: y = sum_l softmax(weights)_l * LN?(x_l)      (HBM-roofline kernels)
// Replaces WeightedSumLayer.forward of the reference (avssl/module/weighted_sum.py:26-45) and its autograd.
//
// Design (see DESIGN.md "S1"): the L layer tensors are never stacked or copied -- the kernel receives the L base
// pointers by value and addresses each as x_l[b*stride_b + t*stride_t + d], which is exactly the (T,B,D)-storage /
// (B,T,D)-view the HuBERT wrapper hands over (speech_encoder_plus.py:596-599).  Every byte is read once with 16-byte
// L1-bypassing loads; algorithmic traffic = L*B*T*D*s_in + B*T*D*s_out.
//   * plain variant: one 16 B vector per thread, all L loads in flight before the first FMA.
//   * LayerNorm variant: one warp per (b,t) row held in registers, mean / variance by warp shuffles (two-pass, as
//     torch's layer_norm), next layer's row prefetched while the current one is reduced.
//   * backward: d_l = <g_y, xhat_l> accumulated per lane/warp, per-block partials, deterministic finalize kernel that also
//     applies the softmax Jacobian  d_weights = w * (d - <w,d>).
//   * S1' (caller tail of FairseqSpeechEncoder_Hubert.forward, speech_encoder_plus.py:572-592) -- the per-layer rescale
//     the reference applies in a Python loop before the sum is fused as two more normalisation modes:
//       SCP_NORM_L2_FRAME ("method1"): x / (||x||_2 + 1e-8) per frame      -> the warp<->row kernel with an L2 statistic
//       SCP_NORM_UTT_MEAN ("method2"): x / mean_t ||x_t||_2 per utterance  -> a statistics pre-pass (scp_wsum_utt_scale)
//                                      writes 1/mean per (layer, utterance); the plain kernels fold it into the weight.
#include "scp_common.cuh"

namespace scp {

struct LayerPtrs {
  const void* p[SCP_MAX_LAYERS];
};
struct LayerOutPtrs {
  float* p[SCP_MAX_LAYERS];
};

constexpr int kWsumBwdBlocks = kNumSMs * 4;
constexpr int kWsumThreads = 256;

__device__ __forceinline__ void softmax_weights_to_smem(const float* __restrict__ weights, int L, float* sw) {
  // L <= 32: one warp computes softmax(weights) (weighted_sum.py:38)
  if (threadIdx.x < 32) {
    float v = (int)threadIdx.x < L ? weights[threadIdx.x] : -INFINITY;
    float m = warp_max(v);
    float e = (int)threadIdx.x < L ? expf(v - m) : 0.f;
    float s = warp_sum(e);
    if ((int)threadIdx.x < L) sw[threadIdx.x] = e / s;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------
// stores 4 consecutive elements (16 B for fp32 out, 8 B for 16-bit out)
template <typename TOut>
__device__ __forceinline__ void store4(TOut* dst, const float* f);
template <>
__device__ __forceinline__ void store4<float>(float* dst, const float* f) {
  st_stream16(dst, Vec16<float>::pack(f));
}
template <>
__device__ __forceinline__ void store4<__half>(__half* dst, const float* f) {
  __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
  uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  *reinterpret_cast<uint2*>(dst) = u;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* dst, const float* f) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  *reinterpret_cast<uint2*>(dst) = u;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kWsumThreads)
wsum_fwd_plain4_kernel(LayerPtrs lp, int L, int64_t n_vec, int vec_per_row, int64_t T, int64_t stride_b,
                       int64_t stride_t, const float* __restrict__ weights, const float* __restrict__ utt_scale,
                       int64_t B, TOut* __restrict__ y) {
  // plain forward: thread <-> one 16 B input vector of one row, all L loads issued before the first FMA;
  // output written in groups of 4 elements (any in/out dtype pair)
  constexpr int NE = Vec16<TIn>::NE;
  __shared__ float sw[SCP_MAX_LAYERS];
  softmax_weights_to_smem(weights, L, sw);
  const int64_t i = (int64_t)blockIdx.x * kWsumThreads + threadIdx.x;
  if (i >= n_vec) return;
  const int64_t r = i / vec_per_row;
  const int v = (int)(i - r * vec_per_row);
  const int64_t b = r / T, t = r - b * T;
  const int64_t off = b * stride_b + t * stride_t + (int64_t)v * NE;
  float acc[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) acc[e] = 0.f;
  for (int l0 = 0; l0 < L; l0 += 8) {
    uint4 raw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (l0 + j < L) raw[j] = ld_stream16(reinterpret_cast<const TIn*>(lp.p[l0 + j]) + off);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (l0 + j < L) {
        float f[NE];
        Vec16<TIn>::unpack(raw[j], f);
        float w = sw[l0 + j];
        if (utt_scale) w *= __ldg(utt_scale + (int64_t)(l0 + j) * B + b);  // method2: 1 / mean_t ||x_{l,b,t}||
#pragma unroll
        for (int e = 0; e < NE; ++e) acc[e] = fmaf(w, f[e], acc[e]);
      }
  }
  TOut* dst = y + r * (int64_t)vec_per_row * NE + (int64_t)v * NE;
#pragma unroll
  for (int c = 0; c < NE / 4; ++c) store4<TOut>(dst + 4 * c, acc + 4 * c);
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm forward: warp <-> row, NV vectors per lane kept in registers
template <typename TIn, int NV>
__device__ __forceinline__ void load_row(const TIn* base, int lane, int vec_per_row, uint4 (&raw)[NV]) {
  constexpr int NE = Vec16<TIn>::NE;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = lane + 32 * j;
    raw[j] = v < vec_per_row ? ld_stream16(base + (int64_t)v * NE) : make_uint4(0, 0, 0, 0);
  }
}

// returns mean and rstd of the row held as x[NV][NE] (invalid vectors are zero and excluded from the variance)
template <int NV, int NE>
__device__ __forceinline__ void row_stats(const float (&x)[NV][NE], int lane, int vec_per_row, int D, float eps,
                                          float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < NE; ++e) s += x[j][e];
  mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (lane + 32 * j < vec_per_row) {
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const float c = x[j][e] - mean;
        q = fmaf(c, c, q);
      }
    }
  const float var = warp_sum(q) / (float)D;
  rstd = 1.0f / sqrtf(var + eps);
}

// L2 norm of the row held as x[NV][NE] (invalid vectors are zero)
template <int NV, int NE>
__device__ __forceinline__ float row_l2(const float (&x)[NV][NE]) {
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < NE; ++e) q = fmaf(x[j][e], x[j][e], q);
  return sqrtf(warp_sum(q));
}

constexpr float kL2FrameEps = 1e-8f;  // speech_encoder_plus.py:582

// MODE = SCP_NORM_LAYERNORM: xhat = (x - mean) * rstd ; MODE = SCP_NORM_L2_FRAME: xhat = x / (||x|| + 1e-8)
template <typename TIn, typename TOut, int NV, int MODE>
__global__ void __launch_bounds__(kWsumThreads)
wsum_fwd_ln_kernel(LayerPtrs lp, int L, int64_t n_rows, int vec_per_row, int D, int64_t T, int64_t stride_b,
                   int64_t stride_t, const float* __restrict__ weights, float eps, TOut* __restrict__ y) {
  constexpr int NE = Vec16<TIn>::NE;
  __shared__ float sw[SCP_MAX_LAYERS];
  softmax_weights_to_smem(weights, L, sw);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (kWsumThreads / 32) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (kWsumThreads / 32);
  for (int64_t r = warp0; r < n_rows; r += n_warps) {
    const int64_t b = r / T, t = r - b * T;
    const int64_t off = b * stride_b + t * stride_t;
    float acc[NV][NE];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e) acc[j][e] = 0.f;
    uint4 cur[NV], nxt[NV];
    load_row<TIn, NV>(reinterpret_cast<const TIn*>(lp.p[0]) + off, lane, vec_per_row, cur);
    for (int l = 0; l < L; ++l) {
      if (l + 1 < L) load_row<TIn, NV>(reinterpret_cast<const TIn*>(lp.p[l + 1]) + off, lane, vec_per_row, nxt);
      float x[NV][NE];
#pragma unroll
      for (int j = 0; j < NV; ++j) Vec16<TIn>::unpack(cur[j], x[j]);
      float mean = 0.f, rstd;
      if (MODE == SCP_NORM_LAYERNORM) row_stats<NV, NE>(x, lane, vec_per_row, D, eps, mean, rstd);
      else rstd = 1.0f / (row_l2<NV, NE>(x) + kL2FrameEps);
      const float a = sw[l] * rstd;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int e = 0; e < NE; ++e) acc[j][e] = fmaf(a, x[j][e] - mean, acc[j][e]);
#pragma unroll
      for (int j = 0; j < NV; ++j) cur[j] = nxt[j];
    }
    TOut* dst = y + r * (int64_t)D;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int v = lane + 32 * j;
      if (v < vec_per_row) {
#pragma unroll
        for (int c = 0; c < NE / 4; ++c) store4<TOut>(dst + (int64_t)v * NE + 4 * c, acc[j] + 4 * c);
      }
    }
  }
}


// loads NE consecutive gradient elements (NE = 4 or 8) as fp32 lanes with 16-byte accesses
template <typename TG, int NE>
__device__ __forceinline__ void load_grad_vec(const TG* p, float* g) {
  constexpr int NEG = Vec16<TG>::NE;
  static_assert(NE % NEG == 0, "gradient dtype wider than the layer dtype is not supported");
#pragma unroll
  for (int c = 0; c < NE / NEG; ++c) Vec16<TG>::unpack(ld_stream16(p + c * NEG), g + c * NEG);
}

// ------------------------------------------------------------------------------------------------------------
// backward, plain: d_l = sum g*x_l ; optional g_l = w_l * g
// LCAP = compile-time bound on the number of layers (16, 25 or 32): the per-thread accumulators are sized by it; G = layers
// loaded per group (G 16-byte loads in flight per thread).  LCAP + 4 G registers must leave the kernel at <= 85 registers:
// three resident blocks per SM instead of two -- half again as many loads in flight, which is what an HBM-bound read-only
// kernel lives on (16 / 8 for HuBERT-base, 25 / 5 for HuBERT-large, 32 / 4 up to SCP_MAX_LAYERS).
template <typename TIn, typename TG, int LCAP, int G = 8, int MINB = 3>
__global__ void __launch_bounds__(kWsumThreads, MINB)
wsum_bwd_plain_kernel(LayerPtrs lp, int L, int64_t n_vec, int vec_per_row, int64_t T, int64_t stride_b,
                      int64_t stride_t, const float* __restrict__ weights, const float* __restrict__ utt_scale,
                      int64_t B, const TG* __restrict__ g_y, float* __restrict__ partials, LayerOutPtrs gl,
                      int write_gl) {
  constexpr int NE = Vec16<TIn>::NE;  // elements handled per thread-iteration
  __shared__ float sw[SCP_MAX_LAYERS];
  __shared__ float sred[kWsumThreads / 32][SCP_MAX_LAYERS];
  softmax_weights_to_smem(weights, L, sw);
  float acc[LCAP];
#pragma unroll
  for (int l = 0; l < LCAP; ++l) acc[l] = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kWsumThreads + threadIdx.x; i < n_vec;
       i += (int64_t)gridDim.x * kWsumThreads) {
    const int64_t r = i / vec_per_row;
    const int v = (int)(i - r * vec_per_row);
    const int64_t b = r / T, t = r - b * T;
    const int64_t off = b * stride_b + t * stride_t + (int64_t)v * NE;
    const int64_t goff = r * (int64_t)vec_per_row * NE + (int64_t)v * NE;
    float g[NE];
    load_grad_vec<TG, NE>(g_y + goff, g);
#pragma unroll
    for (int l0 = 0; l0 < LCAP; l0 += G) {
      if (l0 < L) {
        uint4 raw[G];
#pragma unroll
        for (int j = 0; j < G; ++j)
          if (l0 + j < LCAP && l0 + j < L) raw[j] = ld_stream16(reinterpret_cast<const TIn*>(lp.p[l0 + j]) + off);
#pragma unroll
        for (int j = 0; j < G; ++j)
          if (l0 + j < LCAP && l0 + j < L) {
            float f[NE];
            Vec16<TIn>::unpack(raw[j], f);
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < NE; ++e) s = fmaf(g[e], f[e], s);
            if (utt_scale) s *= __ldg(utt_scale + (int64_t)(l0 + j) * B + b);  // d_l = <g, x_l / mean-norm>
            acc[l0 + j] += s;
            if (write_gl) {
              const float w = sw[l0 + j];
              float o[NE];
#pragma unroll
              for (int e = 0; e < NE; ++e) o[e] = w * g[e];
#pragma unroll
              for (int c = 0; c < NE / 4; ++c) store4<float>(gl.p[l0 + j] + goff + 4 * c, o + 4 * c);
            }
          }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int l = 0; l < LCAP; ++l) {
    if (l < L) {
      const float s = warp_sum(acc[l]);
      if (lane == 0) sred[warp][l] = s;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < L) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWsumThreads / 32; ++w) s += sred[w][threadIdx.x];
    partials[(int64_t)blockIdx.x * SCP_MAX_LAYERS + threadIdx.x] = s;
  }
}

// backward, LayerNorm / per-frame L2: warp <-> row
template <typename TIn, typename TG, int NV, int MODE>
__global__ void __launch_bounds__(kWsumThreads)
wsum_bwd_ln_kernel(LayerPtrs lp, int L, int64_t n_rows, int vec_per_row, int D, int64_t T, int64_t stride_b,
                   int64_t stride_t, const float* __restrict__ weights, float eps, const TG* __restrict__ g_y,
                   float* __restrict__ partials, LayerOutPtrs gl, int write_gl) {
  constexpr int NE = Vec16<TIn>::NE;
  __shared__ float sw[SCP_MAX_LAYERS];
  __shared__ float sred[kWsumThreads / 32][SCP_MAX_LAYERS];
  softmax_weights_to_smem(weights, L, sw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < SCP_MAX_LAYERS) sred[warp][lane] = 0.f;
  __syncwarp();
  const int64_t warp0 = (int64_t)blockIdx.x * (kWsumThreads / 32) + warp;
  const int64_t n_warps = (int64_t)gridDim.x * (kWsumThreads / 32);
  for (int64_t r = warp0; r < n_rows; r += n_warps) {
    const int64_t b = r / T, t = r - b * T;
    const int64_t off = b * stride_b + t * stride_t;
    // gradient row (fp32 lanes), laid out like the layer vectors
    float g[NV][NE];
    float gsum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int v = lane + 32 * j;
#pragma unroll
      for (int e = 0; e < NE; ++e) g[j][e] = 0.f;
      if (v < vec_per_row) load_grad_vec<TG, NE>(g_y + r * (int64_t)D + (int64_t)v * NE, g[j]);
#pragma unroll
      for (int e = 0; e < NE; ++e) gsum += g[j][e];
    }
    gsum = warp_sum(gsum);
    uint4 cur[NV], nxt[NV];
    load_row<TIn, NV>(reinterpret_cast<const TIn*>(lp.p[0]) + off, lane, vec_per_row, cur);
    for (int l = 0; l < L; ++l) {
      if (l + 1 < L) load_row<TIn, NV>(reinterpret_cast<const TIn*>(lp.p[l + 1]) + off, lane, vec_per_row, nxt);
      float x[NV][NE];
#pragma unroll
      for (int j = 0; j < NV; ++j) Vec16<TIn>::unpack(cur[j], x[j]);
      float mean = 0.f, rstd, l2 = 0.f;
      if (MODE == SCP_NORM_LAYERNORM) {
        row_stats<NV, NE>(x, lane, vec_per_row, D, eps, mean, rstd);
      } else {
        l2 = row_l2<NV, NE>(x);
        rstd = 1.0f / (l2 + kL2FrameEps);
      }
      float dot = 0.f;  // sum_d g * xhat
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (lane + 32 * j < vec_per_row) {
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            x[j][e] = (x[j][e] - mean) * rstd;
            dot = fmaf(g[j][e], x[j][e], dot);
          }
        }
      dot = warp_sum(dot);
      if (lane == 0) sred[warp][l] += dot;
      if (write_gl) {
        // LayerNorm backward with upstream gradient w_l*g:  dx = rstd*w_l*(g - mean(g) - xhat*mean(g*xhat))
        // per-frame L2 (y = x/(n+eps)):                       dx = w_l/(n+eps)*(g - xhat*<g,xhat>*(n+eps)/n), n = 0 -> no 2nd term
        const float a = rstd * sw[l];
        const float mg = MODE == SCP_NORM_LAYERNORM ? gsum / (float)D : 0.f;
        const float mgx = MODE == SCP_NORM_LAYERNORM ? dot / (float)D
                                                      : (l2 > 0.f ? dot * (l2 + kL2FrameEps) / l2 : 0.f);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int v = lane + 32 * j;
          if (v < vec_per_row) {
            float o[NE];
#pragma unroll
            for (int e = 0; e < NE; ++e) o[e] = a * (g[j][e] - mg - x[j][e] * mgx);
#pragma unroll
            for (int c = 0; c < NE / 4; ++c)
              store4<float>(gl.p[l] + r * (int64_t)D + (int64_t)v * NE + 4 * c, o + 4 * c);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) cur[j] = nxt[j];
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < L) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWsumThreads / 32; ++w) s += sred[w][threadIdx.x];
    partials[(int64_t)blockIdx.x * SCP_MAX_LAYERS + threadIdx.x] = s;
  }
}

// two sums reduced together: the shuffles of the two butterflies interleave instead of forming two dependent chains
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}

// backward, weights only (the frozen-HuBERT case of every shipped recipe): d_l = sum_rows <g, xhat_l>.
// <g, xhat> = rstd * <g, x - mean>, so a row needs two reduction rounds for LayerNorm (mean; then variance and the
// dot together) and one for the per-frame L2 norm -- and none of the state of the dense-gradient path: 151 -> ~90
// registers, two resident blocks per SM instead of one (the general kernel measured 1.0 ms = 0.43 of the HBM peak).
template <typename TIn, typename TG, int NV, int MODE>
__global__ void __launch_bounds__(kWsumThreads, 2)
wsum_bwd_ln_w_kernel(LayerPtrs lp, int L, int64_t n_rows, int vec_per_row, int D, int64_t T, int64_t stride_b,
                     int64_t stride_t, float eps, const TG* __restrict__ g_y, float* __restrict__ partials) {
  constexpr int NE = Vec16<TIn>::NE;
  __shared__ float sred[kWsumThreads / 32][SCP_MAX_LAYERS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < SCP_MAX_LAYERS) sred[warp][lane] = 0.f;
  __syncwarp();
  const float inv_d = 1.0f / (float)D;
  const int64_t warp0 = (int64_t)blockIdx.x * (kWsumThreads / 32) + warp;
  const int64_t n_warps = (int64_t)gridDim.x * (kWsumThreads / 32);
  for (int64_t r = warp0; r < n_rows; r += n_warps) {
    const int64_t b = r / T, t = r - b * T;
    const int64_t off = b * stride_b + t * stride_t;
    float g[NV][NE];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int v = lane + 32 * j;
#pragma unroll
      for (int e = 0; e < NE; ++e) g[j][e] = 0.f;
      if (v < vec_per_row) load_grad_vec<TG, NE>(g_y + r * (int64_t)D + (int64_t)v * NE, g[j]);
    }
    uint4 cur[NV], nxt[NV];
    load_row<TIn, NV>(reinterpret_cast<const TIn*>(lp.p[0]) + off, lane, vec_per_row, cur);
    for (int l = 0; l < L; ++l) {
      if (l + 1 < L) load_row<TIn, NV>(reinterpret_cast<const TIn*>(lp.p[l + 1]) + off, lane, vec_per_row, nxt);
      float x[NV][NE];
#pragma unroll
      for (int j = 0; j < NV; ++j) Vec16<TIn>::unpack(cur[j], x[j]);
      float mean = 0.f;
      if (MODE == SCP_NORM_LAYERNORM) {
        float sx = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
          for (int e = 0; e < NE; ++e) sx += x[j][e];
        mean = warp_sum(sx) * inv_d;
      }
      float q = 0.f, gd = 0.f;  // sum (x-mean)^2 over the valid vectors, sum g*(x-mean)
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (lane + 32 * j < vec_per_row) {
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            const float c = x[j][e] - mean;
            q = fmaf(c, c, q);
            gd = fmaf(g[j][e], c, gd);
          }
        }
      warp_sum2(q, gd);
      const float rstd = MODE == SCP_NORM_LAYERNORM ? 1.0f / sqrtf(q * inv_d + eps) : 1.0f / (sqrtf(q) + kL2FrameEps);
      if (lane == 0) sred[warp][l] += rstd * gd;
#pragma unroll
      for (int j = 0; j < NV; ++j) cur[j] = nxt[j];
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < L) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWsumThreads / 32; ++w) s += sred[w][threadIdx.x];
    partials[(int64_t)blockIdx.x * SCP_MAX_LAYERS + threadIdx.x] = s;
  }
}

// method2 statistics: utt_scale[l*B + b] = 1 / mean_t ||x_l[b,t,:]||_2   (speech_encoder_plus.py:584-590)
// grid (B, L); warp <-> frame (strided over t), fixed-order block reduction -> deterministic
template <typename TIn>
__global__ void __launch_bounds__(kWsumThreads)
wsum_utt_scale_kernel(LayerPtrs lp, int64_t B, int64_t T, int vec_per_row, int64_t stride_b, int64_t stride_t,
                      float* __restrict__ utt_scale) {
  constexpr int NE = Vec16<TIn>::NE;
  __shared__ float sred[kWsumThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  const int l = blockIdx.y;
  const TIn* base = reinterpret_cast<const TIn*>(lp.p[l]) + b * stride_b;
  float tot = 0.f;
  for (int64_t t = warp; t < T; t += kWsumThreads / 32) {
    const TIn* row = base + t * stride_t;
    float q = 0.f;
    for (int v = lane; v < vec_per_row; v += 32) {
      float f[NE];
      Vec16<TIn>::unpack(ld_stream16(row + (int64_t)v * NE), f);
#pragma unroll
      for (int e = 0; e < NE; ++e) q = fmaf(f[e], f[e], q);
    }
    tot += sqrtf(warp_sum(q));
  }
  if (lane == 0) sred[warp] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWsumThreads / 32; ++w) s += sred[w];
    utt_scale[(int64_t)l * B + b] = (float)T / s;  // no epsilon in the reference: an all-zero utterance gives inf
  }
}

// sums the per-block partials in a fixed order and applies the softmax Jacobian
__global__ void wsum_bwd_finalize_kernel(const float* __restrict__ partials, int n_blocks, int L,
                                         const float* __restrict__ weights, float* __restrict__ d_weights) {
  __shared__ float sw[SCP_MAX_LAYERS];
  __shared__ float sd[SCP_MAX_LAYERS];
  softmax_weights_to_smem(weights, L, sw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 32 warps: warp <-> layer
  if (warp < L) {
    float s = 0.f;
    for (int b = lane; b < n_blocks; b += 32) s += partials[(int64_t)b * SCP_MAX_LAYERS + warp];
    s = warp_sum(s);
    if (lane == 0) sd[warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float wd = (int)threadIdx.x < L ? sw[threadIdx.x] * sd[threadIdx.x] : 0.f;
    const float tot = warp_sum(wd);
    if ((int)threadIdx.x < L) d_weights[threadIdx.x] = sw[threadIdx.x] * (sd[threadIdx.x] - tot);
  }
}

// ------------------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
static int launch_fwd(const LayerPtrs& lp, int L, int64_t B, int64_t T, int64_t D, int64_t sb, int64_t st,
                      const float* weights, int norm_mode, float eps, const float* utt_scale, void* y,
                      cudaStream_t stream) {
  constexpr int NE = Vec16<TIn>::NE;
  const int vec_per_row = (int)(D / NE);
  const int64_t n_rows = B * T;
  if (norm_mode == SCP_NORM_NONE || norm_mode == SCP_NORM_UTT_MEAN) {
    const int64_t n_vec = n_rows * vec_per_row;
    const int64_t blocks = ceil_div(n_vec, kWsumThreads);
    wsum_fwd_plain4_kernel<TIn, TOut><<<(unsigned)blocks, kWsumThreads, 0, stream>>>(
        lp, L, n_vec, vec_per_row, T, sb, st, weights, norm_mode == SCP_NORM_UTT_MEAN ? utt_scale : nullptr, B,
        reinterpret_cast<TOut*>(y));
    SCP_CUDA_LAUNCH_CHECK("wsum_fwd_plain");
    return SCP_OK;
  }
  const int nv = (int)ceil_div(vec_per_row, 32);
  const int64_t blocks = std::min<int64_t>(ceil_div(n_rows, kWsumThreads / 32), (int64_t)kNumSMs * 8);
#define SCP_LN_CASE(NVV)                                                                                         \
  case NVV:                                                                                                      \
    if (norm_mode == SCP_NORM_LAYERNORM)                                                                         \
      wsum_fwd_ln_kernel<TIn, TOut, NVV, SCP_NORM_LAYERNORM><<<(unsigned)blocks, kWsumThreads, 0, stream>>>(     \
          lp, L, n_rows, vec_per_row, (int)D, T, sb, st, weights, eps, reinterpret_cast<TOut*>(y));              \
    else                                                                                                         \
      wsum_fwd_ln_kernel<TIn, TOut, NVV, SCP_NORM_L2_FRAME><<<(unsigned)blocks, kWsumThreads, 0, stream>>>(      \
          lp, L, n_rows, vec_per_row, (int)D, T, sb, st, weights, eps, reinterpret_cast<TOut*>(y));              \
    break;
  switch (nv) {
    SCP_LN_CASE(1) SCP_LN_CASE(2) SCP_LN_CASE(3) SCP_LN_CASE(4) SCP_LN_CASE(5) SCP_LN_CASE(6) SCP_LN_CASE(7)
    SCP_LN_CASE(8)
    default:
      return fail(SCP_ERR_UNSUPPORTED, "wsum per-frame normalisation supports D <= %d for this dtype (got %lld)",
                  8 * 32 * NE, (long long)D);
  }
#undef SCP_LN_CASE
  SCP_CUDA_LAUNCH_CHECK("wsum_fwd_ln");
  return SCP_OK;
}

template <typename TIn, typename TG>
static int launch_bwd(const LayerPtrs& lp, int L, int64_t B, int64_t T, int64_t D, int64_t sb, int64_t st,
                      const float* weights, int norm_mode, float eps, const float* utt_scale, const void* g_y,
                      float* d_weights, const LayerOutPtrs& gl, int write_gl, float* partials, cudaStream_t stream) {
  constexpr int NE = Vec16<TIn>::NE;
  const int vec_per_row = (int)(D / NE);
  const int64_t n_rows = B * T;
  int blocks;
  if (norm_mode == SCP_NORM_NONE || norm_mode == SCP_NORM_UTT_MEAN) {
    const int64_t n_vec = n_rows * vec_per_row;
    if (L <= 16) {
      blocks = (int)std::min<int64_t>(ceil_div(n_vec, kWsumThreads), kNumSMs * 3);  // one wave of three blocks per SM
      wsum_bwd_plain_kernel<TIn, TG, 16><<<blocks, kWsumThreads, 0, stream>>>(
          lp, L, n_vec, vec_per_row, T, sb, st, weights, norm_mode == SCP_NORM_UTT_MEAN ? utt_scale : nullptr, B,
          reinterpret_cast<const TG*>(g_y), partials, gl, write_gl);
    } else {
      // 17..32 layers: shorter load groups keep the kernel at 80 registers = three resident blocks per SM as well
      // (groups of eight need 113: two blocks).  HuBERT-large (25 layers, B*T = 512*249, D = 1024): 2.22 -> 2.01 ms.
      blocks = (int)std::min<int64_t>(ceil_div(n_vec, kWsumThreads), kNumSMs * 3);
#define SCP_WB(LC, GG)                                                                                            \
  wsum_bwd_plain_kernel<TIn, TG, LC, GG, 3><<<blocks, kWsumThreads, 0, stream>>>(                                  \
      lp, L, n_vec, vec_per_row, T, sb, st, weights, norm_mode == SCP_NORM_UTT_MEAN ? utt_scale : nullptr, B,     \
      reinterpret_cast<const TG*>(g_y), partials, gl, write_gl)
      if (L <= 25) SCP_WB(25, 5);
      else SCP_WB(32, 4);
#undef SCP_WB
    }
    SCP_CUDA_LAUNCH_CHECK("wsum_bwd_plain");
  } else {
    const int nv = (int)ceil_div(vec_per_row, 32);
    blocks = (int)std::min<int64_t>(ceil_div(n_rows, kWsumThreads / 32), kWsumBwdBlocks);
#define SCP_LN_CASE(NVV)                                                                                     \
  case NVV:                                                                                                  \
    if (!write_gl && norm_mode == SCP_NORM_LAYERNORM)                                                        \
      wsum_bwd_ln_w_kernel<TIn, TG, NVV, SCP_NORM_LAYERNORM><<<blocks, kWsumThreads, 0, stream>>>(           \
          lp, L, n_rows, vec_per_row, (int)D, T, sb, st, eps, reinterpret_cast<const TG*>(g_y), partials);   \
    else if (!write_gl)                                                                                      \
      wsum_bwd_ln_w_kernel<TIn, TG, NVV, SCP_NORM_L2_FRAME><<<blocks, kWsumThreads, 0, stream>>>(            \
          lp, L, n_rows, vec_per_row, (int)D, T, sb, st, eps, reinterpret_cast<const TG*>(g_y), partials);   \
    else if (norm_mode == SCP_NORM_LAYERNORM)                                                                \
      wsum_bwd_ln_kernel<TIn, TG, NVV, SCP_NORM_LAYERNORM><<<blocks, kWsumThreads, 0, stream>>>(             \
          lp, L, n_rows, vec_per_row, (int)D, T, sb, st, weights, eps, reinterpret_cast<const TG*>(g_y),     \
          partials, gl, write_gl);                                                                           \
    else                                                                                                     \
      wsum_bwd_ln_kernel<TIn, TG, NVV, SCP_NORM_L2_FRAME><<<blocks, kWsumThreads, 0, stream>>>(              \
          lp, L, n_rows, vec_per_row, (int)D, T, sb, st, weights, eps, reinterpret_cast<const TG*>(g_y),     \
          partials, gl, write_gl);                                                                           \
    break;
    switch (nv) {
      SCP_LN_CASE(1) SCP_LN_CASE(2) SCP_LN_CASE(3) SCP_LN_CASE(4) SCP_LN_CASE(5) SCP_LN_CASE(6) SCP_LN_CASE(7)
      SCP_LN_CASE(8)
      default:
        return fail(SCP_ERR_UNSUPPORTED, "wsum per-frame normalisation supports D <= %d for this dtype (got %lld)",
                    8 * 32 * NE, (long long)D);
    }
#undef SCP_LN_CASE
    SCP_CUDA_LAUNCH_CHECK("wsum_bwd_ln");
  }
  wsum_bwd_finalize_kernel<<<1, 1024, 0, stream>>>(partials, blocks, L, weights, d_weights);
  SCP_CUDA_LAUNCH_CHECK("wsum_bwd_finalize");
  return SCP_OK;
}

static int check_wsum_args(const void* const* layer_ptrs, int L, int64_t B, int64_t T, int64_t D, int64_t sb,
                           int64_t st, int dtype_in, const void* weights) {
  SCP_CHECK_ARG(layer_ptrs && weights, "wsum: null pointer");
  SCP_CHECK_ARG(L >= 1 && L <= SCP_MAX_LAYERS, "wsum: L=%d outside [1,%d]", L, SCP_MAX_LAYERS);
  SCP_CHECK_ARG(B > 0 && T > 0 && D > 0, "wsum: non-positive shape");
  SCP_CHECK_ARG(dtype_in >= SCP_F32 && dtype_in <= SCP_BF16, "wsum: bad dtype_in %d", dtype_in);
  const int ne = dtype_in == SCP_F32 ? 4 : 8;
  if (D % ne || sb % ne || st % ne)
    return fail(SCP_ERR_UNSUPPORTED, "wsum: D and strides must be multiples of %d elements (16 B vectors)", ne);
  for (int l = 0; l < L; ++l) {
    SCP_CHECK_ARG(layer_ptrs[l] != nullptr, "wsum: layer %d is null", l);
    if (reinterpret_cast<uintptr_t>(layer_ptrs[l]) & 15)
      return fail(SCP_ERR_UNSUPPORTED, "wsum: layer %d is not 16-byte aligned", l);
  }
  return SCP_OK;
}

}  // namespace scp

using namespace scp;

static int check_norm_mode(int norm_mode, const float* utt_scale) {
  SCP_CHECK_ARG(norm_mode >= SCP_NORM_NONE && norm_mode <= SCP_NORM_UTT_MEAN, "wsum: bad norm_mode %d", norm_mode);
  SCP_CHECK_ARG(norm_mode != SCP_NORM_UTT_MEAN || utt_scale, "wsum: SCP_NORM_UTT_MEAN needs utt_scale (scp_wsum_utt_scale)");
  return SCP_OK;
}

extern "C" int scp_wsum_utt_scale(const void* const* layer_ptrs, int L, int64_t B, int64_t T, int64_t D,
                                  int64_t stride_b, int64_t stride_t, int dtype_in, float* utt_scale,
                                  scp_stream_t stream) {
  int rc = check_wsum_args(layer_ptrs, L, B, T, D, stride_b, stride_t, dtype_in, utt_scale);
  if (rc) return rc;
  SCP_CHECK_ARG(B <= 2147483647ll, "wsum_utt_scale: B too large");
  LayerPtrs lp{};
  for (int l = 0; l < L; ++l) lp.p[l] = layer_ptrs[l];
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid((unsigned)B, (unsigned)L);
  const int ne = dtype_in == SCP_F32 ? 4 : 8;
  if (dtype_in == SCP_F32)
    wsum_utt_scale_kernel<float><<<grid, kWsumThreads, 0, s>>>(lp, B, T, (int)(D / ne), stride_b, stride_t, utt_scale);
  else if (dtype_in == SCP_F16)
    wsum_utt_scale_kernel<__half><<<grid, kWsumThreads, 0, s>>>(lp, B, T, (int)(D / ne), stride_b, stride_t, utt_scale);
  else
    wsum_utt_scale_kernel<__nv_bfloat16><<<grid, kWsumThreads, 0, s>>>(lp, B, T, (int)(D / ne), stride_b, stride_t, utt_scale);
  SCP_CUDA_LAUNCH_CHECK("wsum_utt_scale");
  return SCP_OK;
}

extern "C" int scp_wsum_fwd(const void* const* layer_ptrs, int L, int64_t B, int64_t T, int64_t D, int64_t stride_b,
                            int64_t stride_t, int dtype_in, const float* weights, int norm_mode, float eps,
                            const float* utt_scale, void* y, int dtype_out, scp_stream_t stream) {
  int rc = check_wsum_args(layer_ptrs, L, B, T, D, stride_b, stride_t, dtype_in, weights);
  if (rc) return rc;
  if ((rc = check_norm_mode(norm_mode, utt_scale))) return rc;
  SCP_CHECK_ARG(y && !(reinterpret_cast<uintptr_t>(y) & 15), "wsum_fwd: y null or misaligned");
  LayerPtrs lp{};
  for (int l = 0; l < L; ++l) lp.p[l] = layer_ptrs[l];
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define SCP_DISPATCH(TI, TO) \
  return launch_fwd<TI, TO>(lp, L, B, T, D, stride_b, stride_t, weights, norm_mode, eps, utt_scale, y, s)
  if (dtype_in == SCP_F32 && dtype_out == SCP_F32) SCP_DISPATCH(float, float);
  if (dtype_in == SCP_F16 && dtype_out == SCP_F32) SCP_DISPATCH(__half, float);
  if (dtype_in == SCP_BF16 && dtype_out == SCP_F32) SCP_DISPATCH(__nv_bfloat16, float);
  if (dtype_in == SCP_F16 && dtype_out == SCP_F16) SCP_DISPATCH(__half, __half);
  if (dtype_in == SCP_BF16 && dtype_out == SCP_BF16) SCP_DISPATCH(__nv_bfloat16, __nv_bfloat16);
#undef SCP_DISPATCH
  return fail(SCP_ERR_UNSUPPORTED, "wsum_fwd: dtype pair in=%d out=%d", dtype_in, dtype_out);
}

extern "C" size_t scp_wsum_bwd_workspace_bytes(int, int64_t, int64_t, int64_t) {
  return (size_t)kWsumBwdBlocks * SCP_MAX_LAYERS * sizeof(float);
}

extern "C" int scp_wsum_bwd(const void* const* layer_ptrs, int L, int64_t B, int64_t T, int64_t D, int64_t stride_b,
                            int64_t stride_t, int dtype_in, const float* weights, int norm_mode, float eps,
                            const float* utt_scale, const void* g_y, int dtype_g, float* d_weights,
                            void* const* g_layers, void* workspace, size_t workspace_bytes, scp_stream_t stream) {
  int rc = check_wsum_args(layer_ptrs, L, B, T, D, stride_b, stride_t, dtype_in, weights);
  if (rc) return rc;
  if ((rc = check_norm_mode(norm_mode, utt_scale))) return rc;
  if (norm_mode == SCP_NORM_UTT_MEAN && g_layers)
    return fail(SCP_ERR_UNSUPPORTED, "wsum_bwd: layer gradients through the per-utterance mean-norm (method2) are not "
                                     "implemented (no shipped recipe trains HuBERT with normalize_type=method2)");
  SCP_CHECK_ARG(g_y && d_weights && workspace, "wsum_bwd: null pointer");
  SCP_CHECK_ARG(!(reinterpret_cast<uintptr_t>(g_y) & 15), "wsum_bwd: g_y misaligned");
  if (workspace_bytes < scp_wsum_bwd_workspace_bytes(L, B, T, D))
    return fail(SCP_ERR_WORKSPACE, "wsum_bwd: workspace %zu < %zu", workspace_bytes,
                scp_wsum_bwd_workspace_bytes(L, B, T, D));
  LayerPtrs lp{};
  LayerOutPtrs gl{};
  for (int l = 0; l < L; ++l) lp.p[l] = layer_ptrs[l];
  if (g_layers)
    for (int l = 0; l < L; ++l) {
      SCP_CHECK_ARG(g_layers[l] && !(reinterpret_cast<uintptr_t>(g_layers[l]) & 15), "wsum_bwd: g_layers[%d]", l);
      gl.p[l] = reinterpret_cast<float*>(g_layers[l]);
    }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* partials = reinterpret_cast<float*>(workspace);
#define SCP_DISPATCH(TI, TG)                                                                                   \
  return launch_bwd<TI, TG>(lp, L, B, T, D, stride_b, stride_t, weights, norm_mode, eps, utt_scale, g_y,      \
                            d_weights, gl, g_layers != nullptr, partials, s)
  if (dtype_in == SCP_F32 && dtype_g == SCP_F32) SCP_DISPATCH(float, float);
  if (dtype_in == SCP_F16 && dtype_g == SCP_F32) SCP_DISPATCH(__half, float);
  if (dtype_in == SCP_BF16 && dtype_g == SCP_F32) SCP_DISPATCH(__nv_bfloat16, float);
  if (dtype_in == SCP_F16 && dtype_g == SCP_F16) SCP_DISPATCH(__half, __half);
  if (dtype_in == SCP_BF16 && dtype_g == SCP_BF16) SCP_DISPATCH(__nv_bfloat16, __nv_bfloat16);
#undef SCP_DISPATCH
  return fail(SCP_ERR_UNSUPPORTED, "wsum_bwd: dtype pair in=%d g=%d", dtype_in, dtype_g);
}
