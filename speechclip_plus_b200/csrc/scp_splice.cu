// N3 -- the step right after the VQ on the cascaded path: ClipModel.encode_keywords' prologue
// (avssl/module/clip_official.py:222-279) and get_keypadding_mask (avssl/util/data_utils.py:6-22).
//
// The reference builds the text-transformer input with a token-embedding lookup of a (B,77) id tensor, then a Python
// loop over the batch that slice-assigns every utterance's keywords (one tiny kernel per sample), then adds the
// positional embedding.  Here ONE pass writes  x[b,l,:] = src(b,l) + pos[l]  with
//     src = E[sot]            l == 0
//           keywords[b,l-1]   1 <= l <= n_b
//           E[eot]            l == n_b + 1
//           E[0]              otherwise (the id tensor is zero-initialised, clip_official.py:240)
// and the EOT gather index n_b + 1 (:274-277).  HBM-bound: B*L*D*s bytes written, every byte once, 16-byte vectors.
#include "scp_common.cuh"

namespace scp {

template <typename T>
__global__ void __launch_bounds__(256)
kw_splice_fwd_kernel(const float* __restrict__ keywords, const int64_t* __restrict__ kw_num, int64_t fixed_num,
                     const T* __restrict__ table, const T* __restrict__ pos, int64_t B, int64_t Kmax, int D, int L,
                     int64_t sot, int64_t eot, T* __restrict__ x, int64_t* __restrict__ eot_index) {
  constexpr int NE = Vec16<T>::NE;
  const int vec = D / NE;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= B * L * vec) return;
  const int v = (int)(i % vec);
  const int64_t row = i / vec;
  const int l = (int)(row % L);
  const int64_t b = row / L;
  int64_t n = kw_num ? kw_num[b] : fixed_num;
  n = n < 0 ? 0 : (n > Kmax ? Kmax : n);
  if (n > L - 2) n = L - 2;
  if (l == 0 && v == 0 && eot_index) eot_index[b] = n + 1;
  float f[NE], p[NE];
  Vec16<T>::unpack(*reinterpret_cast<const uint4*>(pos + (int64_t)l * D + v * NE), p);
  if (l >= 1 && l <= n) {
    const float* src = keywords + (b * Kmax + (l - 1)) * D + v * NE;
#pragma unroll
    for (int c = 0; c < NE / 4; ++c) {
      const float4 k4 = *reinterpret_cast<const float4*>(src + 4 * c);
      f[4 * c] = k4.x; f[4 * c + 1] = k4.y; f[4 * c + 2] = k4.z; f[4 * c + 3] = k4.w;
    }
    if (NE == 8) {  // the reference assigns the keywords INTO the embedding tensor: they take its dtype first
      const uint4 r = Vec16<T>::pack(f);
      Vec16<T>::unpack(r, f);
    }
  } else {
    const int64_t id = l == 0 ? sot : (l == n + 1 ? eot : 0);
    Vec16<T>::unpack(*reinterpret_cast<const uint4*>(table + id * D + v * NE), f);
  }
#pragma unroll
  for (int e = 0; e < NE; ++e) f[e] += p[e];
  st_stream16(x + row * D + v * NE, Vec16<T>::pack(f));
}

// g_keywords[b,j,:] = g_x[b,1+j,:] for j < n_b, else 0   (token table and positional embedding are frozen)
template <typename T>
__global__ void __launch_bounds__(256)
kw_splice_bwd_kernel(const T* __restrict__ gx, const int64_t* __restrict__ kw_num, int64_t fixed_num, int64_t B,
                     int64_t Kmax, int D, int L, float* __restrict__ g_keywords) {
  constexpr int NE = Vec16<T>::NE;
  const int vec = D / NE;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= B * Kmax * vec) return;
  const int v = (int)(i % vec);
  const int64_t row = i / vec;
  const int64_t j = row % Kmax, b = row / Kmax;
  int64_t n = kw_num ? kw_num[b] : fixed_num;
  n = n < 0 ? 0 : (n > Kmax ? Kmax : n);
  if (n > L - 2) n = L - 2;
  float f[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) f[e] = 0.f;
  if (j < n) Vec16<T>::unpack(ld_stream16(gx + (b * L + 1 + j) * D + v * NE), f);
  float* dst = g_keywords + row * D + v * NE;
#pragma unroll
  for (int c = 0; c < NE / 4; ++c)
    *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
}

// mask[b,j] = (j >= lens[b])   -- True marks padding (data_utils.py:17-20)
__global__ void keypadding_mask_kernel(const int64_t* __restrict__ lens, int64_t B, int64_t max_len,
                                       uint8_t* __restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * max_len) return;
  const int64_t b = i / max_len, j = i - b * max_len;
  mask[i] = j >= lens[b] ? 1 : 0;
}

static int check_splice(int64_t B, int64_t Kmax, int64_t D, int64_t L, int dtype) {
  SCP_CHECK_ARG(B > 0 && Kmax > 0 && D > 0 && L >= 3, "kw_splice: bad shape");
  SCP_CHECK_ARG(dtype >= SCP_F32 && dtype <= SCP_BF16, "kw_splice: bad dtype %d", dtype);
  const int ne = dtype == SCP_F32 ? 4 : 8;
  if (D % ne) return fail(SCP_ERR_UNSUPPORTED, "kw_splice: D must be a multiple of %d", ne);
  return SCP_OK;
}

}  // namespace scp

using namespace scp;

extern "C" int scp_kw_splice_fwd(const float* keywords, const int64_t* kw_num, int64_t fixed_num, const void* table,
                                 const void* pos_emb, int dtype, int64_t B, int64_t Kmax, int64_t D, int64_t L,
                                 int64_t sot_id, int64_t eot_id, void* x, int64_t* eot_index, scp_stream_t stream) {
  int rc = check_splice(B, Kmax, D, L, dtype);
  if (rc) return rc;
  SCP_CHECK_ARG(keywords && table && pos_emb && x, "kw_splice_fwd: null pointer");
  SCP_CHECK_ARG(sot_id >= 0 && eot_id >= 0, "kw_splice_fwd: negative token id");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int ne = dtype == SCP_F32 ? 4 : 8;
  const unsigned blocks = (unsigned)ceil_div(B * L * (D / ne), 256);
#define SCP_CASE(T)                                                                                              \
  kw_splice_fwd_kernel<T><<<blocks, 256, 0, s>>>(keywords, kw_num, fixed_num, reinterpret_cast<const T*>(table), \
                                                 reinterpret_cast<const T*>(pos_emb), B, Kmax, (int)D, (int)L,   \
                                                 sot_id, eot_id, reinterpret_cast<T*>(x), eot_index)
  if (dtype == SCP_F32) SCP_CASE(float);
  else if (dtype == SCP_F16) SCP_CASE(__half);
  else SCP_CASE(__nv_bfloat16);
#undef SCP_CASE
  SCP_CUDA_LAUNCH_CHECK("kw_splice_fwd");
  return SCP_OK;
}

extern "C" int scp_kw_splice_bwd(const void* g_x, int dtype, const int64_t* kw_num, int64_t fixed_num, int64_t B,
                                 int64_t Kmax, int64_t D, int64_t L, float* g_keywords, scp_stream_t stream) {
  int rc = check_splice(B, Kmax, D, L, dtype);
  if (rc) return rc;
  SCP_CHECK_ARG(g_x && g_keywords, "kw_splice_bwd: null pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int ne = dtype == SCP_F32 ? 4 : 8;
  const unsigned blocks = (unsigned)ceil_div(B * Kmax * (D / ne), 256);
  if (dtype == SCP_F32)
    kw_splice_bwd_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float*>(g_x), kw_num, fixed_num, B, Kmax, (int)D, (int)L, g_keywords);
  else if (dtype == SCP_F16)
    kw_splice_bwd_kernel<__half><<<blocks, 256, 0, s>>>(reinterpret_cast<const __half*>(g_x), kw_num, fixed_num, B, Kmax, (int)D, (int)L, g_keywords);
  else
    kw_splice_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(g_x), kw_num, fixed_num, B, Kmax, (int)D, (int)L, g_keywords);
  SCP_CUDA_LAUNCH_CHECK("kw_splice_bwd");
  return SCP_OK;
}

extern "C" int scp_keypadding_mask(const int64_t* lens, int64_t B, int64_t max_len, uint8_t* mask, scp_stream_t stream) {
  SCP_CHECK_ARG(lens && mask && B > 0 && max_len > 0, "keypadding_mask: bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  keypadding_mask_kernel<<<(unsigned)ceil_div(B * max_len, 256), 256, 0, s>>>(lens, B, max_len, mask);
  SCP_CUDA_LAUNCH_CHECK("keypadding_mask");
  return SCP_OK;
}
