// Packed gradient exchange + fused Adam step for the hot path's own trainable parameters
// (audio_encoder.weightedsum_layer.weights, criterion.temperature, and whatever else the caller registers): the tail of
// the data-parallel step of KWClipBase (avssl/model/kwClip.py:149-193 gather point, :636-668 single Adam group,
// config/speechCLIP+/model_base/spchclip_c+.yaml:119-123 lr 1e-4, weight_decay 1e-6).
//
// The reference lets nn.DataParallel reduce_add_coalesced every replica's gradients onto GPU 0 and runs torch.optim.Adam
// there, one multi-tensor launch chain per step.  Here each process packs its few small gradients into one flat buffer
// (one launch), the caller all-reduces that buffer over NCCL (one collective), and one launch applies Adam to every
// registered tensor in place.  The step counter lives on the device so that a captured CUDA graph advances it on replay.
#include "scp_common.cuh"

namespace scp {

struct PackList {
  const float* src[SCP_MAX_PACKED];
  float* dst[SCP_MAX_PACKED];
  int64_t offset[SCP_MAX_PACKED + 1];  // prefix sums of the element counts
  int n;
};

__device__ __forceinline__ int pack_find(const PackList& pl, int64_t e) {
  int t = 0;
#pragma unroll
  for (int i = 1; i < SCP_MAX_PACKED; ++i) t += (i < pl.n && e >= pl.offset[i]) ? 1 : 0;
  return t;
}

// packed[e] = scale * grads[t][e - offset[t]]   (a null gradient pointer packs zeros)
__global__ void grad_pack_kernel(PackList pl, float scale, float* __restrict__ packed) {
  const int64_t total = pl.offset[pl.n];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int t = pack_find(pl, e);
    const float* g = pl.src[t];
    packed[e] = g ? g[e - pl.offset[t]] * scale : 0.f;
  }
}

// torch.optim.Adam (no amsgrad, L2 weight decay): g += wd p ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).  `step` is read, and advanced by thread 0 of block 0 after
// every block has read it (single launch, grid of one block for the sizes this path has).
__global__ void adam_packed_kernel(PackList pl, const float* __restrict__ packed, int n_shards, int64_t shard_stride,
                                   float grad_scale,
                                   float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                   int64_t* __restrict__ step, const float* __restrict__ lr_ptr, float lr, float beta1,
                                   float beta2, float eps, float weight_decay) {
  const int64_t total = pl.offset[pl.n];
  const int64_t t_now = *step + 1;
  const float rate = lr_ptr ? *lr_ptr : lr;
  const float bc1 = 1.0f - powf(beta1, (float)t_now);
  const float bc2 = 1.0f - powf(beta2, (float)t_now);
  const float step_size = rate / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
    const int t = pack_find(pl, e);
    float* p = pl.dst[t] + (e - pl.offset[t]);
    float g = 0.f;
    for (int r = 0; r < n_shards; ++r) g += packed[(int64_t)r * shard_stride + e];  // fixed order: every rank identical
    g *= grad_scale;
    const float w = *p;
    g = fmaf(weight_decay, w, g);
    const float m = beta1 * exp_avg[e] + (1.0f - beta1) * g;
    const float v = beta2 * exp_avg_sq[e] + (1.0f - beta2) * g * g;
    exp_avg[e] = m;
    exp_avg_sq[e] = v;
    *p = w - step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
  }
  __syncthreads();
  if (threadIdx.x == 0) *step = t_now;
}

static int make_pack_list(PackList* pl, const float* const* src, float* const* dst, const int64_t* sizes, int n) {
  SCP_CHECK_ARG(n >= 1 && n <= SCP_MAX_PACKED && sizes, "packed parameters: 1..%d tensors", SCP_MAX_PACKED);
  pl->n = n;
  pl->offset[0] = 0;
  for (int i = 0; i < SCP_MAX_PACKED; ++i) {
    pl->src[i] = (src && i < n) ? src[i] : nullptr;
    pl->dst[i] = (dst && i < n) ? dst[i] : nullptr;
    if (i < n) {
      SCP_CHECK_ARG(sizes[i] > 0, "packed parameters: tensor %d has %lld elements", i, (long long)sizes[i]);
      pl->offset[i + 1] = pl->offset[i] + sizes[i];
    } else {
      pl->offset[i + 1] = pl->offset[i];
    }
  }
  return SCP_OK;
}

}  // namespace scp

using namespace scp;

extern "C" int scp_grad_pack(const float* const* grads, const int64_t* sizes, int n, float scale, float* packed,
                             scp_stream_t stream) {
  SCP_CHECK_ARG(grads && packed, "grad_pack: null pointer");
  PackList pl;
  int rc = make_pack_list(&pl, grads, nullptr, sizes, n);
  if (rc) return rc;
  const int64_t total = pl.offset[n];
  grad_pack_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 64), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      pl, scale, packed);
  SCP_CUDA_LAUNCH_CHECK("grad_pack");
  return SCP_OK;
}

extern "C" int scp_adam_packed(float* const* params, const int64_t* sizes, int n, const float* packed_grads,
                               int n_shards, int64_t shard_stride, float grad_scale, float* exp_avg, float* exp_avg_sq,
                               int64_t* step,
                               const float* lr_device, float lr, float beta1, float beta2, float eps,
                               float weight_decay, scp_stream_t stream) {
  SCP_CHECK_ARG(params && packed_grads && exp_avg && exp_avg_sq && step, "adam_packed: null pointer");
  SCP_CHECK_ARG(n_shards >= 1 && n_shards <= 1024, "adam_packed: n_shards %d", n_shards);
  PackList pl;
  int rc = make_pack_list(&pl, nullptr, params, sizes, n);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) SCP_CHECK_ARG(params[i], "adam_packed: params[%d] null", i);
  SCP_CHECK_ARG(pl.offset[n] <= (1 << 20), "adam_packed: %lld elements (this entry point is for the path's small tensors)",
                (long long)pl.offset[n]);
  SCP_CHECK_ARG(n_shards == 1 || shard_stride >= pl.offset[n], "adam_packed: shard stride %lld < %lld elements",
                (long long)shard_stride, (long long)pl.offset[n]);
  adam_packed_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pl, packed_grads, n_shards, shard_stride, grad_scale, exp_avg,
                                                                             exp_avg_sq, step, lr_device, lr, beta1, beta2,
                                                                             eps, weight_decay);
  SCP_CUDA_LAUNCH_CHECK("adam_packed");
  return SCP_OK;
}
