// N1 -- keyword batch-norm prologue of the cascaded branch: Kw_BatchNorm / Kw_BatchNorm_dynamic of the reference
// (avssl/module/speechclip_c_modules/kw_bn.py:97-164, :216-228), i.e. nn.BatchNorm1d statistics over the batch for every
// (keyword slot, channel) ["eachKw"] or over all keyword rows for every channel ["same" / dynamic], forward and backward.
//
// The tensors are tiny ((M,D) = 2048 x 512 fp32 = 4 MB, L2 resident): these kernels are latency-bound, so the design
// goal is few launches, coalesced 128-byte row segments and DETERMINISTIC results:
//   forward : kwbn_stats (per row-slice (n, mean, M2), two passes inside the slice) -> kwbn_finalize (Chan combination
//             of the slices in a fixed order, running-statistics update) -> kwbn_apply (16-byte vectors)
//   backward: kwbn_bwd_reduce (per-slice sums of g and g*xhat) -> kwbn_bwd_finalize (d_gamma, d_beta) -> kwbn_bwd_apply
// Rows m = b*K + k; group(m) = m % n_groups ("eachKw": n_groups = K; "same": n_groups = 1).  Parameters and running
// statistics are addressed in place as p[g*gstride + d*dstride] so that the reference's three parameter layouts
// (BatchNorm1d(D), BatchNorm1d(D*K) with index d*K + k, K stacked BatchNorm1d(D)) need no copies.
#include "scp_common.cuh"

namespace scp {

constexpr int kBnMaxSlices = 64;

struct BnLayout {
  int64_t gstride, dstride;
  int n_groups;
};

__device__ __forceinline__ bool bn_row_valid(const uint8_t* row_valid, int64_t m) { return !row_valid || row_valid[m]; }

// grid (ceil(D/32), n_groups, n_slices), block (32, 8): thread (x, y) owns channel d = 32*bx + x and the rows
// y, y+8, ... of its slice.  Two passes over the slice (it is a few KB, L1 resident): slice mean, then centred squares.
__global__ void __launch_bounds__(256)
kwbn_stats_kernel(const float* __restrict__ x, int64_t M, int D, BnLayout lay, const uint8_t* __restrict__ row_valid,
                  int n_slices, float* __restrict__ part /* (n_groups, n_slices, 3, D): n, mean, M2 */) {
  __shared__ float s_a[8][32], s_b[8][32];
  const int d = blockIdx.x * 32 + threadIdx.x;
  const int g = blockIdx.y, sl = blockIdx.z;
  const int64_t rows_g = (M - g + lay.n_groups - 1) / lay.n_groups;  // rows m = g + j*n_groups < M
  const int64_t j0 = rows_g * sl / n_slices, j1 = rows_g * (sl + 1) / n_slices;
  float sum = 0.f, cnt = 0.f;
  if (d < D)
    for (int64_t j = j0 + threadIdx.y; j < j1; j += 8) {
      const int64_t m = g + j * lay.n_groups;
      if (bn_row_valid(row_valid, m)) { sum += x[m * D + d]; cnt += 1.f; }
    }
  s_a[threadIdx.y][threadIdx.x] = sum;
  s_b[threadIdx.y][threadIdx.x] = cnt;
  __syncthreads();
  float tot = 0.f, n = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { tot += s_a[i][threadIdx.x]; n += s_b[i][threadIdx.x]; }
  const float mean = n > 0.f ? tot / n : 0.f;
  __syncthreads();
  float m2 = 0.f;
  if (d < D)
    for (int64_t j = j0 + threadIdx.y; j < j1; j += 8) {
      const int64_t m = g + j * lay.n_groups;
      if (bn_row_valid(row_valid, m)) { const float c = x[m * D + d] - mean; m2 = fmaf(c, c, m2); }
    }
  s_a[threadIdx.y][threadIdx.x] = m2;
  __syncthreads();
  if (threadIdx.y == 0 && d < D) {
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) q += s_a[i][threadIdx.x];
    float* o = part + (((int64_t)g * n_slices + sl) * 3) * D + d;
    o[0] = n; o[D] = mean; o[2 * D] = q;
  }
}

// thread <-> (g, d): combine the slices (Chan et al., fixed order), write mean / rstd, update the running statistics
// exactly like torch.nn.BatchNorm1d in training mode (biased variance for normalisation, unbiased for running_var).
// eval mode (training == 0): mean / rstd come from the running statistics and `part` is not read.
__global__ void kwbn_finalize_kernel(const float* __restrict__ part, int n_slices, int D, BnLayout lay, int training,
                                     float momentum, float eps, float* __restrict__ running_mean,
                                     float* __restrict__ running_var, float* __restrict__ save_mean,
                                     float* __restrict__ save_rstd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)lay.n_groups * D) return;
  const int g = (int)(i / D), d = (int)(i - (int64_t)g * D);
  const int64_t pi = g * lay.gstride + d * lay.dstride;
  if (!training) {
    save_mean[i] = running_mean[pi];
    save_rstd[i] = 1.0f / sqrtf(running_var[pi] + eps);
    return;
  }
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int sl = 0; sl < n_slices; ++sl) {
    const float* o = part + (((int64_t)g * n_slices + sl) * 3) * D + d;
    const float nb = o[0], mb = o[D], qb = o[2 * D];
    if (nb > 0.f) {
      const float nn = n + nb, delta = mb - mean;
      mean += delta * (nb / nn);
      m2 += qb + delta * delta * (n * nb / nn);
      n = nn;
    }
  }
  const float var = n > 0.f ? m2 / n : 0.f;
  save_mean[i] = mean;
  save_rstd[i] = 1.0f / sqrtf(var + eps);
  if (running_mean) {
    const float unbiased = n > 1.f ? m2 / (n - 1.f) : var;
    running_mean[pi] = (1.f - momentum) * running_mean[pi] + momentum * mean;
    running_var[pi] = (1.f - momentum) * running_var[pi] + momentum * unbiased;
  }
}

// thread <-> 4 consecutive channels of one row: y = (x - mean) * rstd * gamma + beta ; invalid rows pass through
__global__ void kwbn_apply_kernel(const float* __restrict__ x, int64_t M, int D, BnLayout lay,
                                  const uint8_t* __restrict__ row_valid, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, const float* __restrict__ save_mean,
                                  const float* __restrict__ save_rstd, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int vec = D >> 2;
  if (i >= M * vec) return;
  const int64_t m = i / vec;
  const int d0 = (int)(i - m * vec) * 4;
  const float4 xv = *reinterpret_cast<const float4*>(x + m * D + d0);
  float o[4] = {xv.x, xv.y, xv.z, xv.w};
  if (bn_row_valid(row_valid, m)) {
    const int g = (int)(m % lay.n_groups);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t pi = g * lay.gstride + (int64_t)(d0 + e) * lay.dstride;
      const float w = gamma ? gamma[pi] : 1.f, b = beta ? beta[pi] : 0.f;
      o[e] = (o[e] - save_mean[(int64_t)g * D + d0 + e]) * save_rstd[(int64_t)g * D + d0 + e] * w + b;
    }
  }
  *reinterpret_cast<float4*>(y + m * D + d0) = make_float4(o[0], o[1], o[2], o[3]);
}

// backward stage 1: per-slice sum_m g and sum_m g * xhat      (same decomposition as kwbn_stats_kernel)
__global__ void __launch_bounds__(256)
kwbn_bwd_reduce_kernel(const float* __restrict__ gy, const float* __restrict__ x, int64_t M, int D, BnLayout lay,
                       const uint8_t* __restrict__ row_valid, const float* __restrict__ save_mean,
                       const float* __restrict__ save_rstd, int n_slices,
                       float* __restrict__ part /* (n_groups, n_slices, 3, D): sum g, sum g*xhat, rows */) {
  __shared__ float s_a[8][32], s_b[8][32], s_c[8][32];
  const int d = blockIdx.x * 32 + threadIdx.x;
  const int g = blockIdx.y, sl = blockIdx.z;
  const int64_t rows_g = (M - g + lay.n_groups - 1) / lay.n_groups;
  const int64_t j0 = rows_g * sl / n_slices, j1 = rows_g * (sl + 1) / n_slices;
  float sg = 0.f, sgx = 0.f, cnt = 0.f;
  if (d < D) {
    const float mean = save_mean[(int64_t)g * D + d], rstd = save_rstd[(int64_t)g * D + d];
    for (int64_t j = j0 + threadIdx.y; j < j1; j += 8) {
      const int64_t m = g + j * lay.n_groups;
      if (bn_row_valid(row_valid, m)) {
        const float gv = gy[m * D + d];
        sg += gv;
        sgx = fmaf(gv, (x[m * D + d] - mean) * rstd, sgx);
        cnt += 1.f;
      }
    }
  }
  s_a[threadIdx.y][threadIdx.x] = sg;
  s_b[threadIdx.y][threadIdx.x] = sgx;
  s_c[threadIdx.y][threadIdx.x] = cnt;
  __syncthreads();
  if (threadIdx.y == 0 && d < D) {
    float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += s_a[i][threadIdx.x]; b += s_b[i][threadIdx.x]; c += s_c[i][threadIdx.x]; }
    float* o = part + (((int64_t)g * n_slices + sl) * 3) * D + d;
    o[0] = a; o[D] = b; o[2 * D] = c;
  }
}

// backward stage 2: thread <-> (g, d): d_beta = sum g, d_gamma = sum g*xhat (fixed slice order); the thread of
// channel 0 also totals the number of rows that entered the statistics
__global__ void kwbn_bwd_finalize_kernel(const float* __restrict__ part, int n_slices, int D, BnLayout lay,
                                         float* __restrict__ sums /* (n_groups, 2, D): sum g, sum g*xhat */,
                                         float* __restrict__ counts /* (n_groups,) */,
                                         float* __restrict__ g_gamma, float* __restrict__ g_beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)lay.n_groups * D) return;
  const int g = (int)(i / D), d = (int)(i - (int64_t)g * D);
  float a = 0.f, b = 0.f, c = 0.f;
  for (int sl = 0; sl < n_slices; ++sl) {
    const float* o = part + (((int64_t)g * n_slices + sl) * 3) * D + d;
    a += o[0]; b += o[D]; c += o[2 * D];
  }
  if (d == 0) counts[g] = fmaxf(c, 1.f);
  sums[((int64_t)g * 2) * D + d] = a;
  sums[((int64_t)g * 2 + 1) * D + d] = b;
  const int64_t pi = g * lay.gstride + d * lay.dstride;
  if (g_beta) g_beta[pi] = a;
  if (g_gamma) g_gamma[pi] = b;
}

// backward stage 3: dx = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)) in training mode, gamma*rstd*g in eval mode;
// rows outside the statistics pass the gradient through unchanged
__global__ void kwbn_bwd_apply_kernel(const float* __restrict__ gy, const float* __restrict__ x, int64_t M, int D,
                                      BnLayout lay, const uint8_t* __restrict__ row_valid,
                                      const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                      const float* __restrict__ save_rstd, const float* __restrict__ sums,
                                      const float* __restrict__ counts /* (n_groups,) rows in the statistics */,
                                      int training, float* __restrict__ gx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int vec = D >> 2;
  if (i >= M * vec) return;
  const int64_t m = i / vec;
  const int d0 = (int)(i - m * vec) * 4;
  const float4 gv = *reinterpret_cast<const float4*>(gy + m * D + d0);
  float o[4] = {gv.x, gv.y, gv.z, gv.w};
  if (bn_row_valid(row_valid, m)) {
    const int g = (int)(m % lay.n_groups);
    const float4 xv = *reinterpret_cast<const float4*>(x + m * D + d0);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const float inv_n = 1.0f / counts[g];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t si = (int64_t)g * D + d0 + e;
      const float rstd = save_rstd[si];
      const float w = gamma ? gamma[g * lay.gstride + (int64_t)(d0 + e) * lay.dstride] : 1.f;
      if (training) {
        const float xhat = (xs[e] - save_mean[si]) * rstd;
        const float sg = sums[((int64_t)g * 2) * D + d0 + e], sgx = sums[((int64_t)g * 2 + 1) * D + d0 + e];
        o[e] = w * rstd * (o[e] - sg * inv_n - xhat * sgx * inv_n);
      } else {
        o[e] = w * rstd * o[e];
      }
    }
  }
  *reinterpret_cast<float4*>(gx + m * D + d0) = make_float4(o[0], o[1], o[2], o[3]);
}

static int bn_slices(int64_t M, int n_groups) {
  const int64_t rows_g = ceil_div(M, n_groups);
  int s = (int)ceil_div(rows_g, 64);  // ~64 rows (8 per thread) per slice
  if (s < 1) s = 1;
  if (s > kBnMaxSlices) s = kBnMaxSlices;
  return s;
}

static int check_bn(const void* x, int64_t M, int64_t D, int n_groups, int64_t gstride, int64_t dstride) {
  SCP_CHECK_ARG(x != nullptr, "kwbn: null input");
  SCP_CHECK_ARG(M > 0 && D > 0 && D % 4 == 0, "kwbn: M=%lld D=%lld (D must be a multiple of 4)", (long long)M, (long long)D);
  SCP_CHECK_ARG(n_groups >= 1 && M % n_groups == 0, "kwbn: M=%lld is not a multiple of n_groups=%d", (long long)M, n_groups);
  SCP_CHECK_ARG(dstride >= 1 && gstride >= 0, "kwbn: bad parameter strides");
  SCP_CHECK_ARG(n_groups <= 65535, "kwbn: too many groups");
  return SCP_OK;
}

}  // namespace scp

using namespace scp;

extern "C" size_t scp_kwbn_workspace_bytes(int64_t M, int64_t D, int n_groups) {
  const int s = bn_slices(M, n_groups < 1 ? 1 : n_groups);
  // forward partials (3 per slice) / backward partials (2 per slice) share the front; then sums (2) and counts
  return ((size_t)n_groups * s * 3 * D + (size_t)n_groups * 2 * D + (size_t)n_groups + 64) * sizeof(float);
}

extern "C" int scp_kwbn_fwd(const float* x, int64_t M, int64_t D, int n_groups, int64_t gstride, int64_t dstride,
                            const uint8_t* row_valid, const float* gamma, const float* beta, float* running_mean,
                            float* running_var, int training, float momentum, float eps, float* y, float* save_mean,
                            float* save_rstd, void* workspace, size_t workspace_bytes, scp_stream_t stream) {
  int rc = check_bn(x, M, D, n_groups, gstride, dstride);
  if (rc) return rc;
  SCP_CHECK_ARG(y && save_mean && save_rstd && workspace, "kwbn_fwd: null pointer");
  SCP_CHECK_ARG(training || (running_mean && running_var), "kwbn_fwd: eval mode needs the running statistics");
  if (workspace_bytes < scp_kwbn_workspace_bytes(M, D, n_groups))
    return fail(SCP_ERR_WORKSPACE, "kwbn_fwd: workspace %zu < %zu", workspace_bytes, scp_kwbn_workspace_bytes(M, D, n_groups));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const BnLayout lay{gstride, dstride, n_groups};
  const int slices = bn_slices(M, n_groups);
  float* part = reinterpret_cast<float*>(workspace);
  if (training) {
    const dim3 grid((unsigned)ceil_div(D, 32), (unsigned)n_groups, (unsigned)slices), block(32, 8);
    kwbn_stats_kernel<<<grid, block, 0, s>>>(x, M, (int)D, lay, row_valid, slices, part);
    SCP_CUDA_LAUNCH_CHECK("kwbn_stats");
  }
  kwbn_finalize_kernel<<<(unsigned)ceil_div((int64_t)n_groups * D, 256), 256, 0, s>>>(
      part, slices, (int)D, lay, training, momentum, eps, running_mean, running_var, save_mean, save_rstd);
  SCP_CUDA_LAUNCH_CHECK("kwbn_finalize");
  kwbn_apply_kernel<<<(unsigned)ceil_div(M * (D / 4), 256), 256, 0, s>>>(x, M, (int)D, lay, row_valid, gamma, beta,
                                                                        save_mean, save_rstd, y);
  SCP_CUDA_LAUNCH_CHECK("kwbn_apply");
  return SCP_OK;
}

extern "C" int scp_kwbn_bwd(const float* g_y, const float* x, int64_t M, int64_t D, int n_groups, int64_t gstride,
                            int64_t dstride, const uint8_t* row_valid, const float* gamma, const float* save_mean,
                            const float* save_rstd, int training, float* g_x, float* g_gamma, float* g_beta,
                            void* workspace, size_t workspace_bytes, scp_stream_t stream) {
  int rc = check_bn(x, M, D, n_groups, gstride, dstride);
  if (rc) return rc;
  SCP_CHECK_ARG(g_y && g_x && save_mean && save_rstd && workspace, "kwbn_bwd: null pointer");
  if (workspace_bytes < scp_kwbn_workspace_bytes(M, D, n_groups))
    return fail(SCP_ERR_WORKSPACE, "kwbn_bwd: workspace %zu < %zu", workspace_bytes, scp_kwbn_workspace_bytes(M, D, n_groups));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const BnLayout lay{gstride, dstride, n_groups};
  const int slices = bn_slices(M, n_groups);
  float* part = reinterpret_cast<float*>(workspace);
  float* sums = part + (size_t)n_groups * slices * 3 * D;
  float* counts = sums + (size_t)n_groups * 2 * D;
  const dim3 grid((unsigned)ceil_div(D, 32), (unsigned)n_groups, (unsigned)slices), block(32, 8);
  kwbn_bwd_reduce_kernel<<<grid, block, 0, s>>>(g_y, x, M, (int)D, lay, row_valid, save_mean, save_rstd, slices, part);
  SCP_CUDA_LAUNCH_CHECK("kwbn_bwd_reduce");
  kwbn_bwd_finalize_kernel<<<(unsigned)ceil_div((int64_t)n_groups * D, 256), 256, 0, s>>>(
      part, slices, (int)D, lay, sums, counts, g_gamma, g_beta);
  SCP_CUDA_LAUNCH_CHECK("kwbn_bwd_finalize");
  kwbn_bwd_apply_kernel<<<(unsigned)ceil_div(M * (D / 4), 256), 256, 0, s>>>(
      g_y, x, M, (int)D, lay, row_valid, gamma, save_mean, save_rstd, sums, counts, training, g_x);
  SCP_CUDA_LAUNCH_CHECK("kwbn_bwd_apply");
  return SCP_OK;
}
