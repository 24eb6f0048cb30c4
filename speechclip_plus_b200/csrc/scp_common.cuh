// Shared helpers for libscp_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/scp_b200.h"

namespace scp {

// ---- error reporting (thread-local detail string; codes are the SCP_ERR_* values) ---------------------------
void set_error_detail(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
int check_device_arch();  // SCP_OK on a compute-capability-10.x device

// Per-device helper stream + fork/join events (created lazily, never destroyed): lets one entry point run two
// independent kernels concurrently.  fork_to_side() makes the helper stream wait for everything enqueued on `main` so
// far and returns it; join_from_side() makes `main` wait for everything enqueued on the helper stream.  Both are
// capturable (event record / wait only).  Returns nullptr when the helper objects cannot be created.
cudaStream_t fork_to_side(cudaStream_t main);
int join_from_side(cudaStream_t main);

#define SCP_CHECK_ARG(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) return ::scp::fail(SCP_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define SCP_CUDA_LAUNCH_CHECK(name)                                                          \
  do {                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) return ::scp::fail(SCP_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
    ::scp::count_launch();                                                                   \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// Timing-ablation switches (SCP_DEBUG_ABLATE, SCP_VQ_S1_DBG, SCP_VQ_ST_DBG, SCP_PIPE_DEBUG) skip parts of a kernel and make
// its results INVALID; they are how the "MMAs alone / epilogue alone" figures in DESIGN.md were taken.  A regular build
// ignores them: only a library built with -DSCP_ABLATION (SCP_BUILD_ABLATION=1 python -m speechclip_plus_b200.build)
// reads the environment.
static inline int ablation_env(const char* name) {
#ifdef SCP_ABLATION
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
#else
  (void)name;
  return 0;
#endif
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// ---- device helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16-byte streaming load (read once: bypass L1 allocation) and store
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w));
}

// A 16-byte vector of NE elements of type T, converted to/from fp32 lanes.
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int NE = 4;
  __device__ __forceinline__ static void unpack(const uint4& u, float* f) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  __device__ __forceinline__ static uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <>
struct Vec16<__half> {
  static constexpr int NE = 8;
  __device__ __forceinline__ static void unpack(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ static uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int NE = 8;
  __device__ __forceinline__ static void unpack(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ static uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

static inline size_t dtype_size(int dt) { return dt == SCP_F32 ? 4 : 2; }

}  // namespace scp
