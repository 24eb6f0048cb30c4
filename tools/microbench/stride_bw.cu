// Micro-benchmark: DRAM throughput of the row-wise epilogue access pattern on an (M, Vp) fp16 matrix (M = 2048, Vp = 49408).
// A warp owns 32 rows (lane <-> row) and a range of 32-column chunks; per chunk every lane moves its 64 B (two 32-byte
// accesses), exactly like the sweep epilogues.  Three layouts of the same matrix:
//   0 row-major    : addr(m, c) = m * 2 Vp + 64 c                                   (pieces 2 Vp bytes apart)
//   1 tile-major   : 64 KB tiles [128 rows][256 cols]                                (pieces 512 B apart)
//   2 chunk-major  : 64 KB tiles [8 chunks][128 rows][32 cols]                       (a warp's 32 pieces are 2 KB contiguous)
// mode 0 = read, 1 = write, 2 = read one matrix + write another.       nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr long long M = 2048, Vp = 49408, NCH = Vp / 32, NT = Vp / 256;
__device__ __forceinline__ long long addr16(int layout, long long m, long long c) {  // in 16-byte units
  long long b;
  if (layout == 0) b = m * Vp * 2 + c * 64;
  else {
    const long long tile = (m >> 7) * NT + (c >> 3);
    b = layout == 1 ? tile * 65536 + (m & 127) * 512 + (c & 7) * 64 : tile * 65536 + (c & 7) * 8192 + (m & 127) * 64;
  }
  return b >> 4;
}
// row-major, but every lane moves PIECE bytes (64 * CPP) of its row per step: CPP consecutive chunks at once
template <int CPP>
__global__ void kp(const uint4* __restrict__ src, uint4* __restrict__ dst, int mode, int vsplit, unsigned* sink) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long g = warp / vsplit, r = warp % vsplit;
  if (g >= M / 32) return;
  const long long per = ((NCH + vsplit - 1) / vsplit + CPP - 1) / CPP * CPP;
  const long long c0 = r * per, c1 = min(NCH, c0 + per);
  const long long m = g * 32 + lane;
  unsigned acc = 0;
  for (long long c = c0; c + CPP <= c1; c += CPP) {
    const long long o = addr16(0, m, c);
    uint4 v[4 * CPP];
    if (mode != 1) {
#pragma unroll
      for (int i = 0; i < 4 * CPP; ++i) v[i] = __ldcs(src + o + i);
#pragma unroll
      for (int i = 0; i < 4 * CPP; ++i) acc += v[i].x ^ v[i].w;
    }
    if (mode != 0) {
      const uint4 w = make_uint4(acc, lane, (unsigned)c, 1u);
#pragma unroll
      for (int i = 0; i < 4 * CPP; ++i) dst[o + i] = w;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}
__global__ void k(const uint4* __restrict__ src, uint4* __restrict__ dst, int layout, int mode, int vsplit, unsigned* sink) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long g = warp / vsplit, r = warp % vsplit;
  if (g >= M / 32) return;
  const long long per = (NCH + vsplit - 1) / vsplit;
  const long long c0 = r * per, c1 = min(NCH, c0 + per);
  const long long m = g * 32 + lane;
  unsigned acc = 0;
  for (long long c = c0; c < c1; ++c) {
    const long long o = addr16(layout, m, c);
    if (mode != 1) {
      const uint4 a = __ldcs(src + o), b = __ldcs(src + o + 1), cc = __ldcs(src + o + 2), d = __ldcs(src + o + 3);
      acc += a.x ^ b.y ^ cc.z ^ d.w;
    }
    if (mode != 0) {
      const uint4 v = make_uint4(acc, lane, (unsigned)c, 1u);
      dst[o] = v; dst[o + 1] = v; dst[o + 2] = v; dst[o + 3] = v;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}
int main() {
  const size_t bytes = (size_t)M * Vp * 2;
  uint4 *a, *b; unsigned* sink;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 4);
  cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[3] = {"row-major", "tile-major", "chunk-major"};
  for (int vsplit : {19, 37})
  for (int layout = 0; layout < 3; ++layout) for (int mode = 0; mode < 3; ++mode) {
    const long long n_warps = (M / 32) * vsplit;
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      k<<<(unsigned)((n_warps + 7) / 8), 256>>>(a, b, layout, mode, vsplit, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double moved = (double)bytes * (mode == 2 ? 2 : 1);
    printf("vsplit %2d %-12s %s: %.3f ms  %.2f TB/s\n", vsplit, names[layout], mode == 0 ? "read " : mode == 1 ? "write" : "r + w", best, moved / best / 1e9);
  }
  for (int cpp : {2, 4}) for (int mode = 0; mode < 3; ++mode) {
    const int vsplit = 19;
    const long long n_warps = (M / 32) * vsplit;
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (cpp == 2) kp<2><<<(unsigned)((n_warps + 7) / 8), 256>>>(a, b, mode, vsplit, sink);
      else kp<4><<<(unsigned)((n_warps + 7) / 8), 256>>>(a, b, mode, vsplit, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double moved = (double)bytes * (mode == 2 ? 2 : 1);
    printf("vsplit 19 row-major, %3d-byte pieces %s: %.3f ms  %.2f TB/s (approx.)\n", 64 * cpp, mode == 0 ? "read " : mode == 1 ? "write" : "r + w", best, moved / best / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
