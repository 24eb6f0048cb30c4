// How many LSU wavefronts does one warp-wide global access cost?  (ncu: l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum)
// Each kernel issues ONE access instruction per thread per iteration; lane stride selects scattered (one row per lane,
// `stride` bytes apart) or contiguous lanes.    nvcc -gencode arch=compute_100a,code=sm_100a -O3 lsu_wavefronts.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int BYTES, bool STORE>
__global__ void k(uint8_t* base, long long lane_stride, int iters) {
  const int lane = threadIdx.x & 31;
  uint8_t* p = base + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (1 << 20) + (size_t)lane * lane_stride;
  uint32_t r[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    uint8_t* q = p + (size_t)i * 32 * 32;  // next 1 KB (contiguous case) / next piece of the row
    if (STORE) {
      if (BYTES == 32) asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(q), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
      else asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(q), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
    } else {
      if (BYTES == 32) asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(q));
      else asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(q));
      acc += r[0] ^ r[3];
    }
  }
  if (acc == 0x1234567u) *reinterpret_cast<uint32_t*>(base) = acc;
}
int main() {
  uint8_t* buf; cudaMalloc(&buf, (size_t)64 << 20); cudaMemset(buf, 0, (size_t)64 << 20);
  // 8 warps x 4 blocks, 16 iterations: 32 * 16 = 512 warp-level instructions per kernel
  k<32, false><<<4, 256>>>(buf, 32, 16);      // LDG.256, lanes contiguous (1 KB per instruction)
  k<32, false><<<4, 256>>>(buf, 4096, 16);    // LDG.256, one row per lane
  k<16, false><<<4, 256>>>(buf, 16, 16);      // LDG.128, lanes contiguous (512 B per instruction)
  k<16, false><<<4, 256>>>(buf, 4096, 16);    // LDG.128, one row per lane
  k<32, true><<<4, 256>>>(buf, 32, 16);       // STG.256 contiguous
  k<32, true><<<4, 256>>>(buf, 4096, 16);     // STG.256 scattered
  k<16, true><<<4, 256>>>(buf, 16, 16);       // STG.128 contiguous
  k<16, true><<<4, 256>>>(buf, 4096, 16);     // STG.128 scattered
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
