// One-shot all-gather over NVLink peer memory for the path's small exchanges (G0: the packed normalised features of
// avssl/model/kwClip.py:149-169, the (3, n) InfoNCE statistics of the sharded forward, the packed parameter gradients).
//
// These payloads are 56 B .. 1 MB per rank: an NCCL collective inside the step's CUDA graph costs 25-50 us of latency each
// at 8 ranks (measured: the loss group grows from 0.08 ms at 1 GPU to 0.22 ms at 8, the optimiser from 10 to 37 us), which is
// what bends the 1 -> 8 GPU curve.  Here every rank PUSHES its payload straight into the gather buffer of every peer with
// ordinary vector stores on peer-mapped pointers (NVSwitch gives every pair full bandwidth), then raises a flag in each
// peer; the same launch waits for the flags of all peers.  One launch + one collect launch, no host involvement,
// capturable in a CUDA graph (the epoch lives on the device).
//
// Buffer (identical on every rank, allocated as symmetric memory by the caller and exchanged once):
//   [parity 0: world slots of slot_stride bytes][parity 1: world slots][flags: world x u32, one per SOURCE rank]
// Step k uses parity k & 1: a fast peer may already push step k+1 while this rank still reads step k; it cannot push
// step k+2 before this rank has pushed step k+1, i.e. finished reading step k.
// Ordering: payload stores -> __threadfence_system() -> flag store (release.sys); reader: flag load (acquire.sys) -> payload.
#include "scp_common.cuh"

namespace scp {

__device__ __forceinline__ uint64_t p2p_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// state[0] = epoch of the last completed exchange, state[1] = block ticket
__global__ void __launch_bounds__(256)
p2p_allgather_kernel(const uint8_t* __restrict__ src, size_t nbytes, uint8_t* const* __restrict__ peer_bufs, int rank,
                     int world, size_t slot_stride, size_t flag_offset, uint32_t* __restrict__ state) {
  __shared__ bool s_last;
  const uint32_t epoch = state[0] + 1u;  // state[0] is advanced by the last block only after every block has read it (ticket)
  const size_t parity_off = (size_t)(epoch & 1u) * world * slot_stride;
  const size_t n16 = nbytes >> 4;
  const uint4* s16 = reinterpret_cast<const uint4*>(src);
  for (int pi = 0; pi < world; ++pi) {
    const int peer = (rank + pi) % world;  // start with the local copy, spread the ranks over the links
    uint4* d16 = reinterpret_cast<uint4*>(peer_bufs[peer] + parity_off + (size_t)rank * slot_stride);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) d16[i] = s16[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&state[1], 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();  // the other blocks' fenced stores happen-before their ticket, hence before the flags below
  if (threadIdx.x < world) {
    uint32_t* flag = reinterpret_cast<uint32_t*>(peer_bufs[threadIdx.x] + flag_offset) + rank;
    st_release_sys(flag, epoch);
  }
  if (threadIdx.x < world) {
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(peer_bufs[rank] + flag_offset) + threadIdx.x;
    const uint64_t t0 = p2p_timer_ns();
    // epochs only grow; a peer that is already one exchange ahead shows epoch + 1 (wrap-safe comparison)
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (p2p_timer_ns() - t0 > 20000000000ull) {
        printf("scp: p2p all-gather timed out waiting for rank %d (epoch %u)\n", (int)threadIdx.x, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    state[1] = 0u;
    __threadfence();
    state[0] = epoch;
  }
}

// out (world, nbytes) <- the parity half of the local buffer that the last exchange filled
__global__ void __launch_bounds__(256)
p2p_collect_kernel(const uint8_t* __restrict__ local_buf, size_t nbytes, int world, size_t slot_stride,
                   const uint32_t* __restrict__ state, uint8_t* __restrict__ out) {
  const uint32_t epoch = state[0];
  const uint8_t* base = local_buf + (size_t)(epoch & 1u) * world * slot_stride;
  const size_t n16 = nbytes >> 4;
  for (int r = blockIdx.y; r < world; r += gridDim.y) {
    const uint4* s16 = reinterpret_cast<const uint4*>(base + (size_t)r * slot_stride);
    uint4* d16 = reinterpret_cast<uint4*>(out + (size_t)r * nbytes);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) d16[i] = s16[i];
  }
}

// Segment-major collect: the payload of every rank is a sequence of n_seg segments (e.g. [feature 0 | feature 1 | ids]);
// out holds, segment after segment, the world copies of that segment in rank order -- every segment is then ONE contiguous
// (world * n, ...) tensor for the caller, no strided copies afterwards.
struct P2pSegs {
  int n;
  unsigned long long off[4];   // byte offset of segment i inside a rank's payload (off[n] = nbytes)
  unsigned long long end[4];
};
__global__ void __launch_bounds__(256)
p2p_collect_seg_kernel(const uint8_t* __restrict__ local_buf, size_t nbytes, int world, size_t slot_stride,
                       const uint32_t* __restrict__ state, uint8_t* __restrict__ out, P2pSegs segs) {
  const uint32_t epoch = state[0];
  const uint8_t* base = local_buf + (size_t)(epoch & 1u) * world * slot_stride;
  const size_t n16 = nbytes >> 4;
  for (int r = blockIdx.y; r < world; r += gridDim.y) {
    const uint4* s16 = reinterpret_cast<const uint4*>(base + (size_t)r * slot_stride);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
      const unsigned long long b = (unsigned long long)i << 4;
      int sg = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (k < segs.n && b >= segs.off[k]) sg = k;
      const unsigned long long size = segs.end[sg] - segs.off[sg];
      // segment sg starts at world * off[sg] in `out`; rank r's copy follows r * size further on
      *reinterpret_cast<uint4*>(out + (size_t)world * segs.off[sg] + (size_t)r * size + (b - segs.off[sg])) = s16[i];
    }
  }
}

}  // namespace scp

using namespace scp;

extern "C" size_t scp_p2p_buffer_bytes(int world, size_t nbytes_per_rank) {
  const size_t slot = (nbytes_per_rank + 255) & ~size_t(255);
  return 2 * (size_t)world * slot + 256 + (((size_t)world * 4 + 255) & ~size_t(255));
}

extern "C" int scp_p2p_allgather(const void* src, size_t nbytes, void* const* peer_bufs_device, void* local_buf, int rank,
                                 int world, size_t nbytes_capacity, uint32_t* state, void* out, scp_stream_t stream) {
  SCP_CHECK_ARG(src && peer_bufs_device && local_buf && state && out, "p2p_allgather: null pointer");
  SCP_CHECK_ARG(world >= 1 && world <= 64 && rank >= 0 && rank < world, "p2p_allgather: bad rank / world");
  SCP_CHECK_ARG(nbytes > 0 && nbytes % 16 == 0 && nbytes <= nbytes_capacity, "p2p_allgather: payload %zu B (multiple of 16, <= %zu)",
                nbytes, nbytes_capacity);
  SCP_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "p2p_allgather: src / out must be 16-byte aligned");
  const size_t slot = (nbytes_capacity + 255) & ~size_t(255);
  const size_t flag_offset = 2 * (size_t)world * slot + 256;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>(32, (nbytes / 16 + 255) / 256));
  p2p_allgather_kernel<<<blocks, 256, 0, s>>>(static_cast<const uint8_t*>(src), nbytes,
                                              reinterpret_cast<uint8_t* const*>(peer_bufs_device), rank, world, slot,
                                              flag_offset, state);
  SCP_CUDA_LAUNCH_CHECK("p2p_allgather");
  dim3 grid((unsigned)std::max<size_t>(1, std::min<size_t>(16, (nbytes / 16 + 255) / 256)), (unsigned)std::min(world, 8));
  p2p_collect_kernel<<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(local_buf), nbytes, world, slot, state,
                                          static_cast<uint8_t*>(out));
  SCP_CUDA_LAUNCH_CHECK("p2p_collect");
  return SCP_OK;
}

extern "C" int scp_p2p_allgather_segments(const void* src, size_t nbytes, void* const* peer_bufs_device, void* local_buf,
                                          int rank, int world, size_t nbytes_capacity, uint32_t* state,
                                          const int64_t* seg_bytes, int n_seg, void* out, scp_stream_t stream) {
  SCP_CHECK_ARG(src && peer_bufs_device && local_buf && state && out && seg_bytes, "p2p_allgather_segments: null pointer");
  SCP_CHECK_ARG(world >= 1 && world <= 64 && rank >= 0 && rank < world, "p2p_allgather_segments: bad rank / world");
  SCP_CHECK_ARG(n_seg >= 1 && n_seg <= 4, "p2p_allgather_segments: 1..4 segments, got %d", n_seg);
  SCP_CHECK_ARG(nbytes > 0 && nbytes % 16 == 0 && nbytes <= nbytes_capacity,
                "p2p_allgather_segments: payload %zu B (multiple of 16, <= %zu)", nbytes, nbytes_capacity);
  SCP_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "p2p_allgather_segments: src / out must be 16-byte aligned");
  P2pSegs segs{};
  segs.n = n_seg;
  unsigned long long off = 0;
  for (int i = 0; i < n_seg; ++i) {
    SCP_CHECK_ARG(seg_bytes[i] > 0 && seg_bytes[i] % 16 == 0, "p2p_allgather_segments: segment %d has %lld B (multiple of 16)", i,
                  (long long)seg_bytes[i]);
    segs.off[i] = off;
    off += (unsigned long long)seg_bytes[i];
    segs.end[i] = off;
  }
  SCP_CHECK_ARG(off == nbytes, "p2p_allgather_segments: segments add up to %llu B, payload is %zu B", off, nbytes);
  const size_t slot = (nbytes_capacity + 255) & ~size_t(255);
  const size_t flag_offset = 2 * (size_t)world * slot + 256;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>(32, (nbytes / 16 + 255) / 256));
  p2p_allgather_kernel<<<blocks, 256, 0, s>>>(static_cast<const uint8_t*>(src), nbytes,
                                              reinterpret_cast<uint8_t* const*>(peer_bufs_device), rank, world, slot,
                                              flag_offset, state);
  SCP_CUDA_LAUNCH_CHECK("p2p_allgather");
  dim3 grid((unsigned)std::max<size_t>(1, std::min<size_t>(16, (nbytes / 16 + 255) / 256)), (unsigned)std::min(world, 8));
  p2p_collect_seg_kernel<<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(local_buf), nbytes, world, slot, state,
                                              static_cast<uint8_t*>(out), segs);
  SCP_CUDA_LAUNCH_CHECK("p2p_collect_seg");
  return SCP_OK;
}
