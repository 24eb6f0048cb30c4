// Row-streaming tcgen05 GEMM engine shared by the VQ and InfoNCE kernels.
//
// One CTA owns a 128-row tile of X (rows = TMEM lanes) and walks a list of BN-row tiles of Y, computing
//     S_x = X_x[tile] * Y[tile]^T        (fp16 operands, fp32 accumulation in TMEM, x < NX accumulators sharing Y)
// over a K range, and hands every finished accumulator to an epilogue functor that consumes it ROW-WISE
// (thread <-> row, 32 columns at a time), so that row reductions (max / arg-max / log-sum-exp / dot products) are
// thread-serial and the logits never leave the SM.
//
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..9 = epilogue: warp w reads TMEM lanes 32*(w%4) .. +32; the two warps that share a lane quadrant split the
// tile's 32-column chunks between them ("halves"), so every SM sub-partition has two epilogue warps to hide latency.
// Pipelines: STAGES-deep smem ring (full/empty mbarriers, TMA <-> MMA) and ACC_STAGES TMEM accumulator sets
// (tmem_full/tmem_empty mbarriers, MMA <-> epilogue) so that the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "scp_tc.cuh"

namespace scp {
namespace tc {

struct GemmMaps {
  CUtensorMap x[2];
  CUtensorMap y;
  CUtensorMap o;  // optional output map for epilogues that store tiles with TMA
};

constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kEpiBarrierId = 1;

struct EpiCtx {
  int row_in_tile;  // 0..127: TMEM lane == X row inside the CTA tile
  int half;         // 0/1: which half of every tile's column chunks this warp consumes
  int tid;          // 0..255 inside the epilogue group
  uint8_t* smem;    // Epi::kSmemBytes bytes, 1024-byte aligned, shared by the epilogue group
  const GemmMaps* maps;
};

// CTA -> work decomposition.  blockIdx.x = m_tile + m_tiles * (n_group + n_groups * k_split)
struct Sched {
  int m_tiles;   // 128-row tiles of X
  int n_tiles;   // BN-row tiles of Y
  int n_groups;  // the n_tiles are partitioned into n_groups contiguous ranges
  int k_chunks;  // total 64-element K chunks
  int k_splits;  // the k_chunks are partitioned into k_splits contiguous ranges
  // two-direction launches (InfoNCE): CTAs whose m_tile >= m_half walk Y tiles shifted by n_upper_off
  int m_half;           // == m_tiles when unused
  int n_upper_off;      // == 0 when unused
  int x_upper_row_off;  // extra X row offset of the upper-half tiles (== 0 when unused)
  __host__ __device__ int grid() const { return m_tiles * n_groups * k_splits; }
};

struct WorkInfo {
  int m_tile, n_group, k_split;
  int x_row;     // first X row of this CTA's tile (TMA coordinate)
  int nt0, nt1;  // N-tile range
  int kc0, kc1;  // K-chunk range
};

__device__ __forceinline__ WorkInfo decode_work(const Sched& s) {
  WorkInfo w;
  int b = blockIdx.x;
  w.m_tile = b % s.m_tiles;
  b /= s.m_tiles;
  w.n_group = b % s.n_groups;
  w.k_split = b / s.n_groups;
  w.nt0 = (int)((long long)s.n_tiles * w.n_group / s.n_groups);
  w.nt1 = (int)((long long)s.n_tiles * (w.n_group + 1) / s.n_groups);
  w.x_row = w.m_tile * kTileM;
  if (w.m_tile >= s.m_half) {
    w.nt0 += s.n_upper_off;
    w.nt1 += s.n_upper_off;
    w.x_row += s.x_upper_row_off;
  }
  w.kc0 = (int)((long long)s.k_chunks * w.k_split / s.k_splits);
  w.kc1 = (int)((long long)s.k_chunks * (w.k_split + 1) / s.k_splits);
  return w;
}

constexpr int kGemmThreads = 64 + kEpiThreads;
constexpr int kXTileBytes = kTileM * kChunkK * 2;  // 16 KB

template <int BN, int NX, int STAGES>
struct GemmCfg {
  static constexpr int kYTileBytes = BN * kChunkK * 2;
  static constexpr int kStageBytes = NX * kXTileBytes + kYTileBytes;
  static constexpr int kAccCols = NX * BN;
  static constexpr int kAccStages = (int)kTmemCols / kAccCols >= 2 ? 2 : 1;
  static constexpr int kBarrierBytes = 1024;  // mbarriers + TMEM slot; keeps the epilogue scratch 1024-aligned
  // 1024 B slack for manual alignment of the dynamic smem base
  static constexpr int smem_bytes(int epi_bytes) { return 1024 + STAGES * kStageBytes + kBarrierBytes + epi_bytes; }
  static_assert(kAccCols <= (int)kTmemCols, "accumulators exceed TMEM");
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN");
  static_assert((2 * STAGES + 4) * 8 + 16 <= kBarrierBytes, "barrier area");
  static_assert(STAGES * kStageBytes <= 220 * 1024, "smem ring too large");
};

// Epi interface (all __device__ __forceinline__):
//   struct Params;                      POD passed by value to the kernel
//   static constexpr int kSmemBytes;    extra shared memory (shared by the 8 epilogue warps)
//   Epi(const Params&, const WorkInfo&, const EpiCtx&)
//   void tile_begin(int n_tile);
//   void chunk(int col0, float (&v)[NX][32]);      // columns [col0, col0+32) of Y-row space (global index)
//   void tile_end(int n_tile);
//   void finish();
// Per-row state lives in the two warps ("halves") that own the row; epilogues combine the halves themselves
// (separate partial slots, or through ctx.smem + named_bar_sync(kEpiBarrierId, kEpiThreads)).
template <int BN, int NX, int STAGES, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
stream_gemm_kernel(const __grid_constant__ GemmMaps maps, const Sched sched, const typename Epi::Params ep) {
  using Cfg = GemmCfg<BN, NX, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_smem = smem + STAGES * Cfg::kStageBytes + Cfg::kBarrierBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WorkInfo work = decode_work(sched);

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int x = 0; x < NX; ++x) prefetch_tmap(&maps.x[x]);
    prefetch_tmap(&maps.y);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], kEpiWarps);  // one arrive per epilogue warp
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int nt = work.nt0; nt < work.nt1; ++nt) {
        for (int kc = work.kc0; kc < work.kc1; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
          for (int x = 0; x < NX; ++x)
            tma_load_2d(st + x * kXTileBytes, &maps.x[x], kc * kChunkK, work.x_row, &full_bar[stage]);
          tma_load_2d(st + NX * kXTileBytes, &maps.y, kc * kChunkK, nt * BN, &full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BN);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int nt = work.nt0; nt < work.nt1; ++nt) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        for (int kc = work.kc0; kc < work.kc1; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t b_desc = make_kmajor_sw128_desc(st + NX * kXTileBytes);
#pragma unroll
          for (int x = 0; x < NX; ++x) {
            const uint64_t a_desc = make_kmajor_sw128_desc(st + x * kXTileBytes);
            const uint32_t d = tmem_base + (uint32_t)(as * Cfg::kAccCols + x * BN);
#pragma unroll
            for (int k = 0; k < kChunkK / kUmmaK; ++k)
              umma_f16(d, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kc > work.kc0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator set complete
        if (++as == Cfg::kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ---------------- epilogue warps ----------------
    const int quad = warp & 3;
    EpiCtx ctx;
    ctx.row_in_tile = quad * 32 + lane;
    ctx.half = (warp - 2) >> 2;
    ctx.tid = threadIdx.x - 64;
    ctx.smem = epi_smem;
    ctx.maps = &maps;
    Epi epi(ep, work, ctx);
    constexpr int kChunksPerHalf = BN / 64;
    int as = 0;
    uint32_t aphase = 0;
    for (int nt = work.nt0; nt < work.nt1; ++nt) {
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      epi.tile_begin(nt);
      const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * Cfg::kAccCols);
#pragma unroll 1
      for (int cc = 0; cc < kChunksPerHalf; ++cc) {
        const int c = ctx.half * kChunksPerHalf + cc;
        float v[NX][32];
        __syncwarp();
#pragma unroll
        for (int x = 0; x < NX; ++x) tmem_ld32(tbase + (uint32_t)(x * BN + c * 32), v[x]);
        epi.chunk(nt * BN + c * 32, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      epi.tile_end(nt);
      if (++as == Cfg::kAccStages) { as = 0; aphase ^= 1; }
    }
    epi.finish();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int BN, int NX, int STAGES, class Epi>
int launch_stream_gemm(const GemmMaps& maps, const Sched& sched, const typename Epi::Params& ep, cudaStream_t stream,
                       const char* name) {
  using Cfg = GemmCfg<BN, NX, STAGES>;
  auto kern = stream_gemm_kernel<BN, NX, STAGES, Epi>;
  const int smem = Cfg::smem_bytes(Epi::kSmemBytes);
  static thread_local bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(SCP_ERR_CUDA, "%s: cudaFuncSetAttribute(%d B): %s", name, smem, cudaGetErrorString(e));
    configured = true;
  }
  if (sched.grid() <= 0) return SCP_OK;
  kern<<<sched.grid(), kGemmThreads, smem, stream>>>(maps, sched, ep);
  SCP_CUDA_LAUNCH_CHECK(name);
  return SCP_OK;
}

}  // namespace tc
}  // namespace scp
