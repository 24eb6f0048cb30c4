// S2 backward as ONE producer/consumer pipeline over the vocabulary (included by scp_vq.cu).
//
// The backward of the keyword quantiser (avssl/model/kw_branches.py:181-197 + my_vector_quantizer.py:130-136, closed form
// in SURVEY.md section 8(a) V4) needs, per keyword row m,
//     U_m = sum_v Q~[m,v] ehat_v ,  W_m = sum_v P~[m,v] ehat_v ,  sum_v Q~ , sum_v P~
// with P~ = 2^14 softmax_tau(c)[m,v], Q~ = P~ (T'[m,v] - s0_m), c = khat_m . ehat_v, T' = (ghat_m . ehat_v) |e_v| / norm_ref.
// The (M,V) matrices P~, Q~ never reach HBM here.  The grid is split into PIPELINES of two CTA pairs:
//
//   producer pair (tcgen05 cta_group::2)   A = Ehat tile, 256 vocabulary rows (128 per CTA) = TMEM lanes, streamed by TMA
//                                          B = [khat | ghat] of one 128-row keyword tile, RESIDENT in shared memory
//                                              (CTA 0 holds khat, CTA 1 holds ghat: the two halves of N = 256)
//                                          one N = 256 MMA per K step yields c (columns 0..127) and T (128..255) together;
//                                          the epilogue (thread <-> vocabulary row) forms P~, Q~, writes them TRANSPOSED
//                                          ([v][m], 64 contiguous bytes per thread and chunk) into a ring slot in global
//                                          memory and reduces the per-keyword sums across the warp
//   consumer pair (tcgen05 cta_group::2)   M = 256 = [Q~^T rows of the keyword tile (CTA 0) ; P~^T rows (CTA 1)],
//                                          both operands MN-major straight from the ring slot / the unit table (no
//                                          transposed table copy), N = D: CTA 0 accumulates U, CTA 1 accumulates W in TMEM
//                                          over every step of the keyword tile, then dumps a split-K partial
//
// A pipeline owns a contiguous run of (keyword tile, vocabulary tile) steps; the ring (a few slots of 128 KB per pipeline,
// ~19 MB in total) stays in L2, so the only DRAM traffic of the whole backward is the unit table and the small operands.
// Hand-off: `ready[q][slot]` (+1 per producer epilogue warp each time the slot is written, release: the k-th use of a slot is
// complete at 16 k -- a per-pipeline step counter would not do, the epilogue warps drift up to two steps apart) /
// `done[q]` (+1 per step once the consumer's TMA loads of the slot have landed, release); waits are acquire loads, bounded
// (trap after 4 s).  The two roles execute the
// same number of MMA cycles per step (4 M V D FLOP each in total), so the split is 1:1.
//
// All CTAs are co-resident by construction (one CTA per SM, grid <= number of SMs); if some SMs are busy with another
// kernel the late pairs simply start later: a producer blocks on a full ring, a consumer on an empty one, never both.
#pragma once

namespace scp {
namespace pipe {

constexpr int kStepV = 256;                     // vocabulary rows per step (one pair-wide MMA tile)
constexpr int kSlotHalfs = 2 * kStepV * 128;    // [matrix: Q~^T, P~^T][v][m] fp16 = 128 KB
constexpr int kMaxStages = 8;
constexpr int kBarBytes = 1024;
constexpr int kWarpsPerStep = 2 * tc::kEpiWarps;  // producer epilogue warps of a pair

struct PipeMaps {
  CUtensorMap tab_k;   // Ehat (Vp, D), box {64, 128}: producer A operand (K-major)
  CUtensorMap kw;      // khat (Mp, D), box {64, 128}: resident B operand, CTA 0
  CUtensorMap gh;      // ghat (Mp, D), box {64, 128}: resident B operand, CTA 1
  CUtensorMap scr;     // ring / scratch (slots * 512, 128) as (64, row, 2 blocks), box {64, 64, 2}: consumer A operand
  CUtensorMap tab_mn;  // Ehat (Vp, D) as (64, row, D/64 blocks), box {64, 64, nb}: consumer B operand (both MN-major)
};

struct PipeParams {
  int MT, NVT, NP, KC, D, V;
  int64_t M, Mp;
  int fused;      // 1: both roles in this launch (ring + flags); 0: one role per launch, every step has its own slot
  int role;       // fused == 0: 0 = producers, 1 = consumers
  int ring;       // slots per pipeline (fused)
  int sa, sc;     // smem ring stages of the producer / consumer
  int uw_slots;   // max number of pipelines that share one keyword tile
  int debug;      // timing ablations (env SCP_PIPE_DEBUG, results invalid): 1 no epilogue maths, 2 no MMAs, 4 no P~/Q~ stores
  __half* scratch;
  unsigned int* ready;      // [NP][ring]: per ring slot, +1 per producer epilogue warp each time the slot has been written
  unsigned int* done;       // [NP]: steps whose slot the consumer has loaded
  const float* row_stats;   // (M,4): [1] = lse at tau
  const float* g_aux;       // (Mp,2): |g|, s0
  const float* table_norm;  // (Vp,)
  const float* table_mean;  // [D] = norm_ref
  const float* tau;
  float* sums;              // (Mp, uw_slots*8, 4): sum Q~, sum P~, sum Q~ c, sum P~ c
  float* uw;                // (uw_slots, 2, Mp, D): U, W partials
  MaskedCols mc;
};

// pipeline q owns the steps [total*q/NP, total*(q+1)/NP) of the list (keyword tile major, vocabulary tile minor)
__host__ __device__ inline int pipe_of_step(long long s, long long total, int NP) {
  return (int)(((s + 1) * NP - 1) / total);
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded spin: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void spin_until_ge(const unsigned int* p, unsigned int target) {
  if (ld_acquire_gpu(p) >= target) return;
  const uint64_t t0 = tc::global_timer_ns();
  while (ld_acquire_gpu(p) < target) {
    __nanosleep(64);
    if (tc::global_timer_ns() - t0 > 4000000000ull) {
      printf("scp: vq_bwd pipeline flag wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// MN-major SWIZZLE_128B operand: rows of 128 B = 64 consecutive M/N elements of ONE k index, 8-row groups 1024 B apart
// (SBO), 64-element M/N blocks `lbo_bytes` apart (LBO).  This is what TMA boxes {64, rows} with SWIZZLE_128B produce when
// the global matrix is stored with the M/N extent contiguous.
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// pair MMA, fp16 x fp16 -> fp32, BOTH operands MN-major (a_major bit 15, b_major bit 16)
__host__ __device__ constexpr uint32_t make_idesc_f16_pair_mn(int n) {
  return tc::make_idesc_f16_pair(n) | (1u << 15) | (1u << 16);
}

// ---- per-keyword sums over the vocabulary ------------------------------------------------------------------------------
// The producer epilogue holds lane <-> vocabulary row, register <-> keyword column, so sum_v P~[m,v] is a CROSS-LANE sum.
// Measured alternatives (B200, M = 2048, V = 49408, D = 512, producer alone on 148 SMs):
//   * shuffle transpose-reduce per step (31 x {2 FSEL, SHFL, FADD} per quantity and 32 columns): FSEL + SHFL + FADD were
//     49 % of the stall samples, 222 us;
//   * staging the fp16 tile in shared memory and summing it with ones x tile on mma.sync (ldmatrix + HMMA.16816): 291 us --
//     the legacy HMMA path shares the tensor pipe with the tcgen05 MMAs that keep it busy, so the epilogue stalls on it;
//   * (this version) thread-private fp32 running sums, one packed add per column pair and quantity, reduced across lanes
//     ONCE per keyword tile.  The 128 accumulator registers per thread come from setmaxnreg: the two single-lane warps
//     (and two idle warps that complete their warpgroup) shrink to 40 registers, the eight epilogue warps grow to 232.
// Lane j of the warp receives the sum over the 32 lanes of v[j] (31 shuffles instead of 32 x 5).
template <int O>
__device__ __forceinline__ void tr_step(float (&v)[32], int lane) {
  const bool up = (lane & O) != 0;
#pragma unroll
  for (int i = 0; i < O; ++i) {
    const float a = v[i], b = v[i + O];
    const float send = up ? a : b;
    const float keep = up ? b : a;
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
  }
}
__device__ __forceinline__ float lane_column_sum(float (&v)[32], int lane) {
  tr_step<16>(v, lane);
  tr_step<8>(v, lane);
  tr_step<4>(v, lane);
  tr_step<2>(v, lane);
  tr_step<1>(v, lane);
  return v[0];
}
__device__ __forceinline__ float lane_column_sum2(const tc::f32x2 (&a)[16], int lane) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) tc::unpack2(a[i], v[2 * i], v[2 * i + 1]);
  return lane_column_sum(v, lane);
}
constexpr int kPipeThreads = 384;      // 8 epilogue warps + TMA warp + MMA warp + 2 idle warps (whole warpgroups)
constexpr int kRegsEpilogue = 232;     // 8 * 32 * 232 + 4 * 32 * 40 = 384 * 168: the launch-time allocation, redistributed
constexpr int kRegsOther = 40;
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <bool WANT_TAU>
__global__ void __launch_bounds__(kPipeThreads, 1)
vq_bwd_pipe_kernel(const __grid_constant__ PipeMaps maps, const PipeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bfree_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfree_bar + 1);
  uint8_t* data = base + kBarBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)tc::cluster_ctarank();
  const int pair = (int)blockIdx.x >> 1;
  const int role = p.fused ? (pair & 1) : p.role;
  const int q = p.fused ? (pair >> 1) : pair;
  const long long total = (long long)p.MT * p.NVT;
  const int t0 = (int)(total * q / p.NP), t1 = (int)(total * (q + 1) / p.NP);
  constexpr uint16_t kPairMask = 3;

  if (warp == tc::kProducerWarp && lane == 0) {
    tc::prefetch_tmap(role == 0 ? &maps.tab_k : &maps.scr);
    tc::prefetch_tmap(role == 0 ? (rank == 0 ? &maps.kw : &maps.gh) : &maps.tab_mn);
  }
  if (warp == tc::kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kMaxStages; ++s) {
        tc::mbar_init(&full_bar[s], 1);
        tc::mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        tc::mbar_init(&tfull_bar[a], 1);
        tc::mbar_init(&tempty_bar[a], kWarpsPerStep);  // one arrive per epilogue warp of both CTAs
      }
      tc::mbar_init(bfree_bar, 1);
      tc::fence_barrier_init();
    }
    __syncwarp();
    tc::tmem_alloc_pair(tmem_slot, tc::kTmemCols);
  }
  tc::tc_fence_before();
  __syncwarp();
  tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (role == 0) {
    // =============================================================================================================
    // producer: logits -> P~^T, Q~^T
    // =============================================================================================================
    uint8_t* res_b = data;                                   // KC chunks x 16 KB: this CTA's half of [khat | ghat]
    uint8_t* ring_a = data + (size_t)p.KC * tc::kXTileBytes;  // sa stages x 16 KB of Ehat rows
    float* vec_bias = reinterpret_cast<float*>(ring_a + (size_t)p.sa * tc::kXTileBytes);
    float* vec_ns0 = vec_bias + 128;
    if (warp >= tc::kEpiWarps) {
    setmaxnreg_dec<kRegsOther>();  // warpgroup 2 (TMA warp, MMA warp, two idle warps) hands its registers to the epilogue
    if (warp == tc::kProducerWarp) {
      if (lane == 0) {
        int stage = 0, prev_mt = -1;
        uint32_t phase = 0, bphase = 0;
        for (int t = t0; t < t1; ++t) {
          const int mt = t / p.NVT, vt = t - mt * p.NVT;
          const bool newseg = mt != prev_mt;
          if (newseg && prev_mt >= 0) {  // every MMA that reads the old keyword tile has retired
            tc::mbar_wait(bfree_bar, bphase);
            bphase ^= 1;
          }
          prev_mt = mt;
          for (int kc = 0; kc < p.KC; ++kc) {
            tc::mbar_wait(&empty_bar[stage], phase ^ 1);
            const uint32_t tx = (uint32_t)tc::kXTileBytes * (newseg ? 2u : 1u);
            if (rank == 0) tc::mbar_arrive_expect_tx(&full_bar[stage], 2 * tx);
            const uint32_t lead_bar = tc::mapa_u32(&full_bar[stage], 0);
            tc::tma_load_2d_pair(ring_a + (size_t)stage * tc::kXTileBytes, &maps.tab_k, kc * tc::kChunkK,
                                 vt * kStepV + rank * 128, lead_bar);
            if (newseg)
              tc::tma_load_2d_pair(res_b + (size_t)kc * tc::kXTileBytes, rank == 0 ? &maps.kw : &maps.gh,
                                   kc * tc::kChunkK, mt * 128, lead_bar);
            if (++stage == p.sa) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == tc::kMmaWarp) {
      if (lane == 0 && rank == 0) {
        constexpr uint32_t idesc = tc::make_idesc_f16_pair(256);
        int stage = 0, as = 0;
        uint32_t phase = 0, aphase = 0;
        for (int t = t0; t < t1; ++t) {
          const int mt = t / p.NVT;
          tc::mbar_wait(&tempty_bar[as], aphase ^ 1);
          tc::tc_fence_after();
          for (int kc = 0; kc < p.KC; ++kc) {
            tc::mbar_wait(&full_bar[stage], phase);
            tc::tc_fence_after();
            const uint64_t a_desc = tc::make_kmajor_sw128_desc(tc::smem_u32(ring_a + (size_t)stage * tc::kXTileBytes));
            const uint64_t b_desc = tc::make_kmajor_sw128_desc(tc::smem_u32(res_b + (size_t)kc * tc::kXTileBytes));
            const uint32_t d = tmem_base + (uint32_t)(as * 256);
#pragma unroll
            for (int k = 0; k < tc::kChunkK / tc::kUmmaK; ++k) {
              if (p.debug & 2) continue;
              tc::umma_f16_pair(d, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kc > 0 || k > 0) ? 1u : 0u);
            }
            tc::umma_commit_pair_mc(&empty_bar[stage], kPairMask);
            if (++stage == p.sa) { stage = 0; phase ^= 1; }
          }
          tc::umma_commit_pair_mc(&tfull_bar[as], kPairMask);
          if (t + 1 < t1 && (t + 1) / p.NVT != mt) tc::umma_commit_pair_mc(bfree_bar, kPairMask);
          if (++as == 2) { as = 0; aphase ^= 1; }
        }
      }
    }
    } else {
      setmaxnreg_inc<kRegsEpilogue>();
      const int quad = warp & 3, half = warp >> 2;
      const int vrow = rank * 128 + quad * 32 + lane;  // row inside the step's 256 vocabulary rows
      const float k_tau = kLog2e / __ldg(p.tau);
      const float inv_norm_ref = 1.0f / __ldg(p.table_mean + p.D);
      const tc::f32x2 kt2 = tc::pack2(k_tau, k_tau);
      const int sum_stride = p.uw_slots * 8;
      // running sums of this thread's vocabulary rows: [chunk group][column pair] (WANT_TAU: plain per-step reduction)
      tc::f32x2 acc_q[WANT_TAU ? 1 : 2][16], acc_p[WANT_TAU ? 1 : 2][16];
      float sum_q[2] = {0.f, 0.f}, sum_p[2] = {0.f, 0.f}, sum_qc[2] = {0.f, 0.f}, sum_pc[2] = {0.f, 0.f};
      if constexpr (!WANT_TAU) {
#pragma unroll
        for (int gi = 0; gi < 2; ++gi)
#pragma unroll
          for (int i = 0; i < 16; ++i) acc_q[gi][i] = acc_p[gi][i] = tc::pack2(0.f, 0.f);
      }
      int as = 0, prev_mt = -1;
      uint32_t aphase = 0;
      float rv_next = t0 < t1 ? __ldg(p.table_norm + (t0 % p.NVT) * kStepV + vrow) : 0.f;
      auto flush_sums = [&](int mt) {
        const int slot_s = (q - pipe_of_step((long long)mt * p.NVT, total, p.NP)) * 8 + rank * 4 + quad;
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
          if constexpr (!WANT_TAU) {  // the one cross-lane reduction of the keyword tile
            sum_q[gi] = lane_column_sum2(acc_q[gi], lane);
            sum_p[gi] = lane_column_sum2(acc_p[gi], lane);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc_q[gi][i] = acc_p[gi][i] = tc::pack2(0.f, 0.f);
          }
          const int64_t m = (int64_t)mt * 128 + (half * 2 + gi) * 32 + lane;
          *reinterpret_cast<float4*>(p.sums + (m * sum_stride + slot_s) * 4) =
              make_float4(sum_q[gi], sum_p[gi], sum_qc[gi], sum_pc[gi]);
          sum_q[gi] = sum_p[gi] = sum_qc[gi] = sum_pc[gi] = 0.f;
        }
      };
      for (int t = t0; t < t1; ++t) {
        const int mt = t / p.NVT, vt = t - mt * p.NVT;
        if (mt != prev_mt) {
          if (prev_mt >= 0) flush_sums(prev_mt);
          tc::named_bar_sync(tc::kEpiBarrierId, tc::kEpiThreads);  // the old per-keyword vectors are no longer read
          if (threadIdx.x < 128) {
            const int64_t m = (int64_t)mt * 128 + threadIdx.x;
            const bool mv = m < p.M;
            // P~ = 2^(c * k_tau + bias); 2^14 = kPScale keeps the softmax row in the fp16 normal range; padding rows: 0
            vec_bias[threadIdx.x] = mv ? 14.0f - __ldg(p.row_stats + m * 4 + 1) * kLog2e : -1.0e30f;
            vec_ns0[threadIdx.x] = -__ldg(p.g_aux + m * 2 + 1);
          }
          tc::named_bar_sync(tc::kEpiBarrierId, tc::kEpiThreads);
          prev_mt = mt;
        }
        const int v = vt * kStepV + vrow;  // < Vp
        const bool valid = v < p.V && !is_masked(p.mc, v);
        const float rv = rv_next * inv_norm_ref;  // T' = (ghat . ehat_v) * |e_v| / norm_ref - s0
        if (t + 1 < t1) rv_next = __ldg(p.table_norm + ((t + 1) % p.NVT) * kStepV + vrow);  // fetched one step ahead
        const tc::f32x2 rv2 = tc::pack2(rv, rv);
        const int li = t - t0;
        if (p.fused && li >= p.ring) {  // the slot's previous occupant has been loaded by the consumer
          if (lane == 0) spin_until_ge(p.done + q, (unsigned)(li - p.ring + 1));
          __syncwarp();
        }
        __half* slot = p.scratch + (size_t)(p.fused ? (long long)q * p.ring + li % p.ring : (long long)t) * kSlotHalfs;
        tc::mbar_wait(&tfull_bar[as], aphase);
        tc::tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256);
#pragma unroll  // (fully unrolled: the running sums are indexed by gi and must stay in registers)
        for (int gi = 0; gi < 2; ++gi) {
          const int g = half * 2 + gi;
          __half* dq = slot + (size_t)vrow * 128 + g * 32;
          const uint32_t bsm = tc::smem_u32(vec_bias) + (uint32_t)(g * 128), ssm = tc::smem_u32(vec_ns0) + (uint32_t)(g * 128);
          float tp[WANT_TAU ? 32 : 1], tqv[WANT_TAU ? 32 : 1], tpc[WANT_TAU ? 32 : 1], tqc[WANT_TAU ? 32 : 1];
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {  // 16 columns of c and of T at a time
            uint32_t rc[16], rt[16];
            __syncwarp();
            if (!(p.debug & 8)) {
              tc::tmem_ld16_issue(tbase + (uint32_t)(g * 32 + hb * 16), rc);
              tc::tmem_ld16_issue(tbase + (uint32_t)(128 + g * 32 + hb * 16), rt);
              tc::tmem_ld16_wait(rc);
              tc::tmem_ld16_wait(rt);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) { rc[i] = (uint32_t)(i + lane); rt[i] = (uint32_t)(i * lane); }
            }
            if (gi == 1 && hb == 1) {  // the accumulator set is in registers: hand it back to the MMA issuer before the maths
              tc::tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (rank != 0) tc::mbar_arrive_cluster(tc::mapa_u32(&tempty_bar[as], 0));
                else tc::mbar_arrive(&tempty_bar[as]);
              }
            }
            uint32_t hq[8], hp[8];
            if (p.debug & 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { hp[i] = rc[i]; hq[i] = rt[i]; }
            } else if (valid) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {  // 4 columns per iteration
                const float4 b4 = tc::lds128f(bsm + (4 * hb + i) * 16), s4 = tc::lds128f(ssm + (4 * hb + i) * 16);
                const tc::f32x2 bb[2] = {tc::pack2(b4.x, b4.y), tc::pack2(b4.z, b4.w)};
                const tc::f32x2 ss[2] = {tc::pack2(s4.x, s4.y), tc::pack2(s4.z, s4.w)};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const int e = 4 * i + 2 * j;       // column inside this half
                  const int o = 8 * hb + (e >> 1);   // packed pair inside the chunk group
                  const tc::f32x2 cc = tc::pack2(__uint_as_float(rc[e]), __uint_as_float(rc[e + 1]));
                  const tc::f32x2 pj = tc::ex2_2(tc::fma2(cc, kt2, bb[j]));
                  const tc::f32x2 tj = tc::fma2(tc::pack2(__uint_as_float(rt[e]), __uint_as_float(rt[e + 1])), rv2, ss[j]);
                  const tc::f32x2 qj = tc::mul2(pj, tj);
                  hp[e >> 1] = tc::cvt_f16x2(pj);
                  hq[e >> 1] = tc::cvt_f16x2(qj);
                  if constexpr (WANT_TAU) {  // also the c-weighted sums of the learnable-temperature gradient
                    tc::unpack2(pj, tp[2 * o], tp[2 * o + 1]);
                    tc::unpack2(qj, tqv[2 * o], tqv[2 * o + 1]);
                    tc::unpack2(tc::mul2(pj, cc), tpc[2 * o], tpc[2 * o + 1]);
                    tc::unpack2(tc::mul2(qj, cc), tqc[2 * o], tqc[2 * o + 1]);
                  } else {
                    acc_p[gi][o] = tc::add2(acc_p[gi][o], pj);
                    acc_q[gi][o] = tc::add2(acc_q[gi][o], qj);
                  }
                }
              }
            } else {  // masked / padding vocabulary row: contributes nothing
#pragma unroll
              for (int i = 0; i < 8; ++i) { hp[i] = 0u; hq[i] = 0u; }
              if constexpr (WANT_TAU) {
#pragma unroll
                for (int i = 16 * hb; i < 16 * hb + 16; ++i) tp[i] = tqv[i] = tpc[i] = tqc[i] = 0.f;
              }
            }
            if (!(p.debug & 4)) {
              tc::stg256(dq + 16 * hb, hq);
              tc::stg256(dq + kSlotHalfs / 2 + 16 * hb, hp);
            } else if (hq[0] == 0x12345678u && hp[3] == 0x9abcdef0u) {  // keep the values alive without storing them
              tc::stg256(dq, hq);
            }
          }
          if constexpr (WANT_TAU) {  // rare path (no shipped recipe learns the VQ temperature): reduce every step
            sum_p[gi] += lane_column_sum(tp, lane);
            sum_q[gi] += lane_column_sum(tqv, lane);
            sum_pc[gi] += lane_column_sum(tpc, lane);
            sum_qc[gi] += lane_column_sum(tqc, lane);
          }
        }
        if (p.fused) {  // publish this warp's part of the slot
          // bar.warp.sync orders the 32 lanes' stores before lane 0's release (cumulativity): no blocking fence needed
          __syncwarp();
          if (lane == 0) red_release_gpu_add(p.ready + (size_t)q * p.ring + li % p.ring, 1u);
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      if (prev_mt >= 0) flush_sums(prev_mt);
    }
  } else {
    // =============================================================================================================
    // consumer: U += Q~^T-tile x Ehat-tile, W += P~^T-tile x Ehat-tile
    // =============================================================================================================
    const int n_mma = p.D > 256 ? 2 : 1;
    const int n_each = p.D / n_mma;     // N of one MMA
    const int nh = n_each / 2;          // table columns of one MMA staged by this CTA
    const int nb = nh / 64;             // 64-column blocks of 8 KB
    const uint32_t stage_bytes = (uint32_t)tc::kXTileBytes + (uint32_t)p.D * 64u;
    uint8_t* ring = data;
    if (warp >= tc::kEpiWarps) {
    setmaxnreg_dec<kRegsOther>();
    if (warp == tc::kProducerWarp) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t0; t < t1; ++t) {
          const int mt = t / p.NVT, vt = t - mt * p.NVT;
          const int li = t - t0;
          (void)mt;
          if (p.fused) {
            spin_until_ge(p.ready + (size_t)q * p.ring + li % p.ring, (unsigned)(kWarpsPerStep * (li / p.ring + 1)));
            fence_proxy_async_all();  // the slot was written through the generic proxy, TMA reads through the async proxy
          }
          const long long slot = p.fused ? (long long)q * p.ring + li % p.ring : (long long)t;
          for (int kc = 0; kc < kStepV / 64; ++kc) {
            tc::mbar_wait(&empty_bar[stage], phase ^ 1);
            if (rank == 0) tc::mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_bytes);
            const uint32_t lead_bar = tc::mapa_u32(&full_bar[stage], 0);
            uint8_t* st = ring + (size_t)stage * stage_bytes;
            const int srow = (int)(slot * 512 + rank * 256 + kc * 64);
            tc::tma_load_3d_pair(st, &maps.scr, 0, srow, 0, lead_bar);  // both 64-keyword blocks: [block][v][64]
            const int vrow0 = vt * kStepV + kc * 64;
            for (int h = 0; h < n_mma; ++h)  // this CTA's nb 64-column blocks of MMA h
              tc::tma_load_3d_pair(st + tc::kXTileBytes + (size_t)(h * nb) * 8192, &maps.tab_mn, 0, vrow0,
                                   (h * n_each + rank * nh) / 64, lead_bar);
            if (++stage == p.sc) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == tc::kMmaWarp) {
      if (lane == 0 && rank == 0) {
        const uint32_t idesc = make_idesc_f16_pair_mn(n_each);
        int stage = 0, prev_mt = -1;
        uint32_t phase = 0, sphase = 0;
        for (int t = t0; t < t1; ++t) {
          const int mt = t / p.NVT;
          const bool seg_first = mt != prev_mt;
          prev_mt = mt;
          if (seg_first) {  // the previous keyword tile's accumulators have been read out
            tc::mbar_wait(&tempty_bar[0], sphase ^ 1);
            tc::tc_fence_after();
          }
          for (int kc = 0; kc < kStepV / 64; ++kc) {
            tc::mbar_wait(&full_bar[stage], phase);
            tc::tc_fence_after();
            if (p.fused && kc == kStepV / 64 - 1) red_release_gpu_add(p.done + q, 1u);  // the slot is in shared memory
            const uint32_t sa = tc::smem_u32(ring + (size_t)stage * stage_bytes);
            const uint32_t sb = sa + (uint32_t)tc::kXTileBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t acc = (seg_first && kc == 0 && k == 0) ? 0u : 1u;
              const uint64_t a_desc = make_mn_sw128_desc(sa + (uint32_t)k * 2048u, 8192u);
              for (int h = 0; h < n_mma; ++h) {
                if (p.debug & 2) continue;
                const uint64_t b_desc = make_mn_sw128_desc(sb + (uint32_t)(h * nb) * 8192u + (uint32_t)k * 2048u, 8192u);
                tc::umma_f16_pair(tmem_base + (uint32_t)(h * n_each), a_desc, b_desc, idesc, acc);
              }
            }
            tc::umma_commit_pair_mc(&empty_bar[stage], kPairMask);
            if (++stage == p.sc) { stage = 0; phase ^= 1; }
          }
          if (t + 1 == t1 || (t + 1) / p.NVT != mt) {
            tc::umma_commit_pair_mc(&tfull_bar[0], kPairMask);
            sphase ^= 1;
          }
        }
      }
    }
    } else {
      setmaxnreg_inc<kRegsEpilogue>();
      const int quad = warp & 3, half = warp >> 2;
      if (t1 > t0) {
        const int mt_a = t0 / p.NVT, mt_b = (t1 - 1) / p.NVT;
        uint32_t sphase = 0;
        const int n_chunks = p.D / 32, c_lo = half * (n_chunks / 2), c_hi = (half + 1) * (n_chunks / 2);
        for (int mt = mt_a; mt <= mt_b; ++mt) {
          const int slot_s = q - pipe_of_step((long long)mt * p.NVT, total, p.NP);
          tc::mbar_wait(&tfull_bar[0], sphase);
          tc::tc_fence_after();
          const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16);
          const int64_t m = (int64_t)mt * 128 + quad * 32 + lane;
          float* dst = p.uw + (((int64_t)slot_s * 2 + rank) * p.Mp + m) * p.D;  // rank 0: U, rank 1: W
#pragma unroll 1
          for (int cc = c_lo; cc < c_hi; ++cc) {
            float v[32];
            __syncwarp();
            tc::tmem_ld32(tbase + (uint32_t)(cc * 32), v);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(dst + cc * 32 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank != 0) tc::mbar_arrive_cluster(tc::mapa_u32(&tempty_bar[0], 0));
            else tc::mbar_arrive(&tempty_bar[0]);
          }
          sphase ^= 1;
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncwarp();
  tc::cluster_sync_all();
  if (warp == tc::kMmaWarp) tc::tmem_dealloc_pair(tmem_base, tc::kTmemCols);
}

// block (D/4 threads) <-> keyword row: combine the pipelines' partials of this row's keyword tile in a fixed order,
//   g_khat = (U - s W) * scale / tau ;  g_kw = (g_khat - <g_khat,khat> khat) / ||kw||     (see vq_bwd_finalize_kernel)
__global__ void __launch_bounds__(256)
vq_bwd_pipe_finalize_kernel(const float* __restrict__ uw, int uw_slots, int NVT, int MT, int NP, int64_t M, int64_t Mp,
                            int D, const float* __restrict__ sums, const float* __restrict__ g_aux,
                            const float* __restrict__ kw, const float* __restrict__ row_stats,
                            const float* __restrict__ table_mean, const float* __restrict__ tau_ptr,
                            float* __restrict__ g_kw, float* __restrict__ g_tau) {
  __shared__ float s_red[8];
  const int64_t m = blockIdx.x;
  const int mt = (int)(m >> 7);
  const long long total = (long long)MT * NVT;
  const int qa = pipe_of_step((long long)mt * NVT, total, NP), qb = pipe_of_step((long long)(mt + 1) * NVT - 1, total, NP);
  const int ns = qb - qa + 1;  // pipelines that worked on this keyword tile
  const int d0 = threadIdx.x * 4;
  const float tau = *tau_ptr;
  float sq = 0.f, sp = 0.f, sqc = 0.f, spc = 0.f;
  const float4* srow = reinterpret_cast<const float4*>(sums) + m * (int64_t)(uw_slots * 8);
  for (int j = 0; j < ns * 8; ++j) {  // same address in every thread: broadcast loads
    const float4 t = srow[j];
    sq += t.x; sp += t.y; sqc += t.z; spc += t.w;
  }
  const float s_adj = sp > 0.f ? sq / sp : 0.f;
  const float scale = g_aux[m * 2] * table_mean[D] / (kPScale * tau);
  const float inv_norm = row_stats[m * 4 + 3];
  float4 u = make_float4(0.f, 0.f, 0.f, 0.f), w = u;
  for (int ks = 0; ks < ns; ++ks) {
    const float4 a = *reinterpret_cast<const float4*>(uw + (((int64_t)ks * 2 + 0) * Mp + m) * D + d0);
    const float4 b = *reinterpret_cast<const float4*>(uw + (((int64_t)ks * 2 + 1) * Mp + m) * D + d0);
    u.x += a.x; u.y += a.y; u.z += a.z; u.w += a.w;
    w.x += b.x; w.y += b.y; w.z += b.z; w.w += b.w;
  }
  const float4 k4 = *reinterpret_cast<const float4*>(kw + m * D + d0);
  const float gk[4] = {(u.x - s_adj * w.x) * scale, (u.y - s_adj * w.y) * scale, (u.z - s_adj * w.z) * scale,
                       (u.w - s_adj * w.w) * scale};
  const float kh[4] = {k4.x * inv_norm, k4.y * inv_norm, k4.z * inv_norm, k4.w * inv_norm};
  float proj = gk[0] * kh[0] + gk[1] * kh[1] + gk[2] * kh[2] + gk[3] * kh[3];
  proj = warp_sum(proj);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = proj;
  __syncthreads();
  proj = 0.f;
  for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) proj += s_red[i];  // fixed order
  *reinterpret_cast<float4*>(g_kw + m * D + d0) =
      make_float4((gk[0] - proj * kh[0]) * inv_norm, (gk[1] - proj * kh[1]) * inv_norm,
                  (gk[2] - proj * kh[2]) * inv_norm, (gk[3] - proj * kh[3]) * inv_norm);
  if (g_tau && threadIdx.x == 0) {
    const float contrib = -(g_aux[m * 2] * table_mean[D]) * (sqc - s_adj * spc) / (kPScale * tau * tau);
    atomicAdd(g_tau, contrib);
  }
}

}  // namespace pipe
}  // namespace scp
