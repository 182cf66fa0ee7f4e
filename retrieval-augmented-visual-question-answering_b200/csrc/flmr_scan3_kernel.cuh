// flmr_scan3_kernel.cuh — the fused late-interaction scan with THREE epilogue warpgroups.
//
// Same contract, data layout, TMA producer, TMEM-resident queries, reducer and top-k as flmr_scan_kernel
// (flmr_scan_kernel.cuh, whose constants and helpers this header reuses).  What differs is who drains which
// accumulator:
//   * 16 warps (512 threads, 128 registers each): warps 0..11 = three epilogue warpgroups, 12 = reducer,
//     13 = TMA producer, 14..15 = tcgen05.mma issuers.
//   * STATIC assignment: query tile mt is always drained by warpgroup mt % 3 (accumulators are issued in natural
//     order, a = t * n_mtiles + mt).  The running max of the passage that straddles D tiles therefore never changes
//     hands: it lives in a register of the thread that owns the (query tile, row) — no shared-memory carry, no
//     carry barrier, no rotated issue order.  With five resident tiles the warpgroups drain 2 / 2 / 1 accumulators
//     per D tile instead of 2.5 / 2.5; with three tiles (one Nq = 320 query per pass) one each.
//   * Barriers whose phases every waiter observes (a parity wait only tells consecutive phases apart):
//     `t_full[wg][stage]` — the issuer commits an accumulator to the barrier of the warpgroup that drains it;
//     `t_empty[issuer][stage]` — the drainer of accumulator a arrives on the barrier of the issuer of a + S, the
//     next user of that TMEM stage.  That makes an ODD stage count legal: three query tiles leave room for three
//     accumulator stages (flmr_scan_kernel has to fall back to two).
#pragma once
#include "flmr_scan_kernel.cuh"

namespace flmr {

constexpr int kEpi3Wgs = 3;
constexpr int kEpi3Warps = kEpi3Wgs * 4;
constexpr int kWarp3Red = kEpi3Warps;            // 12
constexpr int kWarp3Producer = kEpi3Warps + 1;   // 13
constexpr int kWarp3Mma = kEpi3Warps + 2;        // 14, 15
constexpr int kScan3Threads = (kWarp3Mma + kMmaWarps) * 32;   // 512

// TMEM column budget with any stage count from 2 to 4 (see the header comment).
__host__ __device__ inline int scan3_acc_stages(int n_mtiles) {
  const int s = (512 - kQCols * n_mtiles) / kTileN;
  return s > 4 ? 4 : s;
}

struct Scan3Smem {
  static constexpr int kOffBars = ScanSmem::kOffCarry;            // (no carry array in this kernel)
  static_assert(kOffBars % 8 == 0, "mbarriers need 8-byte alignment");
  static constexpr int kNumBars = 1 + 2 * kDStages + kEpi3Wgs * kMaxAccStages + kMmaWarps * kMaxAccStages + 4;
  static constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
  static constexpr int kBytes = kOffTmemPtr + 16 + 1024;
  static_assert(kBytes <= ScanSmem::kBytes, "must not need more shared memory than flmr_scan_kernel");
};

// ---- the kernel -------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScan3Threads, 1)
flmr_scan3_kernel(const __grid_constant__ CUtensorMap tmap_d, const ScanParams p) {
  constexpr int dbg = 0;
  using S = ScanSmem;
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment by pointer arithmetic on the __shared__ array (keeps the shared address space
  // visible to the compiler: LDS/STS instead of generic loads/stores)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;

  const uint32_t bar_base = smem_base + Scan3Smem::kOffBars;
  const uint32_t bar_q_full = bar_base;
  auto bar_d_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto bar_d_empty = [&](int s) { return bar_base + 8u * (1 + kDStages + s); };
  constexpr int kB0 = 1 + 2 * kDStages;
  auto bar_t_full = [&](int g, int s) { return bar_base + 8u * (kB0 + g * kMaxAccStages + s); };
  auto bar_t_empty = [&](int i, int s) {
    return bar_base + 8u * (kB0 + kEpi3Wgs * kMaxAccStages + i * kMaxAccStages + s);
  };
  constexpr int kB1 = kB0 + (kEpi3Wgs + kMmaWarps) * kMaxAccStages;
  auto bar_p_full = [&](int b) { return bar_base + 8u * (kB1 + b); };
  auto bar_p_empty = [&](int b) { return bar_base + 8u * (kB1 + 2 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + Scan3Smem::kOffTmemPtr);

  const int32_t row_begin = p.cta_row_begin[cta];
  const int64_t tile_base = p.cta_tile_base[cta];
  const int n_tiles = static_cast<int>(p.cta_tile_base[cta + 1] - tile_base);
  const int n_mtiles = p.n_mtiles;
  const uint32_t acc_stages = static_cast<uint32_t>(scan3_acc_stages(n_mtiles));   // 2, 3 or 4
  const uint32_t acc_col0 = static_cast<uint32_t>(kQCols * n_mtiles);
  const uint32_t tile_stage_step = static_cast<uint32_t>(n_mtiles) % acc_stages;  // stage of (t+1, 0) - stage of (t, 0)

  // ---- one-time setup --------------------------------------------------------------------------
  if (warp == kWarp3Producer && lane == 0) {
    tma_prefetch_desc(&tmap_d);
    mbar_init(bar_q_full, 4);        // one arrive per query-staging warp
    for (int s = 0; s < kDStages; ++s) {
      mbar_init(bar_d_full(s), 1);
      mbar_init(bar_d_empty(s), kMmaWarps);
    }
    for (int s = 0; s < kMaxAccStages; ++s) {
      for (int g = 0; g < kEpi3Wgs; ++g) mbar_init(bar_t_full(g, s), 1);
      for (int i = 0; i < kMmaWarps; ++i) mbar_init(bar_t_empty(i, s), 4);  // one arrive per warp of the drainer
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_p_full(b), kEpi3Warps);  // partial sums of a D tile complete
      mbar_init(bar_p_empty(b), kRedWarps);  // ... and consumed by the reducer warp
    }
    mbar_fence_init();
  }
  if (warp == kWarp3Mma) {  // (the same warp deallocates at the end)
    tmem_alloc<512>(smem_base + Scan3Smem::kOffTmemPtr);
  }
  if (warp < kWarp3Producer) {
    const int et = threadIdx.x;
    constexpr int kInitThreads = (kEpi3Warps + kRedWarps) * 32;
    // top-k lists start empty (key 0 sorts below every real candidate)
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem + S::kOffKeys);
    for (int i = et; i < kNqMax * kMaxK; i += kInitThreads) keys[i] = 0ull;
    if (et < kNqMax) {
      reinterpret_cast<uint64_t*>(smem + S::kOffMinKey)[et] = 0ull;
      reinterpret_cast<int*>(smem + S::kOffMinPos)[et] = 0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == kWarp3Producer) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop (warp-uniform control flow keeps addresses in uniform
    // registers); one elected lane issues the bulk copies.
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % kDStages;
      const uint32_t ph = (t / kDStages) & 1;
      mbar_wait(bar_d_empty(s), ph ^ 1u, p.status, kDevTimeoutProducer);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(bar_d_full(s), kDTileBytes);
        const uint32_t dst = smem_base + S::kOffD + s * kDTileBytes;
        const int32_t row = row_begin + t * kTileN;
        tma_load_2d(dst, &tmap_d, bar_d_full(s), 0, row, kPolicyEvictFirst);
        tma_load_2d(dst + kDKBlockBytes, &tmap_d, bar_d_full(s), 64, row, kPolicyEvictFirst);
      }
      __syncwarp();
    }
  } else if (warp >= kWarp3Mma) {
    // ===================== MMA issuers =====================
    // Issuer i owns the accumulators with a = i (mod 2) in issue order (a = t * n_mtiles + mt), as in
    // flmr_scan_kernel.  It waits for its accumulator's TMEM stage on t_empty[i][stage] — the drainer of the
    // previous user of that stage (accumulator a - S) arrives exactly there — and commits the finished
    // accumulator to t_full[mt % 3][stage], the barrier of the warpgroup that drains query tile mt.
    const uint32_t iw = static_cast<uint32_t>(warp - kWarp3Mma);
    constexpr uint32_t idesc = make_idesc_bf16_f32(kTileM, kTileN);
    mbar_wait(bar_q_full, 0, p.status, kDevTimeoutMma);  // queries are in TMEM
    tc_fence_after_sync();
    uint32_t te_parity = 0;                    // bit s: parity of this issuer's next wait on t_empty[iw][s]
    uint32_t stage0 = 0;                       // stage of accumulator (t, 0)
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % kDStages;
      const uint32_t ph = (t / kDStages) & 1;
      mbar_wait(bar_d_full(s), ph, p.status, kDevTimeoutMma);
      tc_fence_after_sync();
      const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_base + S::kOffD + s * kDTileBytes);
      const uint32_t a_first = static_cast<uint32_t>(t) * n_mtiles;
#pragma unroll 1
      for (uint32_t mt = (a_first ^ iw) & 1u; mt < static_cast<uint32_t>(n_mtiles); mt += 2) {
        const uint32_t a = a_first + mt;
        uint32_t as = stage0 + mt;             // (stage0 + mt) mod acc_stages, mt <= 4, acc_stages >= 2
        as -= (as >= acc_stages) ? acc_stages : 0u;
        as -= (as >= acc_stages) ? acc_stages : 0u;
        as -= (as >= acc_stages) ? acc_stages : 0u;
        if (a >= acc_stages) {                 // the stage has been used before: wait for its hand-back
          mbar_wait(bar_t_empty(iw, as), (te_parity >> as) & 1u, p.status, kDevTimeoutMma);
          te_parity ^= 1u << as;
          tc_fence_after_sync();
        }
        const uint32_t d_tmem = tmem_base + acc_col0 + as * kTileN;
        const uint32_t a_tmem = tmem_base + mt * kQCols;
        uint32_t wg = mt;                      // mt % 3 for mt <= 4
        wg -= (wg >= 3u) ? 3u : 0u;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < kDim / 16; ++k) {
            const uint64_t b_desc =
                b_desc0 + static_cast<uint64_t>(((k >> 2) * kDKBlockBytes + (k & 3) * 32) >> 4);
            tc_mma_ts(d_tmem, a_tmem + k * 8, b_desc, idesc, k > 0 ? 1u : 0u);
          }
          tc_commit(bar_t_full(wg, as));       // accumulator complete -> the warpgroup of query tile mt
        }
        __syncwarp();
      }
      // this issuer's MMAs on the D stage are complete -> producer (both issuers must arrive)
      if (elect_one_sync()) tc_commit(bar_d_empty(s));
      __syncwarp();
      stage0 += tile_stage_step;
      stage0 -= (stage0 >= acc_stages) ? acc_stages : 0u;
    }
  } else if (warp < kEpi3Warps) {
    // ===================== epilogue (TMEM drain), three warpgroups =====================
    // Warpgroup g drains query tiles g and g + 3 of every D tile.  The running max of the passage straddling
    // D tiles stays in a register of the thread that owns the (query tile, row).
    const int wg = warp >> 2;         // epilogue warp 0..11 -> warpgroup 0..2
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    float* partial = reinterpret_cast<float*>(smem + S::kOffPartial);
    float* lane_part = reinterpret_cast<float*>(smem + S::kOffLanePart) + lane;
    const float init = p.init_val;

    // ---- stage the resident queries into tensor memory (warpgroup 0: one warp per lane quadrant).
    if (wg == 0) {
      for (int mt = 0; mt < n_mtiles; ++mt) {
        const uint4* src = p.q_pad + (static_cast<int64_t>(mt) * kTileM + quad * 32 + lane) * 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t w[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 x = __ldg(src + h * 8 + i);
            w[4 * i] = x.x;
            w[4 * i + 1] = x.y;
            w[4 * i + 2] = x.z;
            w[4 * i + 3] = x.w;
          }
          FLMR_TMEM_ST32(tmem_base + lane_base + mt * kQCols + h * 32, w);
        }
      }
      tmem_wait_st();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_q_full);
    }

    float m_lo = init, m_hi = init;            // running maxima of query tiles wg and wg + 3
    uint32_t tf_parity = 0;                    // bit s: parity of this warpgroup's next wait on t_full[wg][s]
    uint32_t stage0 = 0;                       // stage of accumulator (t, 0)
    uint32_t mask_next = (n_tiles > 0) ? __ldg(p.tile_end_mask + tile_base) : 0u;
    for (int t = 0; t < n_tiles; ++t) {
      const uint32_t mask = mask_next;
      if (t + 1 < n_tiles) mask_next = __ldg(p.tile_end_mask + tile_base + t + 1);
      const int buf = t & 1;
      // the reducer must have consumed the partial sums of tile t-2 before this buffer is reused
      mbar_wait(bar_p_empty(buf), ((static_cast<uint32_t>(t) >> 1) & 1u) ^ 1u, p.status,
                kDevTimeoutEpilogue);
      const uint32_t a_first = static_cast<uint32_t>(t) * n_mtiles;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int mt = wg + 3 * h;
        if (mt >= n_mtiles) break;
        const uint32_t a = a_first + mt;
        uint32_t as = stage0 + mt;
        as -= (as >= acc_stages) ? acc_stages : 0u;
        as -= (as >= acc_stages) ? acc_stages : 0u;
        as -= (as >= acc_stages) ? acc_stages : 0u;
        float m = h ? m_hi : m_lo;
        mbar_wait(bar_t_full(wg, as), (tf_parity >> as) & 1u, p.status, kDevTimeoutEpilogue);
        tf_parity ^= 1u << as;
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + lane_base + acc_col0 + as * kTileN;
        uint32_t v[kChunks][32];
        float* partial_rb = partial + (buf * kRbMax + mt * 4 + quad) * kSlots;
        float* lane_part_rb = lane_part + ((buf * kFastSlots) * kRbMax + mt * 4 + quad) * kLaneStride;
        int slot = 0;
        // chunk 0 first; the remaining chunks stream in while chunk 0 is folded
        FLMR_TMEM_LD32(v[0], taddr);
        FLMR_TMEM_WAIT_LD32(v[0]);
#pragma unroll
        for (int c = 1; c < kChunks; ++c) FLMR_TMEM_LD32(v[c], taddr + 32 * c);
        process_chunk(v[0], mask & 0xFFu, m, init, partial_rb, lane_part_rb, slot, lane);
#pragma unroll
        for (int c = 1; c < kChunks; ++c) FLMR_TMEM_WAIT_LD32(v[c]);
        // every column is in registers: hand the TMEM stage to the issuer of its next user, a + acc_stages
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_t_empty((a + acc_stages) & 1u, as));
#pragma unroll
        for (int c = 1; c < kChunks; ++c)
          process_chunk(v[c], (mask >> (8 * c)) & 0xFFu, m, init, partial_rb, lane_part_rb, slot, lane);
        if (h) m_hi = m; else m_lo = m;
      }
      // this warp's per-lane maxima / partial sums of tile t are written: order every lane's stores
      // before lane 0's arrive (release); the reducer's wait is the matching acquire
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_full(buf));
      stage0 += tile_stage_step;
      stage0 -= (stage0 >= acc_stages) ? acc_stages : 0u;
    }
    tc_fence_before_sync();
  } else if (warp == kWarp3Red) {
    // ===================== reducer: score finalisation + per-CTA top-k =====================
    // Per D tile and per (query, passage ending in the tile): sum the row-block partials in fixed
    // order (deterministic), add/store partial scores if requested, offer to the top-k list.
    const int rw = 0;
    const float* partial = reinterpret_cast<const float*>(smem + S::kOffPartial);
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem + S::kOffKeys);
    uint64_t* minkey_s = reinterpret_cast<uint64_t*>(smem + S::kOffMinKey);
    int* minpos_s = reinterpret_cast<int*>(smem + S::kOffMinPos);
    const int n_tiles_red = (dbg == 4 || dbg == 5) ? 0 : n_tiles;
    uint32_t mask_next = 0;
    int32_t fpid_next = 0;
    if (n_tiles_red > 0) {
      mask_next = __ldg(p.tile_end_mask + tile_base);
      fpid_next = __ldg(p.tile_first_pid + tile_base);
    }
    for (int t = 0; t < n_tiles_red; ++t) {
      const uint32_t mask = mask_next;
      const int32_t first_pid = fpid_next;
      if (t + 1 < n_tiles_red) {
        mask_next = __ldg(p.tile_end_mask + tile_base + t + 1);
        fpid_next = __ldg(p.tile_first_pid + tile_base + t + 1);
      }
      const int buf = t & 1;
      mbar_wait(bar_p_full(buf), (static_cast<uint32_t>(t) >> 1) & 1u, p.status, kDevTimeoutEpilogue);
      const int n_slots = __popc(mask);
      const float* lane_part0 = reinterpret_cast<const float*>(smem + S::kOffLanePart);
#pragma unroll 1
      for (int slot = 0; slot < n_slots; ++slot) {
        const int64_t pid = static_cast<int64_t>(first_pid) + slot;
        const float* lp_slot = lane_part0 + ((buf * kFastSlots + slot) * kRbMax) * kLaneStride;
        if (slot < kFastSlots && p.rbq <= p.lane_mode_max_rbq) {
          // many short queries per pass: lane = query, each lane sums its query's 32 * rbq lane maxima
          // (transposed read of the stride-33 layout: conflict-free for rbq = 1)
#pragma unroll 1
          for (int b0 = 0; b0 < p.nq_pass; b0 += 32) {
            const int b = b0 + lane;
            const bool valid = b < p.nq_pass;
            float sc = 0.f;
            uint64_t key = 0ull;
            if (valid) {
              const float* lp = lp_slot + (b * p.rbq) * kLaneStride;
              for (int r = 0; r < p.rbq; ++r) {
#pragma unroll 8
                for (int j = 0; j < 32; ++j) sc += lp[r * kLaneStride + j];
              }
              const int64_t gi = static_cast<int64_t>(b) * p.n_passages + pid;
              if (p.acc_in) sc += __ldg(p.acc_in + gi);
              if (p.acc_out) p.acc_out[gi] = sc;
              key = (static_cast<uint64_t>(float_to_ordered(sc)) << 32) |
                    static_cast<uint64_t>(0xFFFFFFFFu - static_cast<uint32_t>(pid));
            }
            if (p.k > 0) {
              uint32_t hits = __ballot_sync(0xffffffffu, valid && key > minkey_s[valid ? b : 0]);
              while (hits) {  // rare after warm-up: one list update at a time
                const int src = __ffs(hits) - 1;
                hits &= hits - 1;
                const uint64_t cand = shfl64(key, src);
                const int cb = b0 + src;
                uint64_t minkey = minkey_s[cb];
                int minpos = minpos_s[cb];
                if (cand > minkey) {
                  topk_replace_min(keys + cb * kMaxK, p.k, cand, minkey, minpos, lane);
                  if (lane == 0) {
                    minkey_s[cb] = minkey;
                    minpos_s[cb] = minpos;
                  }
                  __syncwarp();
                }
              }
            }
          }
        } else {
#pragma unroll 1
          for (int b = 0; b < p.nq_pass; ++b) {
            // score of (query b, passage pid): the whole warp sums the row-block partials in fixed order
            float sc = 0.f;
            if (slot < kFastSlots) {
              const float* lp = lp_slot + (b * p.rbq) * kLaneStride + lane;
#pragma unroll 2
              for (int r = 0; r < p.rbq; ++r) sc += lp[r * kLaneStride];
              sc = warp_sum(sc);
            } else {
              const float* pr = partial + (buf * kRbMax + b * p.rbq) * kSlots + slot;
#pragma unroll 2
              for (int r = 0; r < p.rbq; ++r) sc += pr[r * kSlots];
            }
            const int64_t gi = static_cast<int64_t>(b) * p.n_passages + pid;
            if (p.acc_in) sc += __ldg(p.acc_in + gi);
            if (p.acc_out && lane == 0) p.acc_out[gi] = sc;
            if (p.k > 0) {
              const uint64_t key = (static_cast<uint64_t>(float_to_ordered(sc)) << 32) |
                                   static_cast<uint64_t>(0xFFFFFFFFu - static_cast<uint32_t>(pid));
              uint64_t minkey = minkey_s[b];
              if (key > minkey) {  // warp-uniform; rare after warm-up
                int minpos = minpos_s[b];
                topk_replace_min(keys + b * kMaxK, p.k, key, minkey, minpos, lane);
                if (lane == 0) {
                  minkey_s[b] = minkey;
                  minpos_s[b] = minpos;
                }
                __syncwarp();
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_empty(buf));
    }
    // ---- publish this CTA's candidates ----
    if (p.k > 0) {
      __syncwarp();
      for (int b = rw; b < p.nq_pass; b += kRedWarps) {
        uint64_t* dst = p.cand_keys + (static_cast<int64_t>(cta) * p.cand_q_stride + p.cand_q_first + b) * p.k;
        for (int i = lane; i < p.k; i += 32) dst[i] = keys[b * kMaxK + i];
      }
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  __syncthreads();
  if (warp == kWarp3Mma) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace flmr
