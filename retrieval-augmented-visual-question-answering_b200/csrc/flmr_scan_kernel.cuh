// flmr_scan_kernel.cuh — the fused late-interaction scan:
//
//   for every passage p of this GPU's shard and every resident query b:
//       score[b][p] = sum_i max_{j < len_p} <Q[b][i], D[p][j]>          
//   + per-CTA running top-k of those scores, so no score matrix reaches HBM.
//
// Replaces, in one kernel, the reference's `D_packed @ Q.T` GEMM, the StridedTensor padding +
// colbert_score_reduce masked max/sum (CB/modeling/colbert.py:235-311), the CPU
// segmented_maxsim.cpp reduction (CB/modeling/segmented_maxsim.cpp:22-93) and the per-query
// `scores.sort()` of IndexScorer.rank (CB/search/index_storage.py:95).
//
// Structure (one persistent CTA per SM, 6 warps, warp-specialised):
//   warp 0 / lane 0 : TMA producer.  Loads the resident query tiles once, then streams the CTA's
//                     contiguous range of passage tokens through a ring of D stages
//                     (TILE_N tokens x 128 dims bf16, 128B-swizzled, 2 boxes per stage).
//   warp 1 / lane 0 : tcgen05.mma issuer.  For every D stage and every resident 128-row query tile
//                     issues 8 MMAs (K = 8 x 16 = 128): acc[128 x TILE_N] = Qtile . Dtile^T into
//                     one of 512/TILE_N TMEM accumulator stages; tcgen05.commit signals epilogue
//                     (accumulator full) and producer (D stage free).
//   warps 2..5      : epilogue.  TMEM lane = query token, TMEM column = passage token, so the max
//                     over a passage's tokens is a per-thread running max over columns (FMNMX3, no
//                     shuffles); at a passage end (bit in the tile's end-mask) the warp sums its 32
//                     lanes with shuffles and lane 0 stores one partial per (32-row block, passage).
//                     After all query tiles of a D tile: per (query, passage) the row-block
//                     partials are summed in fixed order (deterministic), optionally
//                     accumulated/stored to HBM, and offered to the per-CTA top-k list.
//
// Layout contracts (see DESIGN.md "Data layout"):
//   * passages are stored back to back, each padded to a multiple of 4 tokens by repeating its last
//     token (duplicates cannot change a max), so passage boundaries fall on 4-column groups and one
//     64-bit mask per tile marks the groups that end a passage;
//   * each query is padded to a multiple of 32 rows with zero rows (contribute exactly 0, like the
//     reference's masked query tokens), so a warp's 32 TMEM lanes always belong to one query.
#pragma once
#include "flmr_device.cuh"

namespace flmr {

constexpr int kDim = 128;
constexpr int kTileM = 128;               // query rows per MMA (UMMA M)
constexpr int kMtMax = 3;                 // resident 128-row query tiles per pass
constexpr int kRbMax = kMtMax * 4;        // resident 32-row blocks
constexpr int kNqMax = kRbMax;            // queries per pass (each has >= 1 row block)
constexpr int kMaxK = 128;                // fused top-k capacity (== FLMR_MAX_K)
constexpr int kGroup = 4;                 // token padding granularity (== FLMR_TOKEN_GROUP)
constexpr int kScanThreads = 192;
constexpr int kQTileBytes = kTileM * kDim * 2;   // 32 KiB: [2 k-blocks][128 rows][64 bf16]
constexpr int kQKBlockBytes = kTileM * 128;      // 16 KiB

struct ScanParams {
  // corpus partition (built once per corpus, see build_partition() in flmr_maxsim.cu)
  const int32_t* cta_row_begin;    // [n_ctas + 1] first stored row of each CTA's passage range
  const int64_t* cta_tile_base;    // [n_ctas + 1] index of each CTA's first tile in tile_* arrays
  const uint64_t* tile_end_mask;   // [n_tiles] bit g set <=> a passage ends with 4-token group g
  const int32_t* tile_first_pid;   // [n_tiles] local id of the first passage ending in the tile
  // resident queries of this pass
  int32_t n_mtiles;                // 128-row query tiles (1..kMtMax)
  int32_t nq_pass;                 // queries resident in this pass (1..kNqMax)
  int32_t rbq;                     // 32-row blocks per query in this pass
  float init_val;                  // -inf (true max) or 0 (reference CPU "ReLU" variant)
  // score plumbing
  const float* acc_in;             // [nq_pass][n_passages] partial scores of earlier row slices, or null
  float* acc_out;                  // [nq_pass][n_passages] scores (or partial scores) out, or null
  int64_t n_passages;
  // fused top-k
  int32_t k;                       // 0 = disabled
  uint64_t* cand_keys;             // [n_ctas][nq_pass][k]  (ordered score << 32 | ~local pid)
  // diagnostics
  int32_t debug_mode;              // 0 = product.  Timing-only experiments (results are garbage):
                                   // 1 = epilogue releases accumulators unread, 2 = TMEM reads but no
                                   // max/flush, 3 = mode 1 + MMA issue skipped (pure TMA streaming)
  int* status;
  uint64_t timeout_ns;
};

template <int TILE_N>
struct ScanCfg {
  static_assert(TILE_N == 128 || TILE_N == 256, "TILE_N must be 128 or 256");
  static constexpr int kDStages = (TILE_N == 128) ? 3 : 1;
  static constexpr int kDTileBytes = TILE_N * kDim * 2;
  static constexpr int kDKBlockBytes = TILE_N * 128;
  static constexpr int kAccStages = 512 / TILE_N;
  static constexpr int kSlots = TILE_N / kGroup;   // passage ends per tile, at most
  static constexpr int kChunks = TILE_N / 32;
  static constexpr int kOffQ = 0;
  static constexpr int kOffD = kMtMax * kQTileBytes;
  static constexpr int kOffPartial = kOffD + kDStages * kDTileBytes;
  static constexpr int kPartialBytes = 2 * kRbMax * kSlots * 4;
  static constexpr int kOffKeys = kOffPartial + kPartialBytes;
  static constexpr int kKeysBytes = kNqMax * kMaxK * 8;
  static constexpr int kOffMinKey = kOffKeys + kKeysBytes;
  static constexpr int kOffMinPos = kOffMinKey + kNqMax * 8;
  static constexpr int kOffCarry = kOffMinPos + kNqMax * 4;          // float[kMtMax * 128]
  static constexpr int kOffBars = (kOffCarry + kMtMax * kTileM * 4 + 7) / 8 * 8;
  static constexpr int kNumBars = 1 + 2 * kDStages + 2 * kAccStages;
  static constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
  static constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;  // + slack for 1024-B alignment
  static_assert(kSmemBytes <= 232448, "exceeds 227 KiB of shared memory per CTA");
};

// ---- epilogue helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float s) {
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  return s;
}

// 32 accumulator columns = 8 groups of 4 tokens.  `bits` bit g: a passage ends with group g.
// Compact on purpose (the epilogue runs one warp per SM sub-partition, so instruction-cache misses
// and branches are paid in full): 16 FMNMX for the 8 group maxima, then either a 3-instruction
// fold into the running max, or -- per passage END in this chunk -- one pass of predicated folds
// + a warp sum.
__device__ __forceinline__ void process_chunk(const uint32_t (&v)[32], uint32_t bits, float& m,
                                              float init, float* partial_rb, int& slot, int lane) {
  float gv[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float a = fmax3(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                          __uint_as_float(v[4 * g + 2]));
    gv[g] = fmaxf(a, __uint_as_float(v[4 * g + 3]));
  }
  if (bits == 0u) {
    const float x = fmax3(gv[0], gv[1], gv[2]);
    const float y = fmax3(gv[3], gv[4], gv[5]);
    m = fmax3(m, x, y);
    m = fmax3(m, gv[6], gv[7]);
    return;
  }
  uint32_t live = 0xFFu;  // groups not yet consumed by a finished passage
#pragma unroll 1
  while (bits) {          // warp-uniform: one iteration per passage ending in this chunk
    const uint32_t upto = bits ^ (bits - 1u);  // groups 0..(lowest set bit)
    const uint32_t seg = upto & live;
    float s = m;
#pragma unroll
    for (int g = 0; g < 8; ++g) s = fmaxf(s, ((seg >> g) & 1u) ? gv[g] : -INFINITY);
    s = warp_sum(s);
    if (lane == 0) partial_rb[slot] = s;
    ++slot;
    m = init;
    live &= ~upto;
    bits &= bits - 1u;
  }
#pragma unroll
  for (int g = 0; g < 8; ++g) m = fmaxf(m, ((live >> g) & 1u) ? gv[g] : -INFINITY);
}

template <int TILE_N>
__device__ __forceinline__ void epilogue_accumulator(uint32_t taddr, uint64_t mask, float& m,
                                                     float init, float* partial_rb, int lane,
                                                     uint32_t t_empty_bar, int debug_mode) {
  constexpr int kChunks = ScanCfg<TILE_N>::kChunks;
  uint32_t va[32], vb[32];
  int slot = 0;
  if (debug_mode == 1 || debug_mode == 3) {  // timing experiment: hand the accumulator back unread
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(t_empty_bar);
    return;
  }
  if (debug_mode == 2) mask = 0ull;          // timing experiment: no passage ends -> no flushes
  FLMR_TMEM_LD32(va, taddr);
#pragma unroll 1
  for (int c = 0; c < kChunks; c += 2) {
    FLMR_TMEM_WAIT_LD32(va);
    FLMR_TMEM_LD32(vb, taddr + (c + 1) * 32);
    process_chunk(va, static_cast<uint32_t>(mask >> (8 * c)) & 0xFFu, m, init, partial_rb, slot,
                  lane);
    FLMR_TMEM_WAIT_LD32(vb);
    if (c + 2 < kChunks) {
      FLMR_TMEM_LD32(va, taddr + (c + 2) * 32);
    } else {
      // every column of this accumulator is in registers: hand the TMEM stage back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty_bar);
    }
    process_chunk(vb, static_cast<uint32_t>(mask >> (8 * (c + 1))) & 0xFFu, m, init, partial_rb,
                  slot, lane);
  }
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  const uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
  const uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Replace the current minimum of an unsorted k-entry list by `cand`, then recompute the minimum.
__device__ __noinline__ void topk_replace_min(uint64_t* keys, int k, uint64_t cand,
                                              uint64_t& minkey, int& minpos, int lane) {
  if (lane == 0) keys[minpos] = cand;
  __syncwarp();
  uint64_t mk = ~0ull;
  int mp = 0x7fffffff;
  for (int i = lane; i < k; i += 32) {
    const uint64_t x = keys[i];
    if (x < mk) {
      mk = x;
      mp = i;
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const uint64_t ok = shfl64(mk, lane ^ off);
    const int op = __shfl_xor_sync(0xffffffffu, mp, off);
    if (ok < mk || (ok == mk && op < mp)) {
      mk = ok;
      mp = op;
    }
  }
  minkey = mk;
  minpos = mp;
  __syncwarp();
}

// ---- the kernel -------------------------------------------------------------------------------
template <int TILE_N>
__global__ void __launch_bounds__(kScanThreads, 1)
flmr_scan_kernel(const __grid_constant__ CUtensorMap tmap_q,
                 const __grid_constant__ CUtensorMap tmap_d, const ScanParams p) {
  using Cfg = ScanCfg<TILE_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;

  const uint32_t bar_base = smem_base + Cfg::kOffBars;
  const uint32_t bar_q_full = bar_base;
  auto bar_d_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto bar_d_empty = [&](int s) { return bar_base + 8u * (1 + Cfg::kDStages + s); };
  auto bar_t_full = [&](int s) { return bar_base + 8u * (1 + 2 * Cfg::kDStages + s); };
  auto bar_t_empty = [&](int s) {
    return bar_base + 8u * (1 + 2 * Cfg::kDStages + Cfg::kAccStages + s);
  };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + Cfg::kOffTmemPtr);

  const int32_t row_begin = p.cta_row_begin[cta];
  const int32_t row_end = p.cta_row_begin[cta + 1];
  const int64_t tile_base = p.cta_tile_base[cta];
  const int n_tiles = static_cast<int>(p.cta_tile_base[cta + 1] - tile_base);
  (void)row_end;

  // ---- one-time setup --------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_d);
    mbar_init(bar_q_full, 1);
    for (int s = 0; s < Cfg::kDStages; ++s) {
      mbar_init(bar_d_full(s), 1);
      mbar_init(bar_d_empty(s), 1);
    }
    for (int s = 0; s < Cfg::kAccStages; ++s) {
      mbar_init(bar_t_full(s), 1);
      mbar_init(bar_t_empty(s), 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc<512>(smem_base + Cfg::kOffTmemPtr);
  }
  if (warp >= 2) {
    // top-k lists start empty (key 0 sorts below every real candidate)
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem + Cfg::kOffKeys);
    for (int i = threadIdx.x - 64; i < kNqMax * kMaxK; i += 128) keys[i] = 0ull;
    uint64_t* minkey = reinterpret_cast<uint64_t*>(smem + Cfg::kOffMinKey);
    int* minpos = reinterpret_cast<int*>(smem + Cfg::kOffMinPos);
    if (threadIdx.x - 64 < kNqMax) {
      minkey[threadIdx.x - 64] = 0ull;
      minpos[threadIdx.x - 64] = 0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q_full, static_cast<uint32_t>(p.n_mtiles) * kQTileBytes);
      for (int mt = 0; mt < p.n_mtiles; ++mt) {
        const uint32_t dst = smem_base + Cfg::kOffQ + mt * kQTileBytes;
        tma_load_2d(dst, &tmap_q, bar_q_full, 0, mt * kTileM, kPolicyEvictLast);
        tma_load_2d(dst + kQKBlockBytes, &tmap_q, bar_q_full, 64, mt * kTileM, kPolicyEvictLast);
      }
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % Cfg::kDStages;
        const uint32_t ph = (t / Cfg::kDStages) & 1;
        mbar_wait(bar_d_empty(s), ph ^ 1u, p.status, kDevTimeoutProducer, p.timeout_ns);
        mbar_arrive_expect_tx(bar_d_full(s), Cfg::kDTileBytes);
        const uint32_t dst = smem_base + Cfg::kOffD + s * Cfg::kDTileBytes;
        const int32_t row = row_begin + t * TILE_N;
        tma_load_2d(dst, &tmap_d, bar_d_full(s), 0, row, kPolicyEvictFirst);
        tma_load_2d(dst + Cfg::kDKBlockBytes, &tmap_d, bar_d_full(s), 64, row, kPolicyEvictFirst);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(kTileM, TILE_N);
      mbar_wait(bar_q_full, 0, p.status, kDevTimeoutMma, p.timeout_ns);
      tc_fence_after_sync();
      uint32_t acc = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % Cfg::kDStages;
        const uint32_t ph = (t / Cfg::kDStages) & 1;
        mbar_wait(bar_d_full(s), ph, p.status, kDevTimeoutMma, p.timeout_ns);
        tc_fence_after_sync();
        const uint32_t d_addr = smem_base + Cfg::kOffD + s * Cfg::kDTileBytes;
        for (int mt = 0; mt < p.n_mtiles; ++mt, ++acc) {
          const uint32_t as = acc % Cfg::kAccStages;
          const uint32_t aph = (acc / Cfg::kAccStages) & 1;
          mbar_wait(bar_t_empty(as), aph ^ 1u, p.status, kDevTimeoutMma, p.timeout_ns);
          tc_fence_after_sync();
          const uint32_t q_addr = smem_base + Cfg::kOffQ + mt * kQTileBytes;
          const uint32_t d_tmem = tmem_base + as * TILE_N;
          if (p.debug_mode != 3) {
#pragma unroll
          for (int k = 0; k < kDim / 16; ++k) {
            const uint64_t a_desc =
                make_kmajor_sw128_desc(q_addr + (k >> 2) * kQKBlockBytes + (k & 3) * 32);
            const uint64_t b_desc =
                make_kmajor_sw128_desc(d_addr + (k >> 2) * Cfg::kDKBlockBytes + (k & 3) * 32);
            tc_mma_ss(d_tmem, a_desc, b_desc, idesc, k > 0 ? 1u : 0u);
          }
          }
          tc_commit(bar_t_full(as));  // accumulator complete -> epilogue
        }
        tc_commit(bar_d_empty(s));    // all MMAs reading this D stage complete -> producer
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;          // epilogue warp index 0..3 (owns queries ew, ew+4, ...)
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access
    float* partial = reinterpret_cast<float*>(smem + Cfg::kOffPartial);
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem + Cfg::kOffKeys);
    uint64_t* minkey_s = reinterpret_cast<uint64_t*>(smem + Cfg::kOffMinKey);
    int* minpos_s = reinterpret_cast<int*>(smem + Cfg::kOffMinPos);
    const float init = p.init_val;
    const int rows_valid_rb = p.nq_pass * p.rbq;  // row blocks that belong to a query

    // running max of the passage that straddles consecutive D tiles, per (query tile, row)
    float* carry = reinterpret_cast<float*>(smem + Cfg::kOffCarry) + (threadIdx.x - 64);
#pragma unroll
    for (int i = 0; i < kMtMax; ++i) carry[i * kTileM] = init;

    uint32_t acc = 0;
    uint64_t mask_next = 0;
    int32_t fpid_next = 0;
    if (n_tiles > 0) {
      mask_next = __ldg(p.tile_end_mask + tile_base);
      fpid_next = __ldg(p.tile_first_pid + tile_base);
    }
    for (int t = 0; t < n_tiles; ++t) {
      const uint64_t mask = mask_next;
      const int32_t first_pid = fpid_next;
      if (t + 1 < n_tiles) {
        mask_next = __ldg(p.tile_end_mask + tile_base + t + 1);
        fpid_next = __ldg(p.tile_first_pid + tile_base + t + 1);
      }
      const int buf = t & 1;
#pragma unroll 1
      for (int mt = 0; mt < p.n_mtiles; ++mt) {
        {
          float m = carry[mt * kTileM];
          const uint32_t as = acc % Cfg::kAccStages;
          const uint32_t aph = (acc / Cfg::kAccStages) & 1;
          mbar_wait(bar_t_full(as), aph, p.status, kDevTimeoutEpilogue, p.timeout_ns);
          tc_fence_after_sync();
          const int rb = mt * 4 + quad;
          float* partial_rb = partial + (buf * kRbMax + rb) * Cfg::kSlots;
          const uint32_t taddr = tmem_base + as * TILE_N + (static_cast<uint32_t>(quad * 32) << 16);
          epilogue_accumulator<TILE_N>(taddr, mask, m, init, partial_rb, lane, bar_t_empty(as),
                                       p.debug_mode);
          carry[mt * kTileM] = m;
          ++acc;
        }
      }
      // all row blocks of this D tile have written their partials
      named_bar_sync(1, 128);

      // ---- finalize: per (query, passage ending in this tile) ----
      const int n_slots = __popcll(mask);
#pragma unroll 1
      for (int b = ew; b < p.nq_pass; b += 4) {
        uint64_t minkey = minkey_s[b];
        int minpos = minpos_s[b];
        bool dirty = false;
#pragma unroll 1
        for (int s0 = 0; s0 < n_slots; s0 += 32) {
          const int slot = s0 + lane;
          const bool valid = slot < n_slots;
          float sc = 0.f;
          uint64_t key = 0ull;
          if (valid) {
            const float* pr = partial + (buf * kRbMax + b * p.rbq) * Cfg::kSlots + slot;
#pragma unroll 2
            for (int r = 0; r < p.rbq; ++r) sc += pr[r * Cfg::kSlots];
            const int64_t pid = static_cast<int64_t>(first_pid) + slot;
            const int64_t gi = static_cast<int64_t>(b) * p.n_passages + pid;
            if (p.acc_in) sc += __ldg(p.acc_in + gi);
            if (p.acc_out) p.acc_out[gi] = sc;
            key = (static_cast<uint64_t>(float_to_ordered(sc)) << 32) |
                  static_cast<uint64_t>(0xFFFFFFFFu - static_cast<uint32_t>(pid));
          }
          if (p.k > 0) {
            uint32_t hits = __ballot_sync(0xffffffffu, valid && key > minkey);
            while (hits) {
              const int src = __ffs(hits) - 1;
              hits &= hits - 1;
              const uint64_t cand = shfl64(key, src);
              if (cand > minkey) {
                topk_replace_min(keys + b * kMaxK, p.k, cand, minkey, minpos, lane);
                dirty = true;
              }
            }
          }
        }
        if (dirty && lane == 0) {
          minkey_s[b] = minkey;
          minpos_s[b] = minpos;
        }
      }
      (void)rows_valid_rb;
    }

    // ---- publish this CTA's candidates ----
    if (p.k > 0) {
      __syncwarp();
      for (int b = ew; b < p.nq_pass; b += 4) {
        uint64_t* dst = p.cand_keys + (static_cast<int64_t>(cta) * p.nq_pass + b) * p.k;
        for (int i = lane; i < p.k; i += 32) dst[i] = keys[b * kMaxK + i];
      }
    }
    tc_fence_before_sync();
  }

  // ---- teardown -----------------------------------------------------------------------------------
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace flmr
