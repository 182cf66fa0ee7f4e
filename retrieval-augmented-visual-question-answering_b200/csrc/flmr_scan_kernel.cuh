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
// Structure (one persistent CTA per SM, 12 warps, warp-specialised; accumulators are numbered in
// issue order, a = t * n_mtiles + j, where j-th in D tile t is query tile mt = j — except that with an
// odd tile count the odd D tiles start with the LAST query tile, see the epilogue):
//   warps 0..7   : epilogue, two warpgroups; warpgroup g drains the accumulators with (a & 1) == g.
//                  TMEM lane = query token, TMEM column = passage token: the max over a passage's
//                  tokens is a per-thread running max over columns (FMNMX3, no shuffles); at a passage
//                  end (bit in the tile's end mask) the warp sums its 32 lanes and lane 0 stores one
//                  partial per (32-row block, passage).  The running max of the passage straddling
//                  two D tiles lives in shared memory per (query tile, row); when n_mtiles is odd the
//                  last query tile's changes hands between the warpgroups through an mbarrier (all
//                  other query tiles always meet the same warpgroup thanks to the rotated order).
//   warp 8       : reducer.  Per D tile (mbarrier-fed, double-buffered partials): per (query, passage
//                  ending in the tile) sums the row-block partials in fixed order (deterministic),
//                  optionally accumulates / stores scores to HBM, and maintains the per-CTA top-k.
//   warp 9       : TMA producer.  Streams the CTA's contiguous range of passage tokens through a
//                  7-deep ring of D stages (96 tokens x 128 dims bf16, 128B-swizzled, 2 boxes).
//   warps 10..11 : tcgen05.mma issuers; issuer i owns the accumulators with (a & 1) == i.  The
//                  queries are STATIONARY IN TENSOR MEMORY (A operand from TMEM: up to 5 tiles of
//                  128 rows x 128 dims = 5 x 64 columns), so shared memory only feeds the B operand.
//                  Per accumulator: 8 MMAs (K = 8 x 16), acc[128 x 96] = Qtile . Dtile^T, into one of
//                  2 or 4 TMEM accumulator stages; tcgen05.commit signals the epilogue / frees the
//                  D stage.  With an even stage count every stage and its mbarrier pair belong to
//                  exactly one issuer and one epilogue warpgroup.
//
// Layout contracts (see DESIGN.md "Data layout"):
//   * passages are stored back to back, each padded to a multiple of 4 tokens by repeating its last
//     token (duplicates cannot change a max), so passage boundaries fall on 4-column groups and one
//     mask per tile marks the groups that end a passage;
//   * each query is padded to a multiple of 32 rows with zero rows (contribute exactly 0, like the
//     reference's masked query tokens), so a warp's 32 TMEM lanes always belong to one query.
#pragma once
#include "../../include/flmr_maxsim.h"
#include "flmr_device.cuh"

namespace flmr {

constexpr int kDim = 128;
constexpr int kTileM = 128;               // query rows per MMA (UMMA M)
constexpr int kTileN = FLMR_TILE_TOKENS;  // passage tokens per streamed tile (UMMA N)
constexpr int kMtMax = 5;                 // resident 128-row query tiles per pass (5 x 64 TMEM cols)
constexpr int kRbMax = kMtMax * 4;        // resident 32-row blocks
constexpr int kNqMax = kRbMax;            // queries per pass (each has >= 1 row block)
constexpr int kMaxK = 128;                // fused top-k capacity (== FLMR_MAX_K)
constexpr int kGroup = 4;                 // token padding granularity (== FLMR_TOKEN_GROUP)
constexpr int kSlots = kTileN / kGroup;   // passage ends per tile, at most (24 <= 32 lanes)
constexpr int kChunks = kTileN / 32;      // 32-column chunks per accumulator
constexpr int kQCols = kDim / 2;          // TMEM columns of one query tile (bf16 pairs)
constexpr int kMaxAccStages = 4;
constexpr int kFastSlots = 4;             // passage ends per tile whose lane sums are left to the reducer
constexpr int kLaneStride = 33;           // padded row of 32 lane values (conflict-free transposed reads)
constexpr int kDStages = (kTileN == 96) ? 7 : 10;
constexpr int kDTileBytes = kTileN * kDim * 2;   // 24 KiB: [2 k-blocks][96 rows][64 bf16]
constexpr int kDKBlockBytes = kTileN * 128;      // 12 KiB
// Warp roles (see the header comment).
constexpr int kEpiWarps = 8;                      // warps 0..7: two warpgroups draining TMEM
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kRedWarps = 1;                      // warp 8: score finalisation + top-k
constexpr int kWarpRed0 = kEpiWarps;
constexpr int kWarpProducer = kEpiWarps + kRedWarps;   // warp 9
constexpr int kMmaWarps = 2;                           // warps 10..11: issuer i owns accumulators a = i (mod 2)
constexpr int kWarpMma = kWarpProducer + 1;
constexpr int kScanThreads = (kWarpMma + kMmaWarps) * 32;

static_assert(kTileN == 96 || kTileN == 64, "tile width must be 96 or 64 tokens");
static_assert(kDKBlockBytes % 1024 == 0, "swizzle atoms need 1024-B aligned k-blocks");

struct ScanParams {
  // corpus partition (built once per corpus, see build_partition() in flmr_maxsim.cu)
  const int32_t* cta_row_begin;    // [n_ctas + 1] first stored row of each CTA's passage range
  const int64_t* cta_tile_base;    // [n_ctas + 1] index of each CTA's first tile in tile_* arrays
  const uint32_t* tile_end_mask;   // [n_tiles] bit g set <=> a passage ends with 4-token group g
  const int32_t* tile_first_pid;   // [n_tiles] local id of the first passage ending in the tile
  // resident queries of this pass
  const uint4* q_pad;              // bf16 [n_mtiles * 128][128], zero-padded (stage_queries kernel)
  int32_t n_mtiles;                // 128-row query tiles (1..kMtMax)
  int32_t nq_pass;                 // queries resident in this pass (1..kNqMax)
  int32_t rbq;                     // 32-row blocks per query in this pass
  int32_t lane_mode_max_rbq;       // reducer: lane-per-query summation when rbq <= this, else warp-per-item
  float init_val;                  // -inf (true max) or 0 (reference CPU "ReLU" variant)
  // score plumbing
  const float* acc_in;             // [nq_pass][n_passages] partial scores of earlier row slices, or null
  float* acc_out;                  // [nq_pass][n_passages] scores (or partial scores) out, or null
  int64_t n_passages;
  // fused top-k
  int32_t k;                       // 0 = disabled
  uint64_t* cand_keys;             // [n_ctas][cand_q_stride][k]  (ordered score << 32 | ~local pid)
  int32_t cand_q_stride;           // queries of the whole call chunk (one merge launch serves them all)
  int32_t cand_q_first;            // index of this pass's first query inside the chunk
  // diagnostics
  int32_t debug_mode;              // 0 = product.  Timing-only experiments (results are garbage):
                                   // 1 = epilogue releases accumulators unread, 2 = TMEM reads but no
                                   // passage ends, 3 = mode 1 + MMA issue skipped (pure TMA streaming),
                                   // 4 = MMA issue never waits for the epilogue (epilogue idle),
                                   // 5 = mode 4 without the per-accumulator commit
                                   // 6 = product + clock64 timestamps of CTA 0's hand-offs -> dbg_ts
  int* status;                     // mapped pinned host word: watchdog code of a trapped launch
  long long* dbg_ts;               // [64 accumulators][8] timestamps (debug_mode 6), else null
};

constexpr uint32_t kDbgAcc0 = 2000;  // first accumulator index recorded in debug_mode 6
template <bool kDebug>
__device__ __forceinline__ void dbg_stamp(const ScanParams& p, int cta, uint32_t a, int slot) {
  if constexpr (kDebug) {
    if (p.debug_mode == 6 && cta == 0 && a >= kDbgAcc0 && a < kDbgAcc0 + 64)
      p.dbg_ts[(a - kDbgAcc0) * 8 + slot] = clock64();
  }
}

struct ScanSmem {
  static constexpr int kOffD = 0;
  static constexpr int kOffPartial = kOffD + kDStages * kDTileBytes;
  static constexpr int kPartialBytes = 2 * kRbMax * kSlots * 4;
  // per-lane passage maxima of the first kFastSlots passage ends of a tile: [2][kFastSlots][kRbMax][33]
  static constexpr int kOffLanePart = kOffPartial + kPartialBytes;
  static constexpr int kLanePartBytes = 2 * kFastSlots * kRbMax * kLaneStride * 4;
  static constexpr int kOffKeys = kOffLanePart + kLanePartBytes;
  static constexpr int kKeysBytes = kNqMax * kMaxK * 8;
  static constexpr int kOffMinKey = kOffKeys + kKeysBytes;           // u64[kNqMax]
  static constexpr int kOffMinPos = kOffMinKey + kNqMax * 8;         // int[kNqMax]
  static constexpr int kOffCarry = kOffMinPos + kNqMax * 4;          // float[kMtMax * 128]
  static constexpr int kOffBars = (kOffCarry + kMtMax * kTileM * 4 + 7) / 8 * 8;
  static constexpr int kNumBars = 1 + 2 * kDStages + 2 * kMaxAccStages + 4 + kMtMax * 4;
  static constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
  static constexpr int kBytes = kOffTmemPtr + 16 + 1024;  // + slack for 1024-B alignment
  static_assert(kBytes <= 232448, "exceeds 227 KiB of shared memory per CTA");
};

// TMEM column budget: query tiles first, accumulator stages in what is left.  The stage count is
// kept EVEN (4 or 2): accumulators alternate between two issuers / two epilogue warpgroups by index
// parity, so with an even count every stage (and its mbarrier pair) belongs to exactly one issuer
// and one warpgroup, which then see its phases strictly in order.  (With 3 stages an agent skips
// phases of a shared barrier and the parity wait can alias: observed as a hang at n_mtiles = 3.)
__host__ __device__ inline int scan_acc_stages(int n_mtiles) {
  return (512 - kQCols * n_mtiles) / kTileN >= 4 ? 4 : 2;
}

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
// Compact on purpose (few epilogue warps per SM sub-partition: instruction-cache misses, branches
// and dependent-issue latency are paid almost in full): 16 FMNMX for the 8 group maxima, then
// either a 4-instruction fold into the running max or, per passage END in this chunk, a selected
// tree max + one warp sum.
__device__ __forceinline__ void process_chunk(const uint32_t (&v)[32], uint32_t bits, float& m,
                                              float init, float* partial_rb, float* lane_part_rb,
                                              int& slot, int lane) {
  if (bits == 0u) {
    // no passage ends in this chunk: fold all 32 columns into the running max (16 FMNMX3)
    float r[11];
#pragma unroll
    for (int i = 0; i < 10; ++i)
      r[i] = fmax3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]),
                   __uint_as_float(v[3 * i + 2]));
    r[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
    const float a = fmax3(r[0], r[1], r[2]);
    const float b = fmax3(r[3], r[4], r[5]);
    const float c = fmax3(r[6], r[7], r[8]);
    const float d = fmax3(r[9], r[10], m);
    m = fmax3(a, b, fmaxf(c, d));
    return;
  }
  float gv[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float a = fmax3(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                          __uint_as_float(v[4 * g + 2]));
    gv[g] = fmaxf(a, __uint_as_float(v[4 * g + 3]));
  }
  uint32_t live = 0xFFu;  // groups not yet consumed by a finished passage
#pragma unroll 1
  while (bits) {          // warp-uniform: one iteration per passage ending in this chunk
    const uint32_t upto = bits ^ (bits - 1u);  // groups 0..(lowest set bit)
    const uint32_t seg = upto & live;
    float sel[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) sel[g] = ((seg >> g) & 1u) ? gv[g] : -INFINITY;
    const float x = fmax3(sel[0], sel[1], sel[2]);
    const float y = fmax3(sel[3], sel[4], sel[5]);
    float s = fmax3(m, x, y);
    s = fmax3(s, sel[6], sel[7]);
    if (slot < kFastSlots) {
      // common case: every lane parks its maximum, the reducer warp sums the 32 lanes later
      // (keeps the 5-step shuffle chain off the epilogue's critical path)
      lane_part_rb[slot * (kRbMax * kLaneStride)] = s;
    } else {
      s = warp_sum(s);
      if (lane == 0) partial_rb[slot] = s;
    }
    ++slot;
    m = init;
    live &= ~upto;
    bits &= bits - 1u;
  }
  {
    float sel[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) sel[g] = ((live >> g) & 1u) ? gv[g] : -INFINITY;
    const float x = fmax3(sel[0], sel[1], sel[2]);
    const float y = fmax3(sel[3], sel[4], sel[5]);
    m = fmax3(m, x, y);
    m = fmax3(m, sel[6], sel[7]);
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
// kDebug = false is the product instantiation: no timing-experiment branches, no timestamps in the
// hot loops (the epilogue is bound by instruction issue slots, every instruction there counts).
//
// kPair = true is the CTA-PAIR EXPERIMENT (launched as clusters of two CTAs): both CTAs of a pair stream the SAME
// token range, each with its own resident queries (twice the queries per corpus pass); every D tile is fetched from
// L2 / HBM once — CTA r loads rows [48 r, 48 r + 48) of the tile and TMA-multicasts them into both CTAs' shared
// memory — and a D stage is recycled when the issuers of BOTH CTAs have committed it (multicast tcgen05.commit).
// The MMAs stay cta_group::1: this isolates what a pair can save (half the HBM / L2 traffic per query) from what it
// cannot (the accumulator drain), see DESIGN.md 4.1.
template <bool kDebug, bool kPair = false>
__global__ void __launch_bounds__(kScanThreads, 1)
flmr_scan_kernel(const __grid_constant__ CUtensorMap tmap_d, const ScanParams p) {
  const int dbg = kDebug ? p.debug_mode : 0;
  using S = ScanSmem;
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment by pointer arithmetic on the __shared__ array (keeps the shared address space
  // visible to the compiler: LDS/STS instead of generic loads/stores)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);   // partition / list index
  const uint32_t pair_rank = kPair ? cluster_ctarank() : 0u;

  const uint32_t bar_base = smem_base + S::kOffBars;
  const uint32_t bar_q_full = bar_base;
  auto bar_d_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto bar_d_empty = [&](int s) { return bar_base + 8u * (1 + kDStages + s); };
  auto bar_t_full = [&](int s) { return bar_base + 8u * (1 + 2 * kDStages + s); };
  auto bar_t_empty = [&](int s) { return bar_base + 8u * (1 + 2 * kDStages + kMaxAccStages + s); };
  auto bar_p_full = [&](int b) { return bar_base + 8u * (1 + 2 * kDStages + 2 * kMaxAccStages + b); };
  auto bar_p_empty = [&](int b) {
    return bar_base + 8u * (1 + 2 * kDStages + 2 * kMaxAccStages + 2 + b);
  };
  // running-max hand-over between the epilogue warpgroups, one per (query tile, lane quadrant)
  auto bar_carry = [&](int mt, int quad) {
    return bar_base + 8u * (1 + 2 * kDStages + 2 * kMaxAccStages + 4 + mt * 4 + quad);
  };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + S::kOffTmemPtr);

  const int32_t row_begin = p.cta_row_begin[cta];
  const int64_t tile_base = p.cta_tile_base[cta];
  const int n_tiles = static_cast<int>(p.cta_tile_base[cta + 1] - tile_base);
  const int n_mtiles = p.n_mtiles;
  const uint32_t acc_stages = static_cast<uint32_t>(scan_acc_stages(n_mtiles));   // 2 or 4
  const uint32_t stage_mask = acc_stages - 1u, stage_shift = (acc_stages == 4u) ? 2u : 1u;
  const uint32_t acc_col0 = static_cast<uint32_t>(kQCols * n_mtiles);

  // ---- one-time setup --------------------------------------------------------------------------
  if (warp == kWarpProducer && lane == 0) {
    tma_prefetch_desc(&tmap_d);
    mbar_init(bar_q_full, 4);        // one arrive per query-staging warp
    for (int s = 0; s < kDStages; ++s) {
      mbar_init(bar_d_full(s), 1);
      mbar_init(bar_d_empty(s), kPair ? 2 * kMmaWarps : kMmaWarps);   // pair: the issuers of both CTAs
    }
    for (int s = 0; s < kMaxAccStages; ++s) {
      mbar_init(bar_t_full(s), 1);
      mbar_init(bar_t_empty(s), 4);  // one arrive per warp of the draining epilogue warpgroup
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_p_full(b), kEpiWarps);   // partial sums of a D tile complete
      mbar_init(bar_p_empty(b), kRedWarps);  // ... and consumed by the reducer warps
    }
    for (int i = 0; i < kMtMax * 4; ++i) mbar_init(bar_carry(i >> 2, i & 3), 1);
    mbar_fence_init();
  }
  if (warp == kWarpMma) {  // (the same warp deallocates at the end)
    tmem_alloc<512>(smem_base + S::kOffTmemPtr);
  }
  if (warp < kWarpProducer) {
    const int et = threadIdx.x;
    constexpr int kInitThreads = kEpiThreads + kRedWarps * 32;
    // top-k lists start empty (key 0 sorts below every real candidate)
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem + S::kOffKeys);
    for (int i = et; i < kNqMax * kMaxK; i += kInitThreads) keys[i] = 0ull;
    if (et < kNqMax) {
      reinterpret_cast<uint64_t*>(smem + S::kOffMinKey)[et] = 0ull;
      reinterpret_cast<int*>(smem + S::kOffMinPos)[et] = 0;
    }
    float* carry0 = reinterpret_cast<float*>(smem + S::kOffCarry);
    for (int i = et; i < kMtMax * kTileM; i += kInitThreads) carry0[i] = p.init_val;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if constexpr (kPair) cluster_sync_all();   // the peer's barriers exist before anything is multicast to them
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == kWarpProducer) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop (warp-uniform control flow keeps addresses in uniform
    // registers); one elected lane issues the bulk copies.
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % kDStages;
      const uint32_t ph = (t / kDStages) & 1;
      mbar_wait(bar_d_empty(s), ph ^ 1u, p.status, kDevTimeoutProducer);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(bar_d_full(s), kDTileBytes);   // (pair: 12 KB from this CTA's loads + 12 KB from the peer's)
        const uint32_t dst = smem_base + S::kOffD + s * kDTileBytes;
        const int32_t row = row_begin + t * kTileN;
        if constexpr (kPair) {   // tmap_d has a (kTileN / 2)-row box here
          const uint32_t half = pair_rank * (kTileN / 2);
          tma_load_2d_multicast(dst + half * 128, &tmap_d, bar_d_full(s), 0, row + half, 3, kPolicyEvictFirst);
          tma_load_2d_multicast(dst + kDKBlockBytes + half * 128, &tmap_d, bar_d_full(s), 64, row + half, 3,
                                kPolicyEvictFirst);
        } else {
          tma_load_2d(dst, &tmap_d, bar_d_full(s), 0, row, kPolicyEvictFirst);
          tma_load_2d(dst + kDKBlockBytes, &tmap_d, bar_d_full(s), 64, row, kPolicyEvictFirst);
        }
      }
      __syncwarp();
    }
  } else if (warp >= kWarpMma) {
    // ===================== MMA issuers =====================
    // A single thread needs ~31 cycles per tcgen05.mma issue plus ~300 cycles of wait / fence /
    // commit per accumulator -- more than the 384 cycles the 8 MMAs of an accumulator execute
    // (profiles/r01_handoff_timeline.md).  So TWO warps issue: issuer i owns the accumulators with
    // a = i (mod 2) in issue order (a = t * n_mtiles + j), i.e. with two TMEM stages each
    // issuer is bound to one stage and to the epilogue warpgroup that drains it.  Loops are
    // warp-uniform; one elected lane (always the same one) issues tcgen05.mma / commit so the
    // descriptors stay in uniform registers.
    const uint32_t iw = static_cast<uint32_t>(warp - kWarpMma);
    constexpr uint32_t idesc = make_idesc_bf16_f32(kTileM, kTileN);
    mbar_wait(bar_q_full, 0, p.status, kDevTimeoutMma);  // queries are in TMEM
    tc_fence_after_sync();
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % kDStages;
      const uint32_t ph = (t / kDStages) & 1;
      mbar_wait(bar_d_full(s), ph, p.status, kDevTimeoutMma);
      tc_fence_after_sync();
      const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_base + S::kOffD + s * kDTileBytes);
      const uint32_t a_first = static_cast<uint32_t>(t) * n_mtiles;
      const bool rot = ((n_mtiles & t) & 1) != 0;  // odd tile, odd tile count: rotated order (see epilogue)
#pragma unroll 1
      for (uint32_t a = a_first + ((a_first ^ iw) & 1u); a < a_first + n_mtiles; a += 2) {
        const uint32_t j = a - a_first;
        const uint32_t mt = rot ? (j == 0 ? static_cast<uint32_t>(n_mtiles) - 1u : j - 1u) : j;
        const uint32_t as = a & stage_mask, aph = (a >> stage_shift) & 1u;  // TMEM stage / phase
        if (dbg != 4 && dbg != 5) {  // (modes 4/5: never wait for the epilogue)
          mbar_wait(bar_t_empty(as), aph ^ 1u, p.status, kDevTimeoutMma);
          tc_fence_after_sync();
        }
        const uint32_t d_tmem = tmem_base + acc_col0 + as * kTileN;
        const uint32_t a_tmem = tmem_base + mt * kQCols;
        if (elect_one_sync()) {
          dbg_stamp<kDebug>(p, cta, a, 0);  // stage free, about to issue
          if (dbg != 3) {
#pragma unroll
            for (int k = 0; k < kDim / 16; ++k) {
              // advance the start-address field (16-byte units) inside the descriptor
              const uint64_t b_desc =
                  b_desc0 + static_cast<uint64_t>(((k >> 2) * kDKBlockBytes + (k & 3) * 32) >> 4);
              tc_mma_ts(d_tmem, a_tmem + k * 8, b_desc, idesc, k > 0 ? 1u : 0u);
            }
          }
          if (dbg != 5) tc_commit(bar_t_full(as));  // accumulator complete -> epilogue
          dbg_stamp<kDebug>(p, cta, a, 1);  // MMAs + commit issued
        }
        __syncwarp();
      }
      // this issuer's MMAs on the D stage are complete -> producer (both issuers must arrive)
      if (elect_one_sync()) {
        if constexpr (kPair) tc_commit_multicast(bar_d_empty(s), 3);
        else tc_commit(bar_d_empty(s));
      }
      __syncwarp();
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue (TMEM drain) =====================
    // Accumulators are numbered in MMA issue order, a = t * n_mtiles + j; warpgroup g drains the
    // accumulators with (a & 1) == g, so the two warpgroups alternate strictly.  The running max of
    // the passage straddling D tiles is per (query tile, row) and must follow the query tile from D
    // tile to D tile.  With an even tile count query tile mt = j always meets the same warpgroup.
    // With an odd count the issue order of odd D tiles is rotated (j = 0 is the LAST query tile, j > 0
    // is tile j - 1): query tiles 0 .. n_mtiles-2 then keep their warpgroup and only the last one
    // changes hands every D tile (shared memory + mbarrier), instead of all of them.
    const int wg = warp >> 2;         // epilogue warp 0..7 -> warpgroup 0/1
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    float* partial = reinterpret_cast<float*>(smem + S::kOffPartial);
    float* lane_part = reinterpret_cast<float*>(smem + S::kOffLanePart) + lane;
    float* carry = reinterpret_cast<float*>(smem + S::kOffCarry) + quad * 32 + lane;
    const bool carry_crosses = (n_mtiles & 1) != 0;  // odd: (t, mt) and (t+1, mt) are drained by different warpgroups
    const float init = p.init_val;

    // ---- stage the resident queries into tensor memory (warpgroup 0: one warp per lane quadrant).
    // Row r of query tile mt lives in TMEM lane r, columns [mt*64, mt*64+64): column c holds the
    // bf16 pair (k = 2c, 2c+1), i.e. the row's 256 bytes verbatim.
    if (wg == 0) {
      for (int mt = 0; mt < n_mtiles; ++mt) {
        const uint4* src = p.q_pad + (static_cast<int64_t>(pair_rank) * (kMtMax * kTileM) +   // pair: CTA r's staging slot
                                      static_cast<int64_t>(mt) * kTileM + quad * 32 + lane) * 16;
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

    const int n_tiles_epi = (dbg == 4 || dbg == 5) ? 0 : n_tiles;  // modes 4/5: epilogue idle
    uint32_t mask_next = (n_tiles_epi > 0) ? __ldg(p.tile_end_mask + tile_base) : 0u;
    for (int t = 0; t < n_tiles_epi; ++t) {
      const uint32_t mask = (dbg == 2) ? 0u : mask_next;
      if (t + 1 < n_tiles_epi) mask_next = __ldg(p.tile_end_mask + tile_base + t + 1);
      const int buf = t & 1;
      // the reducers must have consumed the partial sums of tile t-2 before this buffer is reused
      mbar_wait(bar_p_empty(buf), ((static_cast<uint32_t>(t) >> 1) & 1u) ^ 1u, p.status,
                kDevTimeoutEpilogue);
      const uint32_t a_first = static_cast<uint32_t>(t) * n_mtiles;
      const bool rot = ((n_mtiles & t) & 1) != 0;
#pragma unroll 1
      for (uint32_t a = a_first + ((a_first ^ static_cast<uint32_t>(wg)) & 1u); a < a_first + n_mtiles;
           a += 2) {
        const int j = static_cast<int>(a - a_first);
        const int mt = rot ? (j == 0 ? n_mtiles - 1 : j - 1) : j;
        const bool crosses = carry_crosses && mt == n_mtiles - 1;
        const uint32_t as = a & stage_mask, aph = (a >> stage_shift) & 1u;  // TMEM stage / phase
        // running max handed over by whoever drained (t-1, mt): with an odd number of query tiles
        // that is the other warpgroup (mbarrier arrive/wait = release/acquire); with an even number
        // it is this very warp, and program order suffices
        if (crosses && t > 0)
          mbar_wait(bar_carry(mt, quad), static_cast<uint32_t>(t - 1) & 1u, p.status,
                    kDevTimeoutEpilogue);
        float m = carry[mt * kTileM];
        if (quad == 0 && lane == 0) dbg_stamp<kDebug>(p, cta, a, 2);               // epilogue ready to wait
        mbar_wait(bar_t_full(as), aph, p.status, kDevTimeoutEpilogue);
        tc_fence_after_sync();
        if (quad == 0 && lane == 0) dbg_stamp<kDebug>(p, cta, a, 3);               // accumulator visible
        if (dbg == 1 || dbg == 3) {  // timing experiment: release unread
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_t_empty(as));
        } else {
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
          // every column is in registers: hand the TMEM stage back to the MMA warps
          if (quad == 0 && lane == 0) dbg_stamp<kDebug>(p, cta, a, 4);             // TMEM read done
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_t_empty(as));
          if (quad == 0 && lane == 0) dbg_stamp<kDebug>(p, cta, a, 5);             // stage handed back
#pragma unroll
          for (int c = 1; c < kChunks; ++c)
            process_chunk(v[c], (mask >> (8 * c)) & 0xFFu, m, init, partial_rb, lane_part_rb, slot, lane);
        }
        if (quad == 0 && lane == 0) dbg_stamp<kDebug>(p, cta, a, 6);               // chunk processing done
        carry[mt * kTileM] = m;
        if (crosses) {
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_carry(mt, quad));
        }
      }
      // this warp's per-lane maxima / partial sums of tile t are written: order every lane's stores
      // before lane 0's arrive (release); the reducer's wait is the matching acquire
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_full(buf));
    }
    tc_fence_before_sync();
  } else {
    // ===================== reducers: score finalisation + per-CTA top-k =====================
    // Per D tile and per (query, passage ending in the tile): sum the row-block partials in fixed
    // order (deterministic), add/store partial scores if requested, offer to the top-k list.
    const int rw = warp - kWarpRed0;  // reducer 0..kRedWarps-1 owns queries b = rw (mod kRedWarps)
    const int q_off = static_cast<int>(pair_rank) * p.nq_pass;   // pair: CTA r's queries follow CTA 0's in every per-query array
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
              const int64_t gi = static_cast<int64_t>(q_off + b) * p.n_passages + pid;
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
            const int64_t gi = static_cast<int64_t>(q_off + b) * p.n_passages + pid;
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
        uint64_t* dst = p.cand_keys + (static_cast<int64_t>(cta) * p.cand_q_stride + p.cand_q_first + q_off + b) * p.k;
        for (int i = lane; i < p.k; i += 32) dst[i] = keys[b * kMaxK + i];
      }
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();   // no CTA leaves while its peer may still signal its barriers
  if (warp == kWarpMma) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace flmr
