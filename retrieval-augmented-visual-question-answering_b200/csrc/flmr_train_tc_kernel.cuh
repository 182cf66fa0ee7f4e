// flmr_train_tc_kernel.cuh — the training / re-scoring forward on the tcgen05 tensor cores:
//
//   arg[b, p, i]    = argmax_{j : mask[p, j]} <Q[b, i, :], D[p, j, :]>        (lowest j on ties, -1 if none)
//   rowmax[b, p, i] = that maximum;   sum_i rowmax = MaxSim score[b, p]
//
// Same contract as flmr_argmax_mma_kernel (flmr_train_kernels.cuh), which stays the path for small batches;
// this one is for batches whose contraction is worth a TMA/TMEM pipeline — the global in-batch-negatives matrix
// of a training step with cross-rank negatives (CB/modeling/colbert.py:64-113, 115-163: 8 queries x 832 rows
// against 128 documents x 512 tokens per rank) or an exhaustive evaluation through ColBERT.score
// (src/executors/FLMR_executor.py:826-833).
//
// Two launches:
//   1. flmr_compact_docs_kernel: every document's unmasked tokens (ColBERT.doc masks padding AND punctuation,
//      colbert.py:199-210: holes anywhere) are packed to the front of a 128-token-aligned slot, the slot's tail
//      filled with copies of the last unmasked token (a duplicate cannot change a max and never wins the
//      lowest-index tie-break); `idx_map` remembers each packed token's position in the padded document.
//      After this no mask exists any more: the epilogue below has no per-element mask logic.
//   2. flmr_argmax_tc_kernel: CTA = (query b, 128-row tile of its tokens, a range of documents).  The query tile
//      is the stationary A operand (TMA -> shared memory once); the CTA streams its documents in chunks of 128
//      tokens through a 4-stage ring.  Stage s = a 32 KB shared-memory chunk AND a 128-column TMEM accumulator:
//        warp 8 (producer): TMA, two 128B-swizzled boxes per chunk
//        warp 9 (issuer)  : 8 x tcgen05.mma (M = 128, N = 128, K = 16), tcgen05.commit -> `done[s]`
//        warps 0-7        : two epilogue warpgroups, documents alternate between them (a document's chunks stay
//                           with one warpgroup, so its running (max, index) never changes hands).  TMEM lane =
//                           query row: per 32 columns a compare-select TREE (strict '>' keeps the lower index on
//                           ties) — 31 independent (FSETP, FSEL, SEL) triples of depth 5 instead of a 32-long
//                           dependent chain, which with one warp per scheduler ran at a fifth of the MMA rate —
//                           then one compare against the row's running best; one coalesced store per output at
//                           the document's end.
//      Per chunk the issuer commits twice: `free[s]` returns the shared-memory chunk to the producer, `done[g][s]`
//      hands the accumulator to the warpgroup g that owns the document (one barrier per warpgroup and stage, so
//      every waiter observes every phase of its barrier); `t_empty[s]` (one arrive per warp of the owner) hands
//      the accumulator back to the issuer.
#pragma once
#include "flmr_device.cuh"

namespace flmr {

constexpr int kTcTile = 128;                 // query rows per CTA and document tokens per chunk
constexpr int kTcStages = 4;                 // smem chunks == TMEM accumulator stages (4 x 128 columns)
constexpr int kTcChunkBytes = kTcTile * 128 * 2;   // 32 KiB: [2 k-blocks][128 rows][64 bf16]
constexpr int kTcKBlockBytes = kTcTile * 128;      // 16 KiB
constexpr int kTcEpiWarps = 8;               // two epilogue warpgroups; warpgroup g owns the CTA's documents of parity g
constexpr int kTcWarpProducer = kTcEpiWarps;
constexpr int kTcWarpMma = kTcEpiWarps + 1;
constexpr int kTcThreads = (kTcEpiWarps + 2) * 32;
constexpr int kTcSmemBytes = (1 + kTcStages) * kTcChunkBytes + 512 + 1024;   // A tile + ring + barriers, lengths + align

// grid = total documents; one 256-thread block packs one document.
__global__ void __launch_bounds__(256)
flmr_compact_docs_kernel(const uint4* __restrict__ d, const uint8_t* __restrict__ mask, int nd, int nd_c,
                         uint4* __restrict__ dc, int32_t* __restrict__ idx_map, int32_t* __restrict__ len_out) {
  __shared__ int s_cnt[256];
  __shared__ int s_last;
  const int p = blockIdx.x, tid = threadIdx.x;
  const uint8_t* mp = mask + static_cast<int64_t>(p) * nd;
  // each thread owns a contiguous slice of the document's tokens: exclusive scan of the per-slice counts
  const int per = (nd + 255) / 256;
  const int j0 = tid * per, j1 = min(nd, j0 + per);
  int cnt = 0;
  for (int j = j0; j < j1; ++j) cnt += mp[j] != 0;
  s_cnt[tid] = cnt;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    const int v = tid >= off ? s_cnt[tid - off] : 0;
    __syncthreads();
    s_cnt[tid] += v;
    __syncthreads();
  }
  const int len = s_cnt[255];
  int c = s_cnt[tid] - cnt;
  int32_t* im = idx_map + static_cast<int64_t>(p) * nd_c;
  for (int j = j0; j < j1; ++j)
    if (mp[j]) im[c++] = j;
  if (tid == 0) len_out[p] = len;
  __syncthreads();
  // copy rows: 16 threads per row (16 B each), packed position -> source token through idx_map
  const int len_pad = (len + kTcTile - 1) / kTcTile * kTcTile;
  const uint4* src = d + static_cast<int64_t>(p) * nd * 16;
  uint4* dst = dc + static_cast<int64_t>(p) * nd_c * 16;
  if (tid == 0) s_last = len > 0 ? im[len - 1] : 0;
  __syncthreads();
  for (int r = tid >> 4; r < len_pad; r += 16) {
    const int t = r < len ? im[r] : s_last;
    dst[r * 16 + (tid & 15)] = src[static_cast<int64_t>(t) * 16 + (tid & 15)];
  }
}

struct ArgmaxTcParams {
  const int32_t* doc_len;      // [n_docs_total] unmasked tokens per document
  const int32_t* idx_map;      // [n_docs_total][nd_c] packed position -> token index in the padded document
  int32_t* arg;                // [B][n_per][nq]
  float* rowmax;               // same shape, or null
  int32_t nq, nd_c, n_per, stride_b, docs_per_cta;
  int* status;                 // watchdog word (may be null)
};

__global__ void __launch_bounds__(kTcThreads, 1)
flmr_argmax_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                      const ArgmaxTcParams p) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_smem = smem_base;                                   // query tile
  const uint32_t d_smem0 = smem_base + kTcChunkBytes;                  // ring of document chunks
  const uint32_t bar_base = smem_base + (1 + kTcStages) * kTcChunkBytes;
  const uint32_t bar_q = bar_base;
  auto bar_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto bar_free = [&](int s) { return bar_base + 8u * (1 + kTcStages + s); };          // MMA -> producer
  auto bar_tempty = [&](int s) { return bar_base + 8u * (1 + 2 * kTcStages + s); };    // owner warpgroup -> MMA
  // accumulator ready, one barrier per (warpgroup, stage): a parity wait only tells CONSECUTIVE phases apart, and
  // a warpgroup sees only its own documents' chunks, so it needs a barrier whose every phase it observes
  auto bar_done = [&](int g, int s) { return bar_base + 8u * (1 + 3 * kTcStages + g * kTcStages + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + (1 + kTcStages) * kTcChunkBytes + 8 * (1 + 5 * kTcStages));
  // lengths of this CTA's documents, read from global memory ONCE (every role walks the same document list; a
  // global load per document and role put ~500 cycles of latency in front of each document's first chunk)
  int32_t* s_len = reinterpret_cast<int32_t*>(smem + (1 + kTcStages) * kTcChunkBytes + 8 * (1 + 5 * kTcStages) + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, mt = blockIdx.y;
  const int p_begin = blockIdx.x * p.docs_per_cta;
  const int p_end = min(p.n_per, p_begin + p.docs_per_cta);

  if (warp == kTcWarpProducer && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_d);
    mbar_init(bar_q, 1);
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_free(s), 1);
      mbar_init(bar_done(0, s), 1);
      mbar_init(bar_done(1, s), 1);
      mbar_init(bar_tempty(s), 4);
    }
    mbar_fence_init();
  }
  if (warp == kTcWarpMma) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
  if (warp == 0 && p_begin + lane < p_end)
    s_len[lane] = __ldg(p.doc_len + static_cast<int64_t>(b) * p.stride_b + p_begin + lane);   // docs_per_cta <= 32
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == kTcWarpProducer) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(bar_q, kTcChunkBytes);
      const int32_t row = b * p.nq + mt * kTcTile;
      tma_load_2d(a_smem, &tmap_q, bar_q, 0, row, kPolicyEvictNormal);
      tma_load_2d(a_smem + kTcKBlockBytes, &tmap_q, bar_q, 64, row, kPolicyEvictNormal);
    }
    __syncwarp();
    uint32_t n = 0;                                 // chunks issued so far
    for (int pl = p_begin; pl < p_end; ++pl) {
      const int64_t pg = static_cast<int64_t>(b) * p.stride_b + pl;
      const int nch = (s_len[pl - p_begin] + kTcTile - 1) / kTcTile;
      for (int c = 0; c < nch; ++c, ++n) {
        const int s = n % kTcStages;
        const uint32_t use = n / kTcStages;         // how often this stage was used before
        if (use > 0) mbar_wait(bar_free(s), (use - 1) & 1u, p.status, kDevTimeoutProducer);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar_full(s), kTcChunkBytes);
          const uint32_t dst = d_smem0 + s * kTcChunkBytes;
          const int32_t row = static_cast<int32_t>(pg * p.nd_c + c * kTcTile);
          tma_load_2d(dst, &tmap_d, bar_full(s), 0, row, kPolicyEvictNormal);
          tma_load_2d(dst + kTcKBlockBytes, &tmap_d, bar_full(s), 64, row, kPolicyEvictNormal);
        }
        __syncwarp();
      }
    }
  } else if (warp == kTcWarpMma) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16_f32(kTcTile, kTcTile);
    mbar_wait(bar_q, 0, p.status, kDevTimeoutMma);
    tc_fence_after_sync();
    const uint64_t a_desc0 = make_kmajor_sw128_desc(a_smem);
    uint32_t n = 0;
    for (int pl = p_begin; pl < p_end; ++pl) {
      const int nch = (s_len[pl - p_begin] + kTcTile - 1) / kTcTile;
      for (int c = 0; c < nch; ++c, ++n) {
        const int s = n % kTcStages;
        const uint32_t use = n / kTcStages;
        mbar_wait(bar_full(s), use & 1u, p.status, kDevTimeoutMma);
        if (use > 0) mbar_wait(bar_tempty(s), (use - 1) & 1u, p.status, kDevTimeoutMma);
        tc_fence_after_sync();
        const uint64_t b_desc0 = make_kmajor_sw128_desc(d_smem0 + s * kTcChunkBytes);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t koff = static_cast<uint64_t>(((k >> 2) * kTcKBlockBytes + (k & 3) * 32) >> 4);
            tc_mma_ss(tmem_base + s * kTcTile, a_desc0 + koff, b_desc0 + koff, idesc, k > 0 ? 1u : 0u);
          }
          tc_commit(bar_free(s));                               // shared-memory chunk consumed -> producer
          tc_commit(bar_done((pl - p_begin) & 1, s));           // accumulator complete -> the document's warpgroup
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: running (max, index) per query row =====================
    const int wg = warp >> 2, quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const int i = mt * kTcTile + quad * 32 + lane;               // this thread's query token
    const bool live = i < p.nq;
    uint32_t n = 0;
    uint32_t own_parity = 0;                                     // bit s: parity of this warpgroup's next use of stage s
    for (int pl = p_begin; pl < p_end; ++pl) {
      const int64_t pg = static_cast<int64_t>(b) * p.stride_b + pl;
      const int len = s_len[pl - p_begin];
      const int nch = (len + kTcTile - 1) / kTcTile;
      if (((pl - p_begin) & 1) != wg) {
        n += nch;                                                // the other warpgroup's document
        continue;
      }
      float best = -INFINITY;
      int barg = -1;
      for (int c = 0; c < nch; ++c, ++n) {
        const int s = n % kTcStages;
        mbar_wait(bar_done(wg, s), (own_parity >> s) & 1u, p.status, kDevTimeoutEpilogue);
        own_parity ^= 1u << s;
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + lane_base + s * kTcTile;
        uint32_t v[2][32];
        FLMR_TMEM_LD32(v[0], taddr);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          FLMR_TMEM_WAIT_LD32(v[q & 1]);
          if (q < 3) {
            FLMR_TMEM_LD32(v[(q + 1) & 1], taddr + 32 * (q + 1));
          } else {                                               // every column is in registers: stage back to the issuer
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty(s));
          }
          // compare-select tree over the 32 columns: (value, column) pairs, the lower column wins ties
          float tv[16];
          int ti[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float x0 = __uint_as_float(v[q & 1][2 * e]), x1 = __uint_as_float(v[q & 1][2 * e + 1]);
            const bool g1 = x1 > x0;
            tv[e] = g1 ? x1 : x0;
            ti[e] = g1 ? 2 * e + 1 : 2 * e;
          }
#pragma unroll
          for (int w = 8; w >= 1; w >>= 1) {
#pragma unroll
            for (int e = 0; e < w; ++e) {
              const bool g1 = tv[2 * e + 1] > tv[2 * e];
              tv[e] = g1 ? tv[2 * e + 1] : tv[2 * e];
              ti[e] = g1 ? ti[2 * e + 1] : ti[2 * e];
            }
          }
          if (tv[0] > best) {                                    // strict: earlier chunks / columns keep ties
            best = tv[0];
            barg = c * kTcTile + q * 32 + ti[0];
          }
        }
      }
      if (live) {
        const int64_t o = (static_cast<int64_t>(b) * p.n_per + pl) * p.nq + i;
        p.arg[o] = barg >= 0 ? __ldg(p.idx_map + pg * p.nd_c + barg) : -1;
        if (p.rowmax) p.rowmax[o] = best;
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == kTcWarpMma) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace flmr
