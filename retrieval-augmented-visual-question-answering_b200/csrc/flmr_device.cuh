// flmr_device.cuh — thin inline-PTX wrappers for the sm_100a features the scan kernel uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences).
// Hand-written for this repository; sm_100a only.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace flmr {

// ---- error codes written to ScanParams::status by the device-side watchdog -------------------
enum : int {
  kDevOk = 0,
  kDevTimeoutProducer = 101,
  kDevTimeoutMma = 102,
  kDevTimeoutEpilogue = 103,
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of the (fully converged) warp; the same lane every time, so tcgen05.commit, which tracks
// the MMAs "initiated by the executing thread", sees all of them.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or the
// hint expires, instead of returning (and re-issuing the polling loop) after the short default limit.
// Without the hint ~20 % of all executed warp instructions of the scan kernel were polling iterations.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)   // up to 10 ms per attempt
      : "memory");
  return ok != 0;
}

// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU box.
// `status` receives `code` before the trap so the host can say which role starved.
//
// The polling loop is written in PTX so that one iteration is the try_wait itself plus four instructions: the
// round-1 C++ loop compiled to 15 instructions per iteration, and with waits this fine-grained (a warp is woken
// by every partial arrival on its barrier) the polling iterations were 25 % of all warp instructions the scan
// kernel executed (profiles/r01 source page).  The watchdog reads the timer once per kSpinBlock failed polls.
constexpr uint32_t kSpinBlock = 4096;
__device__ __forceinline__ uint32_t mbar_poll_block(uint32_t bar, uint32_t parity) {
  uint32_t n;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
      "mov.u32 c, 0;\n"
      "FLMR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "@p bra FLMR_DONE;\n\t"
      "add.u32 c, c, 1;\n\t"
      "setp.lt.u32 p, c, %4;\n\t"
      "@p bra FLMR_WAIT;\n"
      "FLMR_DONE:\n\t"
      "mov.u32 %0, c;\n\t}"
      : "=r"(n)
      : "r"(bar), "r"(parity), "r"(0x989680u), "r"(kSpinBlock)
      : "memory");
  return n;   // < kSpinBlock: the phase completed
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int* status, int code) {
  // 20 s of polling = a protocol bug, not a slow wait (time slicing, a sanitizer or a debugger can stretch a
  // legitimate wait to seconds; the longest real wait of a scan is one tile, microseconds)
  constexpr uint64_t timeout_ns = 20000000000ull;
  uint64_t t0 = 0;
  while (mbar_poll_block(bar, parity) >= kSpinBlock) {
    const uint64_t now = global_timer_ns();
    if (t0 == 0) {
      t0 = now;
    } else if (now - t0 > timeout_ns) {
      if (status) *reinterpret_cast<volatile int*>(status) = code;
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* status, int code) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, status, code);
}

// ---- TMA ----------------------------------------------------------------------------------------
constexpr uint64_t kPolicyEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kPolicyEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kPolicyEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// 2-D tiled load global -> shared, completion on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, uint32_t bar,
                                            int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1),
      "l"(policy)
      : "memory");
}

// Same, delivered to the same shared-memory offset (and signalled on the same barrier offset) of EVERY CTA of the
// cluster whose bit is set in `cta_mask`.
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst_smem, const CUtensorMap* map, uint32_t bar,
                                                      int32_t c0, int32_t c1, uint16_t cta_mask, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5, %6;"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask),
      "l"(policy)
      : "memory");
}

// ---- thread-block cluster ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// tcgen05.commit: arrive once on `bar` when all MMAs issued so far by this thread have completed.
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}

// tcgen05.commit that arrives on the barrier at the same offset in every CTA of `cta_mask`.
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask)
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 in / fp32 accumulate, one CTA.
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (128 rows x 16 bf16 per MMA = 8 TMEM columns,
// row = lane, two consecutive K elements per 32-bit column) is read from tensor memory.
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle, rows of 128 B
// (64 bf16), 8-row groups 1024 B apart.  Field layout: start>>4 [0,14) | LBO>>4 [16,30) |
// SBO>>4 [32,46) | version=1 [46,48) | layout type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;             // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;     // SBO: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;             // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major.
// c_format [4,6)=1 (f32) | a_format [7,10)=1 (bf16) | b_format [10,13)=1 | N>>3 [17,23) | M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// 32 lanes x 32 consecutive fp32 columns of the accumulator -> 32 registers per thread
// (thread t of the warp reads TMEM lane (lane-quadrant base + t)).
#define FLMR_TMEM_LD32(v, taddr)                                                                   \
  asm volatile(                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                    \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                    \
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                   \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),        \
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),    \
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), \
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), \
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                         \
      : "r"(taddr)                                                                                 \
      : "memory")

// tcgen05.wait::ld with the destination registers threaded through as "+r" operands so the
// compiler cannot hoist their uses above the wait.
#define FLMR_TMEM_WAIT_LD32(v)                                                                     \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                    \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]),           \
                 "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]),         \
                 "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]),     \
                 "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),     \
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),     \
                 "+r"(v[30]), "+r"(v[31])::"memory")

// 32 registers per thread -> 32 lanes x 32 consecutive 32-bit TMEM columns.
#define FLMR_TMEM_ST32(taddr, v)                                                                   \
  asm volatile(                                                                                    \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                              \
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"                                   \
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                          \
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),   \
      "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),             \
      "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),          \
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),          \
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])                                               \
      : "memory")

__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Order-preserving map float -> uint32 (larger float <-> larger uint).
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t u) {
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

}  // namespace flmr
