// Inline-PTX helpers shared by the tensor-core convolution kernels (mbarrier, cp.async, bulk/TMA copies, tcgen05).
#pragma once
#include <cstdio>
#include "common.cuh"

namespace tsg {

constexpr int TC_BM = 128;
constexpr int TC_KB = 64;                       // channels per unit (128 B of bf16)
constexpr int TC_A_BYTES = TC_BM * TC_KB * 2;   // 16 KB
constexpr int TC_EPI_WARPS = 4;

#ifdef TSG_TC_TRACE
__device__ int g_dbg[160][32][4];  // debugging build: per (block, warp) role state, printed by the hang watchdog
#define TSG_STATE(a, b, c, d)                                                         \
  do {                                                                                \
    if ((threadIdx.x & 31) == 0) {                                                    \
      int *q_ = g_dbg[blockIdx.x % 160][threadIdx.x >> 5];                            \
      q_[0] = (a); q_[1] = (b); q_[2] = (c); q_[3] = (d);                             \
    }                                                                                 \
  } while (0)
#else
#define TSG_STATE(a, b, c, d) do { } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Spin on an mbarrier phase.  try_wait suspends the thread in hardware for a bounded time, so the loop is not hot.  A protocol
// bug must fail loudly, never hang the GPU: the production build counts failed polls (2^26 of them are seconds) and traps;
// the trace build reads the SM cycle counter and prints who hangs on what first.  (The watchdog used to be a clock64() /
// 64-bit compare sequence inlined at each of the ~20 wait sites: 22 % of the kernel's SASS, and code size costs this kernel
// measurable time — profiles/README.md.)
#ifndef TSG_TC_TRACE
template <bool CLUSTER = false>   // CLUSTER marks the waits whose arrivals come from the other CTA of a pair (same instruction:
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {   // see the note in the trace variant below)
  uint32_t done = 0, spin = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done && ++spin < (1u << 26));
  if (!done) __trap();
}
#else
// CLUSTER marks the waits whose arrivals come from the other CTA of a pair.  They use the same CTA-scope acquire as
// the local ones (as CUTLASS's ClusterBarrier does): what they order is shared memory that has physically landed and TMEM
// reads bracketed by tcgen05 fences; the cluster-scope forms compile to MEMBAR.ALL.GPU / CCTL.IVALL per use, which made
// the first pair kernel 2.2x slower than the single-CTA one.
template <bool CLUSTER = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 1;; ++spin) {
    if (CLUSTER)
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
    if (done) break;
    if ((spin & 0xfffu) == 0) {
      const long long now = clock64();
      if (!t0) t0 = now;
#ifdef TSG_TC_TRACE
      else if (now - t0 > (1ll << 30)) {  // debugging build: say who hangs on what, give the printf time to drain, then trap
        if ((threadIdx.x & 31) == 0) {
          const int *q_ = g_dbg[blockIdx.x % 160][threadIdx.x >> 5];
          unsigned long long w_[8];
          for (int z_ = 0; z_ < 8; ++z_) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(w_[z_]) : "r"(1024u + 8u * z_));
          printf("HANG block %d warp %d bar %d parity %u state %d %d %d %d | full %llx %llx %llx %llx %llx | me %llx\n", blockIdx.x, threadIdx.x >> 5,
                 (int)(bar - 1024) / 8, parity, q_[0], q_[1], q_[2], q_[3], w_[0], w_[1], w_[2], w_[3], w_[4], w_[((bar - 1024) / 8) & 7]);
        }
        __nanosleep(1000000);
        if (now - t0 > (1ll << 31)) __trap();
      }
#else
      else if (now - t0 > (1ll << 33)) __trap();
#endif
    }
  }
}
#endif
// one non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` once every cp.async issued so far by this thread has landed (counts against the expected arrivals)
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// ---- CTA-pair (cta_group::2) helpers
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(cta)
      : "memory");
}
// "all MMAs issued so far by this thread have completed" -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// M = 256 over the pair: descriptors address the same shared-memory offsets in both CTAs; each CTA holds N/2 rows of B
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// One K stream accumulated into an output tile: the kernel offsets of a convolution over (in0 | in1), or a folded 1x1x1
// shortcut convolution of a residual block (K = 1, its index line = the centre offset of the block's kernel map).
struct TcPhase {
  const __nv_bfloat16 *in0, *in1;
  const uint8_t *packed_w;
  const int *nbr;                 // NULL = identity map (tile row r reads row r; only without perm)
  long long nbr_stride;           // elements between offsets of nbr; multiple of 256, padding rows hold -1
  unsigned long long slice_need;  // 4 bits per slice j < Q: which of the P offsets of a virtual offset slice j touches
  int c0, c1;
  int pk, kq, cpo;                // K-slice packing: offsets per virtual offset (P), slices per virtual offset (Q), chunks per offset
  int K;
};
struct TcParams {
  TcPhase ph[2];
  int n_phases;
  int c_out;   // row pitch of out / residual / bias
  int n_eff;   // accumulator width of one work item: c_out / ns
  int ns;      // N split: work item = (super tile, column block of n_eff channels)
  int na;      // pipeline stages
  int ksmax;   // most kernel offsets one K slice touches, over both phases (sizes the producers' index buffers)
  const unsigned *tile_mask;
  const int *perm;  // tile row r -> output row (NULL = identity); nbr / tile_mask are indexed by tile row
  long long n_out;        // output rows (capacity when n_out_dev is set)
  const int *n_out_dev;   // sync-free pipeline: the row count lives on the device
  void *out;
  int out_f32;
  const float *bias;
  const __nv_bfloat16 *residual;
  int relu;
  uint32_t tmem_cols;
  int dbg;     // profiling knock-out bits (TSG_TC_DEBUG), 0 in production
  int *sched;  // {next ticket, finished CTAs}, zero between launches; NULL = static round-robin tile assignment
  // K split (tsg_conv_split_items): work items {tile, offset mask, part | parts << 8, partial-sum slot} replace the tile
  // enumeration; tiles with many active offsets are summed by two items (on two SMs) and combined in the epilogue
  const int4 *items;
  const int *n_items;    // device: number of work items
  float *split_scratch;  // slots x split_parts slabs of (n_eff / 4) x 128 float4 partial sums
  int *split_state;      // slots x 8 epilogue warps x {arrivals, parked parts}, zero between launches
  int split_parts;       // slabs per slot (most parts a tile is split into)
  int lookahead;         // planner throttle: a ticket is drawn when fewer than lookahead * stages + 4 planned stages are left
};

struct Ring {  // slot + phase of a circular mbarrier pipeline
  uint32_t slot = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t n) {
    if (++slot == n) {
      slot = 0;
      phase ^= 1;
    }
  }
};

__device__ __forceinline__ int next_bit(unsigned mask, int after) {  // first set bit strictly above `after`, or 32
  const unsigned m = after >= 31 ? 0u : (mask & (0xffffffffu << (after + 1)));
  return m ? __ffs(m) - 1 : 32;
}

// Long waits (an epilogue warp waiting for a whole mainloop): back off between polls so the idle warp does not
// compete for issue slots with the producer / MMA warps that share its scheduler.
#ifndef TSG_TC_TRACE
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spin = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(200);
    if (++spin == (1u << 24)) __trap();   // > 3 s of back-off: a protocol bug
  }
}
#else
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 1;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(200);
    if ((spin & 0x3ffu) == 0) {
      const long long now = clock64();
      if (!t0) t0 = now;
#ifdef TSG_TC_TRACE
      else if (now - t0 > (1ll << 30)) {
        if ((threadIdx.x & 31) == 0) printf("HANG block %d warp %d bar %d parity %u (sleep)\n", blockIdx.x, threadIdx.x >> 5, (int)(bar - 1024) / 8, parity);
        __nanosleep(1000000);
        if (now - t0 > (1ll << 31)) __trap();
      }
#else
      else if (now - t0 > (1ll << 33)) __trap();
#endif
    }
  }
}

#endif

}  // namespace tsg
