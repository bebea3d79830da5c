// Sparse convolution on the 5th-generation tensor cores (sm_100a): output-stationary implicit GEMM,
//   out[row(r),:] = epilogue( sum_k in[nbr[k,r],:] @ W[k] ),   bf16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM walks "super tiles" of G x 128 tile rows (G in {1,2}), handed out by an in-kernel dynamic
// scheduler (heaviest tiles first).  The unit of the shared-memory pipeline is a STAGE = one 64-channel K slice: the
// pre-swizzled weight slice (c_out x 128 B) plus the G gathered 128 x 64 A tiles that multiply it, behind ONE full /
// ONE empty mbarrier.  The K dimension is a flat stream of 16-byte chunks (8 channels): all chunks of offset 0 (first
// source tensor, then the second), then offset 1, ...; a slice is the next 8 chunks of that stream, so every slice is
// full whatever the channel count: 96 channels -> 3 slices per 2 offsets (not 2 per offset), 96 + 32 concatenated ->
// 2 slices per offset (not 3), 32 channels -> 2 offsets per slice, 16 -> 4.  The pattern repeats every P = 8 / gcd(C/8, 8)
// offsets ("virtual offset" kv = k / P) with Q = (C/8) / gcd slices.  Warp roles (768 threads):
//   warps 0-3   epilogue   tcgen05.ld the G 128 x c_out fp32 accumulators (one TMEM lane quadrant per warp),
//                          + bias + residual, ReLU, store bf16/fp32 rows (to row perm[r] when tiles are mask-sorted)
//   warps 4-5   MMA        one issuing thread per sub-tile: tcgen05.mma (M=128, N=c_out, K=16) per 16 input channels,
//                          tcgen05.commit on the stage's empty barrier; warp 4 owns the TMEM allocation
//   warp  6     weights    one thread streams the weight slice into the stage with cp.async.bulk (UBLKCP, complete_tx)
//   warp  7     scheduler  one thread takes super-tile tickets from a global counter and publishes them in a 4-deep
//                          shared-memory ring read by every other warp (the last CTA to finish re-zeroes the counter)
//   warps 8-23  producers  4 independent groups of 4 warps; group g fills every 4th stage: it gathers the neighbour rows
//                          with 16-byte cp.async (LDGSTS, zero-fill for missing neighbours) into the 128B-swizzled K-major
//                          A tiles, prefetches the indices of its next stage while the rows land, then cp.async.wait_all,
//                          fence.proxy.async and ONE mbarrier arrival per warp.
// History (profiles/README.md): v1 paid ~200 producer instructions per 16 KB slot and was bound by producer issue; v3
// became bound by the single MMA thread's instruction latency; v4 (8 producer warps, 64-bit address arithmetic, static
// tile striding) was bound by the producers' own dependent instruction chains (~125 SASS instructions per stage per
// warp, stall_wait) and by a 2x imbalance between SMs.  v5: 16 producer warps with one IMAD.WIDE per copy, slice
// packing, dynamic scheduling.  Knock-out runs of v5 (TSG_TC_DEBUG=15: no gathers, weights, MMAs or stores) still took
// the full kernel time: the per-THREAD cp.async.mbarrier.arrive.noinc (513 arrivals per stage on one barrier word) was
// the bottleneck suspect; arriving once per warp did not help either, and clock64 traces of the hand-offs (TSG_TC_DEBUG
// bit 128) showed ~600 cycles between consecutive stages of an EMPTY pipeline: with all 16 producer warps taking part in
// every stage the per-stage bookkeeping (x16 warps) saturates instruction issue and every hand-off is on the critical
// path.  v6: producer groups own whole stages (hand-shake amortised over 8 G copies per thread, groups overlap).
// Offsets for which a sub-tile has no neighbour at all are skipped by every role (tile_mask); with rows sorted by
// their neighbour bit mask (tsg_kmap_sort_rows) that removes more than half of the (tile, offset) work.
// Accumulators are double buffered in TMEM so the epilogue of super tile t overlaps the mainloop of t+1.
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace tsg {

constexpr int V5_MMA_WARP = TC_EPI_WARPS;        // 4, 5 (one per sub-tile)
constexpr int V5_W_WARP = TC_EPI_WARPS + 2;      // 6
constexpr int V5_SCHED_WARP = TC_EPI_WARPS + 3;  // 7
constexpr int V5_PROD_WARP0 = TC_EPI_WARPS + 4;  // 8
constexpr int V5_GROUPS = 4;                                      // producer groups; each owns every NG-th stage
constexpr int V5_GROUP_WARPS = 4;
constexpr int V5_PROD_WARPS = V5_GROUPS * V5_GROUP_WARPS;         // 16
constexpr int V5_THREADS = 32 * (V5_PROD_WARP0 + V5_PROD_WARPS);  // 768
constexpr int V5_Q = TC_BM / (V5_GROUP_WARPS * 4);                // 8 consecutive tile rows per producer thread
constexpr int V5_MAX_STAGES = 8;
constexpr int V5_SCHED_SLOTS = 3;
constexpr int V5_CONSUMER_WARPS = V5_PROD_WARPS + TC_EPI_WARPS + 3;  // producers + epilogue + 2 MMA + weights

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// Profiling aid (TSG_TC_DEBUG bit 128): CTA 0 records clock64() at the pipeline hand-offs of its first TRACE_N stages.
constexpr int TRACE_N = 96;
__device__ long long g_trace[13][TRACE_N];  // 0 producer group 0 got empty, 1 it arrived on full, 2 MMA got full, 3 MMA committed,
                                           // 4 weights got empty, 5 MMA starts waiting for full
#ifdef TSG_TC_TRACE  // profiling build (TSG_TC_TRACE=1 python -m taseg_b200.build): knock-outs and traces cost nothing otherwise
#define TSG_DBG(bit) (p.dbg & (bit))
#define TSG_TRACE(role, idx)                                                              \
  do {                                                                                    \
    if ((p.dbg & 128) && blockIdx.x == 0 && (idx) < TRACE_N) g_trace[role][idx] = clock64(); \
  } while (0)
#else
#define TSG_DBG(bit) 0
#define TSG_TRACE(role, idx) do { } while (0)
#endif

template <int G>
__global__ void __launch_bounds__(V5_THREADS, 1) conv_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * V5_MAX_STAGES + 4 + 2 * V5_SCHED_SLOTS];
  __shared__ int sched_tile[V5_SCHED_SLOTS];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.pk, Q = p.kq;                                      // offsets per virtual offset, slices per virtual offset
  const unsigned subP = (1u << P) - 1u;
  const unsigned long long slice_need = p.slice_need;                // 4 bits per slice: which of the P offsets it touches
  const uint32_t b_bytes = (uint32_t)p.c_out * 128u;                 // multiple of 2048
  const uint32_t stage_bytes = b_bytes + (uint32_t)G * TC_A_BYTES;   // [W slice][A tile 0]..[A tile G-1]
  const uint32_t nst = (uint32_t)p.na;                               // stages
  const int num_tiles = (int)((p.n_out + TC_BM - 1) / TC_BM);
  const int num_super = (num_tiles + G - 1) / G;
  const unsigned kmask = p.K >= 32 ? 0xffffffffu : ((1u << p.K) - 1u);
  const int KV = (p.K + P - 1) / P;                                  // virtual offsets
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[V5_MAX_STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * V5_MAX_STAGES]), tempty0 = tfull0 + 16;
  const uint32_t sfull0 = tfull0 + 32, sempty0 = sfull0 + 8 * V5_SCHED_SLOTS;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < nst; ++s) {
      mbar_init(full0 + 8 * s, V5_GROUP_WARPS + 1);  // one arrival per warp of the owning producer group + the weight thread
      mbar_init(empty0 + 8 * s, G);                  // one tcgen05.commit per MMA warp
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, G);
      mbar_init(tempty0 + 8 * b, TC_EPI_WARPS * 32);
    }
    for (int s = 0; s < V5_SCHED_SLOTS; ++s) {
      mbar_init(sfull0 + 8 * s, 1);
      mbar_init(sempty0 + 8 * s, V5_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == V5_MMA_WARP) {  // TMEM allocation by the MMA warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // next super tile from the scheduler ring (every consumer warp calls this once per super tile, converged)
  Ring sr;
  auto next_super = [&]() -> int {
    mbar_wait(sfull0 + 8 * sr.slot, sr.phase);
    const int st = *reinterpret_cast<volatile int *>(&sched_tile[sr.slot]);
    __syncwarp();
    if (lane == 0) mbar_arrive(sempty0 + 8 * sr.slot);
    sr.advance(V5_SCHED_SLOTS);
    return st;
  };
  auto next_super_lane = [&]() -> int {  // same, for roles that run in one lane
    mbar_wait(sfull0 + 8 * sr.slot, sr.phase);
    const int st = *reinterpret_cast<volatile int *>(&sched_tile[sr.slot]);
    mbar_arrive(sempty0 + 8 * sr.slot);
    sr.advance(V5_SCHED_SLOTS);
    return st;
  };
  // real-offset masks of the G tiles of a super tile; returns their union
  auto tile_masks = [&](int st, unsigned (&masks)[G]) -> unsigned {
    unsigned um = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int tile = st * G + g;
      masks[g] = tile < num_tiles ? ((p.tile_mask ? __ldg(p.tile_mask + tile) : 0xffffffffu) & kmask) : 0u;
      um |= masks[g];
    }
    return um;
  };
  // virtual offsets with at least one real offset present in `m`
  auto virt_mask = [&](unsigned m) -> unsigned {
    if (P == 1) return m;
    unsigned vm = 0;
    for (int kv = 0; kv < KV; ++kv) vm |= ((m >> (kv * P)) & subP) ? (1u << kv) : 0u;
    return vm;
  };
  // The stages of a super tile, in the order every role walks them: for each virtual offset kv present in the union
  // mask `um`, the slices j whose offsets (bits `need`) intersect it.  All roles enumerate with the same two tests.
  auto group_bits = [&](unsigned m, int kv) -> unsigned { return (m >> (kv * P)) & subP; };
  auto need_of = [&](int j) -> unsigned { return (unsigned)(slice_need >> (4 * j)) & 15u; };

  if (warp < TC_EPI_WARPS) {
    // ================================================================= epilogue
    uint32_t it = 0;
    for (int st = next_super(); st >= 0; st = next_super(), ++it) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      unsigned masks[G];
      long long rows[G];
      tile_masks(st, masks);
#pragma unroll
      for (int g = 0; g < G; ++g) {  // destination rows are fetched before the (long) wait for the accumulators
        const long long r = (long long)(st * G + g) * TC_BM + warp * 32 + lane;
        rows[g] = r < p.n_out ? (p.perm ? (long long)__ldg(p.perm + r) : r) : -1;
      }
      mbar_wait_sleep(tfull0 + 8 * buf, ph);
      tc_fence_after();
      if (threadIdx.x == 0) TSG_TRACE(10, it);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (st * G + g >= num_tiles) break;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (buf * G + g) * (uint32_t)p.c_out;
        int c = 0;
        for (; c + 32 <= p.c_out; c += 32) {
          uint32_t v[32];
          if (masks[g]) {
            tmem_ld32(taddr + c, v);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          if (rows[g] >= 0 && !TSG_DBG(8)) {
            epilogue_store16(p, rows[g], c, *reinterpret_cast<const uint32_t(*)[16]>(&v[0]));
            epilogue_store16(p, rows[g], c + 16, *reinterpret_cast<const uint32_t(*)[16]>(&v[16]));
          }
        }
        if (c < p.c_out) {  // c_out is a multiple of 16
          uint32_t v[16];
          if (masks[g]) {
            tmem_ld16(taddr + c, v);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
          }
          if (rows[g] >= 0) epilogue_store16(p, rows[g], c, v);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty0 + 8 * buf);
      if (threadIdx.x == 0) TSG_TRACE(11, it);
    }
  } else if (warp == V5_MMA_WARP || warp == V5_MMA_WARP + 1) {
    // ================================================================= MMA issuers (one thread per sub-tile)
    // Issuing a tcgen05.mma costs the issuing thread ~56 cycles and a tcgen05.commit ~180 whatever the MMA's size
    // (tools/micro/mma_issue.cu, profiles/README.md), so for c_out <= 128 the instruction stream, not the tensor pipe,
    // bounds a stage: each of the G sub-tiles gets its own issuing warp, and the loop runs in one lane.
    const int g = warp - V5_MMA_WARP;
    if (g < G && lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.c_out >> 3) << 17) | ((TC_BM >> 4) << 24);
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      const uint32_t desc_lo_stage = stage_bytes >> 4;
      const uint32_t b_lo0 = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);          // LBO field = 1 (ignored for swizzled K-major)
      const uint32_t a_lo0 = b_lo0 + (b_bytes >> 4) + g * (TC_A_BYTES >> 4);
      uint32_t slot = 0, phase = 0, it = 0;
      int n_mma = 0;
      for (int st = next_super_lane(); st >= 0; st = next_super_lane(), ++it) {
        const uint32_t buf = it & 1, ph = (it >> 1) & 1;
        unsigned masks[G];
        const unsigned um = tile_masks(st, masks);
        const unsigned umask = virt_mask(um);
        unsigned mg = masks[0];
#pragma unroll
        for (int gg = 1; gg < G; ++gg)
          if (g == gg) mg = masks[gg];
        mbar_wait(tempty0 + 8 * buf, ph ^ 1);
        tc_fence_after();
        if (g == 0) TSG_TRACE(8, it);
        const int n_mma_tile0 = n_mma;
        const uint32_t d_tmem = tmem_base + (buf * G + g) * (uint32_t)p.c_out;
        uint32_t started = 0;
        for (int kv = next_bit(umask, -1); kv < 32; kv = next_bit(umask, kv)) {
          const unsigned gu = group_bits(um, kv), gm = group_bits(mg, kv);
          for (int j = 0; j < Q; ++j) {
            const unsigned need = need_of(j);
            if (!(gu & need)) continue;                 // no tile of the super tile needs this slice: no stage
            const bool act = (gm & need) != 0 && !TSG_DBG(4);
            if (g == 0) TSG_TRACE(5, n_mma);
            mbar_wait(full0 + 8 * slot, phase);  // producers fenced their writes towards the async proxy before arriving
            tc_fence_after();
            if (g == 0) TSG_TRACE(2, n_mma);
            if (act) {  // every slice is a full 64-channel block: four K = 16 MMAs
              const uint32_t b_lo = b_lo0 + slot * desc_lo_stage, a_lo = a_lo0 + slot * desc_lo_stage;
              umma_bf16(d_tmem, make_desc(a_lo, desc_hi), make_desc(b_lo, desc_hi), idesc, started);
              umma_bf16(d_tmem, make_desc(a_lo + 2, desc_hi), make_desc(b_lo + 2, desc_hi), idesc, 1u);
              umma_bf16(d_tmem, make_desc(a_lo + 4, desc_hi), make_desc(b_lo + 4, desc_hi), idesc, 1u);
              umma_bf16(d_tmem, make_desc(a_lo + 6, desc_hi), make_desc(b_lo + 6, desc_hi), idesc, 1u);
              started = 1;
            }
            umma_commit(empty0 + 8 * slot);  // this warp's share of "stage consumed" (arrives once its MMAs have read it)
            if (g == 0) TSG_TRACE(3, n_mma);
            ++n_mma;
            if (++slot == nst) {
              slot = 0;
              phase ^= 1;
            }
          }
        }
        umma_commit(tfull0 + 8 * buf);  // this sub-tile's accumulator is complete (immediately if there was no work)
        if (g == 0) TSG_TRACE(9, it);
#ifdef TSG_TC_TRACE
        if (g == 0 && (p.dbg & 128) && blockIdx.x == 0 && it < TRACE_N) g_trace[12][it] = n_mma - n_mma_tile0;
#endif
      }
    } else if (lane == 0) {
      while (next_super_lane() >= 0) {}  // spare MMA warp (G == 1): keep the scheduler ring moving
    }
    __syncwarp();
  } else if (warp == V5_W_WARP) {
    // ================================================================= weight loader
    if (lane == 0) {
      Ring r;
      int n_w = 0;
      for (int st = next_super_lane(); st >= 0; st = next_super_lane()) {
        unsigned masks[G];
        const unsigned um = tile_masks(st, masks);
        const unsigned umask = virt_mask(um);
        for (int kv = next_bit(umask, -1); kv < 32; kv = next_bit(umask, kv)) {
          const unsigned gu = group_bits(um, kv);
          const uint8_t *wk = p.packed_w + (size_t)kv * Q * b_bytes;
          for (int j = 0; j < Q; ++j) {
            if (!(gu & need_of(j))) continue;
            mbar_wait(empty0 + 8 * r.slot, r.phase ^ 1);
            TSG_TRACE(4, n_w);
            ++n_w;
            if (TSG_DBG(2)) {
              mbar_arrive(full0 + 8 * r.slot);
            } else {
              mbar_arrive_expect_tx(full0 + 8 * r.slot, b_bytes);
              bulk_g2s(smem_base + r.slot * stage_bytes, wk + (size_t)j * b_bytes, b_bytes, full0 + 8 * r.slot);
            }
            r.advance(nst);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == V5_SCHED_WARP) {
    // ================================================================= scheduler
    if (lane == 0) {
      Ring w;
      int static_next = blockIdx.x;
      for (;;) {
        mbar_wait(sempty0 + 8 * w.slot, w.phase ^ 1);
        int t;
        if (p.sched) {
          t = atomicAdd(p.sched, 1);
        } else {
          t = static_next;
          static_next += gridDim.x;
        }
        const int st = t < num_super ? num_super - 1 - t : -1;  // heavy (high-key) tiles first
        *reinterpret_cast<volatile int *>(&sched_tile[w.slot]) = st;
        mbar_arrive(sfull0 + 8 * w.slot);  // release: the tile index is visible to the waiters
        w.advance(V5_SCHED_SLOTS);
        if (st < 0) break;
      }
      if (p.sched) {  // every CTA draws exactly one terminal ticket: the last one re-arms the counters for the next launch
        __threadfence();
        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
          p.sched[0] = 0;
          p.sched[1] = 0;
          __threadfence();
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================= producers
    // V5_GROUPS independent groups of V5_GROUP_WARPS warps; group g owns the global stages s with s % NG == g, so a
    // hand-shake (empty wait, landing wait, proxy fence, one arrival per warp) is paid once per 8 G copies per thread,
    // and while one group waits for its rows to land the others are issuing theirs.
    const int grp = (warp - V5_PROD_WARP0) / V5_GROUP_WARPS;
    const int NG = (int)nst < V5_GROUPS ? (int)nst : V5_GROUPS;
    const int tid = threadIdx.x - 32 * (V5_PROD_WARP0 + grp * V5_GROUP_WARPS);  // 0..127 inside the group
    const int chunk = tid & 7, rsub = tid >> 3;  // 8 lanes cover one 128-byte tile row; thread owns rows 8 rsub .. 8 rsub + 7
    uint32_t dst_off[V5_Q];                      // row 8 rsub + q sits at chunk position chunk ^ q of its 128-byte line
#pragma unroll
    for (int q = 0; q < V5_Q; ++q) dst_off[q] = b_bytes + (uint32_t)(rsub * V5_Q + q) * 128u + (uint32_t)((chunk ^ q) << 4);
    const int cpo = p.cpo, c0c = p.c0 >> 3;      // chunks per offset, chunks of the first source tensor
    const uint32_t rb0 = (uint32_t)p.c0 * 2u, rb1 = (uint32_t)p.c1 * 2u;
    const char *in0 = reinterpret_cast<const char *>(p.in0);
    const char *in1 = reinterpret_cast<const char *>(p.in1);
    const int K = p.K;
    const long long n_out = p.n_out, nbr_stride = p.nbr_stride;
    const int *nbr = p.nbr;
    const bool worker = grp < NG;
    uint32_t slot = (uint32_t)grp, phase = 0;  // ring position of this group's next stage (always NG stages further)
    int cnt = NG - 1;                          // global stage number modulo NG of the iterator's current stage (none yet)
    int n_issued = 0;
    for (int st = next_super(); st >= 0; st = next_super()) {
      if (!worker) continue;
      unsigned masks[G];
      const unsigned um = tile_masks(st, masks);
      const unsigned umask = virt_mask(um);
      // iterator over the stages of this super tile; every group walks all of them and keeps `cnt` in step
      int kv = next_bit(umask, -1), j = -1;
      unsigned gu = kv < 32 ? group_bits(um, kv) : 0u;
      auto next_stage = [&]() -> bool {  // advance to the next stage of the super tile; false at its end
        while (kv < 32) {
          while (++j < Q)
            if (gu & need_of(j)) {
              cnt = cnt + 1 == NG ? 0 : cnt + 1;
              return true;
            }
          kv = next_bit(umask, kv);
          j = -1;
          gu = kv < 32 ? group_bits(um, kv) : 0u;
        }
        return false;
      };
      auto next_mine = [&]() -> bool {   // advance to this group's next stage
        while (next_stage())
          if (cnt == grp) return true;
        return false;
      };
      const long long m0 = (long long)st * G * TC_BM + rsub * V5_Q;  // first of this thread's 8 consecutive tile rows
      int idx[G][V5_Q];  // neighbour row of tile row 8 rsub + q, or -1
      int k_loaded = -1;
      // this lane's chunk of slice j of virtual offset kv: offset k, source tensor, byte offset inside the source row
      int k_cur = 0;
      const char *src_cur = in0;
      uint32_t rb_cur = rb0;
      auto locate = [&]() {
        const int f = 8 * j + chunk, ksub = f / cpo, cc = f - ksub * cpo;
        k_cur = kv * P + ksub;
        const bool second = cc >= c0c;
        src_cur = (second ? in1 : in0) + (second ? cc - c0c : cc) * 16;
        rb_cur = second ? rb1 : rb0;
      };
      // eight consecutive indices = two 16-byte loads; rows past n_out read the -1 padding of the table
      auto load_idx = [&]() {
        if (k_cur == k_loaded) return;
        k_loaded = k_cur;
        const int4 *src = reinterpret_cast<const int4 *>(nbr + (long long)k_cur * nbr_stride + m0);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (k_cur < K && ((masks[g] >> k_cur) & 1u)) {
            if (nbr) {
              const int4 a = __ldg(src + g * (TC_BM / 4)), b = __ldg(src + g * (TC_BM / 4) + 1);
              idx[g][0] = a.x; idx[g][1] = a.y; idx[g][2] = a.z; idx[g][3] = a.w;
              idx[g][4] = b.x; idx[g][5] = b.y; idx[g][6] = b.z; idx[g][7] = b.w;
            } else {  // identity map (1x1x1 convolutions, point MLPs)
#pragma unroll
              for (int q = 0; q < V5_Q; ++q) idx[g][q] = m0 + g * TC_BM + q < n_out ? (int)(m0 + g * TC_BM + q) : -1;
            }
          } else {
#pragma unroll
            for (int q = 0; q < V5_Q; ++q) idx[g][q] = -1;
          }
        }
      };
      bool have = next_mine();
      if (have) {
        locate();
        load_idx();
      }
      while (have) {
        const unsigned need = need_of(j);
        unsigned act = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) act |= ((group_bits(masks[g], kv) & need) ? 1u : 0u) << g;
        const char *bp = src_cur;
        const uint32_t rb = rb_cur;
        mbar_wait(empty0 + 8 * slot, phase ^ 1);
        if (tid == 0 && grp == 0) TSG_TRACE(0, n_issued);
        const uint32_t dst = smem_base + slot * stage_bytes;
        if (!TSG_DBG(1)) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (!((act >> g) & 1u)) continue;
#pragma unroll
            for (int q = 0; q < V5_Q; ++q) {
              const int v = idx[g][q];
              cp_async16(dst + g * TC_A_BYTES + dst_off[q], bp + (unsigned long long)(unsigned)max(v, 0) * rb, v >= 0 ? 16u : 0u);
            }
          }
        }
        if (tid == 0 && grp == 0) TSG_TRACE(6, n_issued);
        // the index registers are free again: fetch the neighbour rows of this group's next stage while the copies land
        have = next_mine();
        if (have) {
          locate();
          load_idx();
        }
        if (tid == 0 && grp == 0) TSG_TRACE(7, n_issued);
        asm volatile("cp.async.wait_all;" ::: "memory");
        fence_async_proxy();  // generic-proxy writes -> visible to the async proxy (tensor core reads)
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * slot);
        if (tid == 0 && grp == 0) TSG_TRACE(1, n_issued);
        ++n_issued;
        slot += NG;
        if (slot >= nst) {
          slot -= nst;
          phase ^= 1;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == V5_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// K-slice packing of a layer with C = c0 + c1 input channels (both multiples of 16): the flat chunk stream repeats every
// P offsets with Q slices; need[j] = which of the P offsets slice j touches.
struct SlicePlan {
  int P, Q, cpo;
  unsigned long long need;
};
static SlicePlan slice_plan(int c0, int c1) {
  SlicePlan sp;
  sp.cpo = (c0 + c1) / 8;
  int g = 8;
  while (sp.cpo % g) g >>= 1;   // gcd(cpo, 8)
  sp.P = 8 / g;
  sp.Q = sp.cpo / g;
  sp.need = 0;
  for (int j = 0; j < sp.Q && j < 16; ++j) {
    const int lo = (8 * j) / sp.cpo, hi = (8 * j + 7) / sp.cpo;
    unsigned long long bits = 0;
    for (int k = lo; k <= hi; ++k) bits |= 1ull << k;
    sp.need |= bits << (4 * j);
  }
  return sp;
}

// W (K, c_in, c_out) fp32 -> per (virtual offset kv, slice j) a [c_out][64] bf16 block, K-major, 128B-swizzled: the byte
// image the MMA reads, so a linear bulk copy stages it.  Column 8 c + e of block (kv, j) is channel 8 cc + e of offset
// kv P + ksub with 8 j + c = ksub cpo + cc.  Optional per-output-channel scale (folded BatchNorm).
__global__ void pack_weights_kernel(const float *__restrict__ w, int K, int c_in, int c_out, int P, int Q, int cpo,
                                    const float *__restrict__ out_scale, __nv_bfloat16 *__restrict__ packed) {
  const int KV = (K + P - 1) / P;
  const long long total = (long long)KV * Q * c_out * TC_KB;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(t % TC_KB);
    const int n = (int)((t / TC_KB) % c_out);
    const int j = (int)((t / ((long long)TC_KB * c_out)) % Q);
    const int kv = (int)(t / ((long long)TC_KB * c_out * Q));
    const int f = 8 * j + (col >> 3), ksub = f / cpo, cc = f - ksub * cpo;
    const int k = kv * P + ksub, ch = cc * 8 + (col & 7);
    float v = 0.f;
    if (k < K && ch < c_in) v = w[((long long)k * c_in + ch) * c_out + n];
    if (out_scale) v *= out_scale[n];
    const long long blk = ((long long)kv * Q + j) * c_out * TC_KB;
    const int sw = (((col >> 3) ^ (n & 7)) << 3) | (col & 7);
    packed[blk + (long long)n * TC_KB + sw] = __float2bfloat16_rn(v);
  }
}

}  // namespace tsg

using namespace tsg;

extern "C" {

size_t tsg_conv_pack_bytes(int k, int c0, int c1, int c_out) {
  const SlicePlan sp = slice_plan(c0, c1);
  return (size_t)((k + sp.P - 1) / sp.P) * sp.Q * c_out * TC_KB * 2;
}

int tsg_conv_pack_weights(const float *weight, int k, int c_in, int c_out, int c0, int c1, const float *out_scale,
                          void *packed, tsg_stream_t stream) {
  if (c0 % 16 || c1 % 16 || c_out % 16 || c_out > 256 || c_out <= 0 || c0 <= 0 || c0 + c1 < c_in) {
    set_error("tsg_conv_pack_weights: need c0,c1,c_out multiples of 16, c_out<=256, c0+c1>=c_in");
    return TSG_ERR_UNSUPPORTED;
  }
  const SlicePlan sp = slice_plan(c0, c1);
  if (sp.Q > 16) {
    set_error("tsg_conv_pack_weights: unsupported channel count (more than 16 slices per offset group)");
    return TSG_ERR_UNSUPPORTED;
  }
  const long long total = (long long)((k + sp.P - 1) / sp.P) * sp.Q * c_out * TC_KB;
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, stream>>>(weight, k, c_in, c_out, sp.P, sp.Q, sp.cpo, out_scale,
                                                                (__nv_bfloat16 *)packed);
  return check_launch("tsg_conv_pack_weights");
}

/* profiling aid, not part of the public header: copies the trace of the last TSG_TC_DEBUG&128 launch to `host` (13 x 96 int64) */
int tsg_debug_conv_trace(long long *host) {
  TSG_CUDA(cudaDeviceSynchronize());
  TSG_CUDA(cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * 13 * TRACE_N));
  return TSG_OK;
}

int tsg_conv_fwd_tc(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                    int c_out, const int32_t *nbr, int64_t nbr_stride, const uint32_t *tile_mask, const int32_t *perm,
                    int64_t n_out, void *out, int out_dtype, const float *bias, const void *residual, int relu,
                    int num_sms_hint, int32_t *sched, tsg_stream_t stream) {
  if (c0 % 16 || c1 % 16 || c_out % 16 || c_out > 256 || c_out <= 0 || c0 <= 0 || k <= 0 || k > 32 ||
      (out_dtype != TSG_BF16 && out_dtype != TSG_F32)) {
    set_error("tsg_conv_fwd_tc: need c0,c1,c_out multiples of 16, c_out<=256, K<=32, out bf16/f32");
    return TSG_ERR_UNSUPPORTED;
  }
  if (n_out <= 0) return TSG_OK;
  if (nbr && (nbr_stride % 256 || nbr_stride < (n_out + 255) / 256 * 256)) {
    set_error("tsg_conv_fwd_tc: nbr_stride must be a multiple of 256 covering n_out (padding rows hold -1)");
    return TSG_ERR_INVALID;
  }
  if (n_out >= (1ll << 31) - 4 * TC_BM || n_in >= (1ll << 31)) {
    set_error("tsg_conv_fwd_tc: tensor too large for 32-bit row arithmetic");
    return TSG_ERR_UNSUPPORTED;
  }
  const SlicePlan sp = slice_plan(c0, c1);
  if (sp.Q > 16) {
    set_error("tsg_conv_fwd_tc: unsupported channel count (more than 16 slices per offset group)");
    return TSG_ERR_UNSUPPORTED;
  }
  TcParams p;
  p.in0 = (const __nv_bfloat16 *)in0;
  p.in1 = (const __nv_bfloat16 *)in1;
  p.c0 = c0;
  p.c1 = c1;
  p.pk = sp.P;
  p.kq = sp.Q;
  p.cpo = sp.cpo;
  p.slice_need = sp.need;
  p.packed_w = (const uint8_t *)packed_w;
  p.K = k;
  p.c_out = c_out;
  p.nbr = nbr;
  p.nbr_stride = nbr_stride;
  p.tile_mask = tile_mask;
  p.perm = perm;
  p.n_out = n_out;
  p.out = out;
  p.out_f32 = out_dtype == TSG_F32;
  p.bias = bias;
  p.residual = (const __nv_bfloat16 *)residual;
  p.relu = relu;
  p.sched = sched;
  static const char *dbg_env = getenv("TSG_TC_DEBUG");  // profiling knock-outs (trace build only, wrong results): 1 no gathers,
  p.dbg = dbg_env ? atoi(dbg_env) : 0;                  // 2 no weight copies, 4 no MMAs, 8 no epilogue stores, 128 trace
  const int sms = num_sms_hint > 0 ? num_sms_hint : num_sms();
  const long long num_tiles = (n_out + TC_BM - 1) / TC_BM;
  // G sub-tiles share every weight slice; bounded by TMEM (2 buffers x G x c_out fp32 columns <= 512) and by the
  // number of super tiles needed to keep every SM busy
  int G = c_out <= 128 ? 2 : 1;  // (G = 4 spills the producers' index registers at 768 threads per CTA)
  while (G > 1 && (num_tiles + G - 1) / G < 2LL * sms) G >>= 1;
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)G * (uint32_t)c_out) cols <<= 1;
  p.tmem_cols = cols;
  const size_t b_bytes = (size_t)c_out * 128, stage_bytes = b_bytes + (size_t)G * TC_A_BYTES, budget = 222 * 1024;
  int stages = (int)(budget / stage_bytes);
  if (stages > V5_MAX_STAGES) stages = V5_MAX_STAGES;
  if (stages < 2) {
    set_error("tsg_conv_fwd_tc: not enough shared memory for the pipeline");
    return TSG_ERR_UNSUPPORTED;
  }
  p.na = stages;
  p.nb = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static bool configured = false;
  if (!configured) {
    TSG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    TSG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured = true;
  }
  const long long num_super = (num_tiles + G - 1) / G;
  const unsigned grid = (unsigned)(num_super < sms ? num_super : sms);
  if (G == 2) conv_tc_kernel<2><<<grid, V5_THREADS, smem, stream>>>(p);
  else conv_tc_kernel<1><<<grid, V5_THREADS, smem, stream>>>(p);
  return check_launch("tsg_conv_fwd_tc");
}

}  // extern "C"
