// Sparse convolution on the 5th-generation tensor cores (sm_100a): output-stationary implicit GEMM,
//   out[row(r),:] = epilogue( sum_k in[nbr[k,r],:] @ W[k] ),   bf16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM (896 threads) walks "super tiles" of G x 128 tile rows (G in {1,2}).  The unit of the
// shared-memory pipeline is a STAGE = one 64-channel K slice: the pre-swizzled weight slice (c_out x 128 B) plus the G
// gathered 128 x 64 A tiles that multiply it, behind ONE full / ONE empty mbarrier.  The K dimension is a flat stream of
// 16-byte chunks (8 channels): all chunks of offset 0 (first source tensor, then the second), then offset 1, ...; a
// slice is the next 8 chunks of that stream, so every slice is full whatever the channel count: 96 channels -> 3 slices
// per 2 offsets (not 2 per offset), 96 + 32 concatenated -> 2 slices per offset (not 3), 32 channels -> 2 offsets per
// slice, 16 -> 4.  The pattern repeats every P = 8 / gcd(C/8, 8) offsets ("virtual offset" kv = k / P) with
// Q = (C/8) / gcd slices.  Offsets for which a sub-tile has no neighbour at all are skipped (tile_mask); with rows sorted
// by their neighbour bit mask (tsg_kmap_sort_rows) that removes about two thirds of the (tile, offset) work.
// Warp roles:
//   warps 0-7    epilogue   two sets over the four TMEM lane quadrants: tcgen05.ld the 128 x c_out fp32 accumulators,
//                           + bias + residual, ReLU, through a padded staging buffer to coalesced 64-byte row segments
//                           (row perm[r] when tiles are mask-sorted)
//   warps 8-9    MMA        one issuing thread each, alternating stages, ordered by an mbarrier hand-off: tcgen05.mma
//                           (M=128, N=c_out, K=16) x 4 x G per stage, tcgen05.commit on the stage's empty barrier;
//                           warp 8 owns the TMEM allocation
//   warp  10     weights    one thread streams the weight slice into the stage with cp.async.bulk (UBLKCP, complete_tx)
//   warp  11     planner    draws super-tile tickets from a global counter (heaviest first) and expands each into a stage
//                           list {kv, slice, active sub-tiles, first-stage flags} in a 4-slot shared-memory ring
//   warps 12-27  producers  4 groups of 4 warps, group g fills every 4th stage: neighbour indices arrive as 128-byte
//                           lines in a private double-buffered buffer (cp.async, one stage ahead); rows are gathered with
//                           16-byte cp.async (LDGSTS, zero-fill for missing neighbours) into the 128B-swizzled K-major A
//                           tiles; hand-off by cp.async.mbarrier.arrive.noinc (nothing waits for rows to land)
// Accumulators are double buffered in TMEM so the epilogue of super tile t overlaps the main loop of t+1.
// v16: a launch may carry a second K PHASE — the 1x1x1 shortcut convolution of a residual block, gathered through the
// centre offset's index line and accumulated into the same TMEM tile (no shortcut launch, no residual read: 66 -> 59
// launches and -0.2 ms per benchmark step).
// v17: conflict-free epilogue staging (64-byte pitch, XOR swizzle); the first ticket of a CTA is its block index; K-split work
// items (tsg_conv_split_items / tsg_conv_fwd_tc4, off by default).
// v18: the planner's LOOK-AHEAD THROTTLE.  The plan ring let every CTA draw four tickets in its first microsecond, so
// launches with 2-3 tiles per SM were assigned statically and the CTAs holding the heaviest tiles got the most tiles
// (per-CTA timelines: stride-8 layers finished between 21 and 50 us).  A ticket is now drawn only when fewer than
// 6 x stages + 4 planned stages remain in front of the MMA issuers (template argument THR, chosen by the host for launches
// with <= 8 work items per CTA): -22 % on the stride-8 layers, 3858 -> 3590 us of convolutions per benchmark step.
// Also v18: the CTA-PAIR variant (template argument PAIR, tcgen05.mma.cta_group::2, see the kernel's comment) — built,
// parity-tested, measured no faster on any layer class and therefore off by default (TSG_TC_PAIR).
// Measured-and-rejected switches stay behind environment variables: (tile, column half) work items (TSG_TC_NSPLIT=2),
// programmatic dependent launch (TSG_TC_PDL=1), K split (TSG_SPLIT_K=1), pairs (TSG_TC_PAIR=1/2); see the host function.
// History and measurements (profiles/README.md): v1-v7 were bound, in turn, by producer instruction count, the single MMA
// thread, per-stage bookkeeping done by all 16 producer warps (~2000 cycles of branchy scalar code per group stage) and a
// thread-per-row epilogue; v9 moved the stage enumeration into the planner warp, v12 doubled and coalesced the epilogue,
// v13 made the producer hand-off asynchronous, v14/v15 alternate two MMA issuers.  What bounds v18 (ncu of every layer
// class, profiles/r02/ncu_conv_final_summary.md): the shared-memory data pipe on the 96-channel layers at strides 1-2
// (LDGSTS + LDS/STS wavefronts 38-44 % of cycles + tensor-core operand reads 24-33 %), L2 -> SM weight traffic on 256 -> 256
// at stride 8 (10.4 TB/s), the producers' latency chain and ~10 us of fixed cost per launch on the other one-tile launches.
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace tsg {

constexpr int V8_EPI_WARPS = 8;                  // 0-3 and 4-7: two sets over the four TMEM lane quadrants (warp % 4)
constexpr int V8_MMA_WARP = V8_EPI_WARPS;        // 8, 9 (one issuer per sub-tile)
constexpr int V8_W_WARP = V8_EPI_WARPS + 2;      // 10
constexpr int V8_SCHED_WARP = V8_EPI_WARPS + 3;  // 11
constexpr int V8_PROD_WARP0 = V8_EPI_WARPS + 4;  // 12
constexpr int V8_GROUPS = 4;                                      // producer groups; each owns every NG-th stage
constexpr int V8_GROUP_WARPS = 4;
constexpr int V8_PROD_WARPS = V8_GROUPS * V8_GROUP_WARPS;         // 16
constexpr int V8_THREADS = 32 * (V8_PROD_WARP0 + V8_PROD_WARPS);  // 896
constexpr int V8_Q = TC_BM / (V8_GROUP_WARPS * 4);                // 8 consecutive tile rows per producer thread
constexpr int V8_MAX_STAGES = 8;
constexpr int V8_PLAN_SLOTS = 4;
constexpr int V8_PLAN_MAX = 512;                                  // stages of one super tile: <= 32 virtual offsets x 16 slices
constexpr int V8_CONSUMER_WARPS = V8_PROD_WARPS + V8_EPI_WARPS + 3;  // producers + epilogue + 2 MMA + weights
// Epilogue staging: per epilogue warp 32 rows x 64 B of output; 16-byte chunk c of row r sits at chunk position
// c ^ ((r >> 1) & 3).  Both access patterns are then conflict-free: eight threads walking their own rows write the same
// chunk c of rows r..r+7 into eight different 16-byte bank groups ((r & 1) * 4 + (c ^ (r >> 1 & 3))), and eight lanes reading
// the four chunks of two consecutive rows cover one 128-byte wavefront.  (v16 padded rows to 80 B instead: the row-pair
// reads were 2-way conflicted, 8 wavefronts per LDS.128 in the ncu source view.)
constexpr int V8_STG_PITCH = 64;
constexpr int V8_STG_BYTES = 32 * V8_STG_PITCH;  // 2048
__device__ __forceinline__ uint32_t stg_chunk(uint32_t stg, int r, int c) {
  return stg + (uint32_t)r * V8_STG_PITCH + (uint32_t)((c ^ ((r >> 1) & 3)) << 4);
}
constexpr int V8_DYN_SMEM = 221 * 1024;          // dynamic shared memory requested per CTA

// What the planner publishes per super tile: the tile, the offset masks of its G tiles and the list of pipeline stages
// (stage = bits 0-4 virtual offset kv, 5-8 slice j, 9-10 which sub-tiles multiply this slice, 11-12 it is the first
// such stage of the super tile for sub-tile 0 / 1, 13 the K phase).
struct __align__(16) Plan {
  int tile, n;      // n: bits 0-15 number of stages, bit 16 column half of the work item (N split)
  unsigned mask[2];
  int split, slot;  // K split (work-item lists): bits 0-7 part of the tile's offsets this item sums, 8-15 number of parts; partial-sum slot
  unsigned short stage[V8_PLAN_MAX];
};

__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Profiling aid (TSG_TC_DEBUG bit 128): CTA 0 records clock64() at the pipeline hand-offs of its first TRACE_N stages.
constexpr int TRACE_N = 96;
__device__ long long g_trace[13][TRACE_N];  // 0 producer group 0 got empty, 1 it arrived on full, 2 MMA got full, 3 MMA committed,
                                           // 4 weights got empty, 5 MMA starts waiting for full, 6 copies issued, 7 next indices requested
__device__ long long g_cta[160][4];         // trace build: per CTA {globaltimer at entry, at exit, stages issued by MMA warp 0, tiles}
__device__ long long g_life[8];             // trace build: CTA 0's clock64 at kernel entry, after the prologue barrier, first plan
                                            // published, first plan seen by producers, last epilogue done, before exit
#ifdef TSG_TC_TRACE  // profiling build (TSG_TC_TRACE=1 python -m taseg_b200.build): knock-outs and traces cost nothing otherwise
#define TSG_CTA(j, v)                                                      \
  do {                                                                     \
    if ((p.dbg & 128) && blockIdx.x < 160) g_cta[blockIdx.x][j] = (v);     \
  } while (0)
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TSG_LIFE(i)                                                        \
  do {                                                                     \
    if ((p.dbg & 128) && blockIdx.x == 0) g_life[i] = clock64();           \
  } while (0)
#define TSG_DBG(bit) (p.dbg & (bit))
#define TSG_TRACE(role, idx)                                                              \
  do {                                                                                    \
    if ((p.dbg & 128) && blockIdx.x == 0 && (idx) < TRACE_N) g_trace[role][idx] = clock64(); \
  } while (0)
#else
#define TSG_DBG(bit) 0
#define TSG_TRACE(role, idx) do { } while (0)
#define TSG_LIFE(i) do { } while (0)
#define TSG_CTA(j, v) do { } while (0)
#endif

// ---- epilogue building blocks: NCOLS accumulator columns of 32 tile rows (one per lane) through the warp's staging buffer
template <int NCOLS>
__device__ __forceinline__ void epi_prefetch_res(uint32_t stg, const char *res_c0, long long pitch, int rows_g, int lane) {
  constexpr int CPR = NCOLS / 8;  // 16-byte chunks of bf16 per row
#pragma unroll
  for (int it = 0; it < CPR; ++it) {
    const int r = it * (32 / CPR) + lane / CPR, q = lane % CPR;
    const int orow = __shfl_sync(0xffffffffu, rows_g, r);
    cp_async16(stg_chunk(stg, r, q), res_c0 + (long long)max(orow, 0) * pitch + q * 16, orow >= 0 ? 16u : 0u);
  }
}

// Wait until *flag == want (set by a warp of another, already running CTA); same watchdog policy as mbar_wait.
__device__ __forceinline__ void spin_until(const int *flag, int want) {
  long long t0 = 0;
  for (uint32_t spin = 1; *reinterpret_cast<const volatile int *>(flag) != want; ++spin) {
    __nanosleep(64);
    if ((spin & 0x3ffu) == 0) {
      const long long now = clock64();
      if (!t0) t0 = now;
      else if (now - t0 > (1ll << 33)) __trap();
    }
  }
}

// K split: every work item of a tile but the last to arrive parks its fp32 accumulators in its own slab of global memory.
// Slab layout: [16-byte column chunk][tile row] float4, so the 32 lanes of a warp (32 consecutive tile rows) write and read
// 512 contiguous bytes per instruction (a row-major slab made the exchange 13 us of a 53 us launch: every lane its own sector).
template <int NCOLS>
__device__ __forceinline__ void epi_store_partial(uint32_t taddr, float4 *slab_row) {   // slab_row: this lane's row, first chunk
#pragma unroll
  for (int cc = 0; cc < NCOLS; cc += 16) {
    uint32_t v[16];
    tmem_ld16(taddr + cc, v);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      __stcg(slab_row + (cc / 4 + j) * TC_BM,
             make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                         __uint_as_float(v[4 * j + 3])));
  }
}

// ... and the last one sums all parts IN PART ORDER (its own from TMEM, the others from their slabs) before bias / residual /
// ReLU: the result does not depend on which item arrived last.
struct SplitParts {
  const float4 *slab_row;   // part 0's slab: this lane's row, first chunk of the column block in hand; nullptr = not split
  int nparts, mine;
  size_t part_stride;       // float4s between the slabs of consecutive parts
};
template <int NCOLS, bool F32>
__device__ __forceinline__ void epi_block(const TcParams &p, uint32_t taddr, bool have_acc, const float *bias_c, uint32_t stg,
                                          int lane, int rows_g, int c0, bool res_staged, bool no_store,
                                          const SplitParts sp = SplitParts{nullptr, 1, 0, 0}) {
  const uint32_t my_row = stg + lane * V8_STG_PITCH;
  const int my_x = (lane >> 1) & 3;                       // chunk c of this thread's row sits at my_row + ((c ^ my_x) << 4)
  auto my_chunk = [&](int c) -> uint32_t { return my_row + (uint32_t)((c ^ my_x) << 4); };
  // (A) every thread finishes NCOLS columns of its own row: accumulator + bias (+ residual), ReLU, convert, into staging
#pragma unroll
  for (int cc = 0; cc < NCOLS; cc += 16) {
    uint32_t v[16];
    if (have_acc) {
      tmem_ld16(taddr + cc, v);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0u;
    }
    if (sp.slab_row) {
      float a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = 0.f;
      for (int q = 0; q < sp.nparts; ++q) {
        if (q == sp.mine) {
#pragma unroll
          for (int j = 0; j < 16; ++j) a[j] += __uint_as_float(v[j]);
        } else {
          const float4 *src = sp.slab_row + (size_t)q * sp.part_stride + (cc / 4) * TC_BM;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 o = __ldcg(src + j * TC_BM);
            a[4 * j] += o.x; a[4 * j + 1] += o.y; a[4 * j + 2] += o.z; a[4 * j + 3] += o.w;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(a[j]);
    }
    float f[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b = *reinterpret_cast<const float4 *>(bias_c + cc + 4 * j);
      f[4 * j] = __uint_as_float(v[4 * j]) + b.x;
      f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
      f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
      f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
    }
    if (p.residual) {
      uint4 r0, r1;
      if (res_staged) {
        r0 = lds128(my_chunk(cc >> 3));
        r1 = lds128(my_chunk((cc >> 3) + 1));
      } else if (rows_g >= 0) {  // fp32 output with a residual (not on the engine's path): direct loads
        const uint4 *rp = reinterpret_cast<const uint4 *>(p.residual + (long long)rows_g * p.c_out + c0 + cc);
        r0 = __ldg(rp);
        r1 = __ldg(rp + 1);
      } else {
        r0 = r1 = make_uint4(0u, 0u, 0u, 0u);
      }
      const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[2 * j] += __uint_as_float(rw[j] << 16);
        f[2 * j + 1] += __uint_as_float(rw[j] & 0xffff0000u);
      }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (F32) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(my_chunk((cc >> 2) + j), make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                                                   __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3])));
    } else {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t *>(&h);
      }
      sts128(my_chunk(cc >> 3), make_uint4(w[0], w[1], w[2], w[3]));
      sts128(my_chunk((cc >> 3) + 1), make_uint4(w[4], w[5], w[6], w[7]));
    }
  }
  __syncwarp();
  // (B) the warp streams the block out: consecutive lanes on consecutive 16-byte chunks of a row
  constexpr int ESZ = F32 ? 4 : 2;
  constexpr int CPR = NCOLS * ESZ / 16;
  char *out_c0 = reinterpret_cast<char *>(p.out) + (long long)c0 * ESZ;
  const long long pitch = (long long)p.c_out * ESZ;
#pragma unroll
  for (int it = 0; it < CPR; ++it) {
    const int r = it * (32 / CPR) + lane / CPR, q = lane % CPR;
    const int orow = __shfl_sync(0xffffffffu, rows_g, r);
    const uint4 val = lds128(stg_chunk(stg, r, q));
    if (orow >= 0 && !no_store) *reinterpret_cast<uint4 *>(out_c0 + (long long)orow * pitch + q * 16) = val;
  }
  __syncwarp();
}

// PAIR (v18): a cluster of two CTAs on the two SMs of a TPC computes TWO tiles with tcgen05.mma.cta_group::2 (M = 256:
// rows 0-127 = the leader CTA's tile, in the leader's TMEM; rows 128-255 = the peer's).  Every CTA gathers its own A
// tile and stages HALF of the weight slice (output channels [rank N/2, rank N/2 + N/2)): the weight bytes an SM pulls
// from the L2, writes to and reads from shared memory are halved — the bound of the c_out >= 128 layers (L2 -> SM
// traffic is two thirds weights at c_out = 256) — and the stage shrinks, so more stages are in flight.  Protocol:
//   * both CTAs plan the same unit (tiles 2u, 2u + 1; static round-robin over clusters, heaviest first) and derive the
//     same stage list = union of the two tiles' needs; a tile that does not need a stage presents zero rows;
//   * full barriers stay CTA-local (cp.async / bulk completions signal only the own CTA); the peer's first MMA warp is a
//     RELAY: wait local full, proxy fence, mbarrier.arrive on the leader's pfull[slot] through the cluster window;
//   * only the leader issues MMAs (two alternating issuers as before) after full + pfull; tcgen05.commit ... multicast
//     arrives on empty[slot] / tfull[buf] of BOTH CTAs;
//   * the peer's epilogue threads release the accumulator buffer on the LEADER's tempty (remote arrive).
// THR: the planner's look-ahead throttle (heaviest-first list scheduling for launches with few tiles per CTA) is compiled in.
// SC: the launch carries the second K phase (a folded shortcut).  A template argument because the phase selects sit in the
// planner's and the producers' hot loops and most launches have one phase: the one-phase kernels are smaller and faster.
// F32: fp32 output rows (the classifier heads, the split-precision fp32 path); ITEMS: work-item list with K split (G = 1).
// Template arguments for the same reason: each removes an epilogue variant (or the partial-sum exchange) from the others.
template <int G, bool PAIR, bool THR, bool SC, bool F32, bool ITEMS>
__global__ void __launch_bounds__(V8_THREADS, 1) conv_tc_kernel(const TcParams p) {
  static_assert(!ITEMS || (G == 1 && !PAIR), "work items address single tiles of a single CTA");
  static_assert(!PAIR || G == 1, "a pair CTA holds one tile");
  constexpr int GP = PAIR ? 2 : G;   // tiles per plan (unit of scheduling)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * V8_MAX_STAGES + 4 + 2 * V8_PLAN_SLOTS + 2];
  __shared__ Plan plans[V8_PLAN_SLOTS];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[256];
  __shared__ uint16_t lut[256];  // per (phase, slice j, 16-byte chunk c of the slice): which offset / source tensor / source chunk
  __shared__ int issued_s;       // global stage number the first MMA issuer has reached (the planner's look-ahead throttle)

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NPH = SC ? 2 : 1;
  const int P0 = p.ph[0].pk, Q0 = p.ph[0].kq, P1 = p.ph[1].pk, Q1 = p.ph[1].kq;  // offsets / slices per virtual offset
  const unsigned long long need0 = p.ph[0].slice_need, need1 = p.ph[1].slice_need;
  const int NS = p.ns, n_eff = p.n_eff;
  uint32_t rank = 0;                                                 // CTA rank in the pair
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int my_g = PAIR ? (int)rank : 0;                             // this CTA's tile among the plan's tiles (PAIR)
  const uint32_t b_bytes = (uint32_t)n_eff * (PAIR ? 64u : 128u);    // weight rows this CTA stages; multiple of 1024
  const uint32_t b_full = (uint32_t)p.c_out * 128u;                  // ... of all output channels (block pitch of packed_w)
  const uint32_t stage_bytes = b_bytes + (uint32_t)G * TC_A_BYTES;   // [W slice][A tile 0]..[A tile G-1]
  const uint32_t nst = (uint32_t)p.na;                               // stages
  const uint32_t stg0 = smem_base + nst * stage_bytes;               // epilogue staging, then the producers' index buffers
  const uint32_t idx0 = stg0 + V8_EPI_WARPS * V8_STG_BYTES;
  const uint32_t idx_warp_bytes = (uint32_t)p.ksmax * G * 128u;      // per producer warp, twice: [offset of the slice][sub-tile][32 rows]
  const unsigned kmask = p.ph[0].K >= 32 ? 0xffffffffu : ((1u << p.ph[0].K) - 1u);
  const int KV0 = (p.ph[0].K + P0 - 1) / P0, KV1 = NPH > 1 ? (p.ph[1].K + P1 - 1) / P1 : 0;  // virtual offsets
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[V8_MAX_STAGES]);
  const uint32_t pfull0 = smem_u32(&bars[2 * V8_MAX_STAGES]);       // PAIR, leader: "the peer's half of the stage is full"
  const uint32_t tfull0 = smem_u32(&bars[3 * V8_MAX_STAGES]), tempty0 = tfull0 + 16;
  const uint32_t sfull0 = tfull0 + 32, sempty0 = sfull0 + 8 * V8_PLAN_SLOTS;
  const uint32_t obar0 = sempty0 + 8 * V8_PLAN_SLOTS;                // issue-order hand-off between the two MMA issuers (G == 1)

  if (threadIdx.x == 0) TSG_LIFE(0);
  if (threadIdx.x == 0) TSG_CTA(0, gtime());
  // programmatic dependent launch: the next kernel of the stream may start its prologue while this grid drains
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < nst; ++s) {
      mbar_init(full0 + 8 * s, V8_GROUP_WARPS * 32 + 1);  // every thread of the owning producer group (async, when its copies land) + the weight thread
      mbar_init(empty0 + 8 * s, 1);                  // one tcgen05.commit, from the issuer that owns the stage
      mbar_init(pfull0 + 8 * s, 1);                  // the peer's relay thread
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 2);                  // both MMA warps commit once per super tile
      mbar_init(tempty0 + 8 * b, V8_EPI_WARPS * 32 * (PAIR ? 2 : 1));   // PAIR: the epilogue threads of both CTAs
      mbar_init(obar0 + 8 * b, 1);
    }
    for (int s = 0; s < V8_PLAN_SLOTS; ++s) {
      mbar_init(sfull0 + 8 * s, 1);
      mbar_init(sempty0 + 8 * s, V8_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ... and nothing produced by the previous kernel is read before it has completed and flushed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long n_rows = dev_count(p.n_out_dev, p.n_out);   // host value, or the device counter of the sync-free pipeline
  const int num_tiles = (int)((n_rows + TC_BM - 1) / TC_BM);
  const int num_super = (num_tiles + GP - 1) / GP;
  if (threadIdx.x == 0) issued_s = 0;
  if (threadIdx.x < 256) bias_s[threadIdx.x] = (p.bias && (int)threadIdx.x < p.c_out) ? __ldg(p.bias + threadIdx.x) : 0.f;
  if (threadIdx.x >= 256 && threadIdx.x < 256 + 256) {
    const int t = threadIdx.x - 256, phi = t >> 7, j = (t >> 3) & 15, c = t & 7;
    uint32_t e = 0;
    if (phi < NPH && j < (phi ? Q1 : Q0)) {
      const int cpo = phi ? p.ph[1].cpo : p.ph[0].cpo, c0c = (phi ? p.ph[1].c0 : p.ph[0].c0) >> 3;
      const int f = 8 * j + c, ksub = f / cpo, cc = f - ksub * cpo, lo = (8 * j) / cpo;
      const bool second = cc >= c0c;
      e = (uint32_t)(ksub - lo) | ((uint32_t)ksub << 2) | ((second ? 1u : 0u) << 4) | ((uint32_t)(second ? cc - c0c : cc) << 5);
    }
    lut[t] = (uint16_t)e;
  }
  if (warp == V8_MMA_WARP) {  // TMEM allocation by the MMA warp (PAIR: the same warp of both CTAs, same shared-memory slot)
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                   "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                   "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) {   // both CTAs' barriers are initialised before anything arrives on them through the cluster window
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) TSG_LIFE(1);

  auto need_of = [&](unsigned phi, int j) -> unsigned { return (unsigned)((phi ? need1 : need0) >> (4 * j)) & 15u; };
  // Plan ring, consumer side: wait for the next plan; release it when the role is done with the super tile.
  Ring pr;
  auto plan_wait = [&]() -> const volatile Plan * {
    mbar_wait(sfull0 + 8 * pr.slot, pr.phase);
    return &plans[pr.slot];
  };
  auto plan_release_lane = [&]() {  // from the one running lane of a single-lane role
    mbar_arrive(sempty0 + 8 * pr.slot);
    pr.advance(V8_PLAN_SLOTS);
  };
  auto plan_release_warp = [&]() {  // converged warp
    __syncwarp();
    if (lane == 0) mbar_arrive(sempty0 + 8 * pr.slot);
    pr.advance(V8_PLAN_SLOTS);
  };

  if (warp < V8_EPI_WARPS) {
    // ================================================================= epilogue
    // Accumulator rows live one per TMEM lane = one per thread, but a thread storing its own row touches 32 different
    // cache lines per warp instruction.  So each warp transposes through its staging buffer, 64 B of output per row at
    // a time (epi_block): the residual block is fetched into the buffer with coalesced cp.async, every thread finishes
    // its row there, and the warp streams the buffer out with consecutive lanes on consecutive chunks of a row.
    // Two sets of four warps (one warp per TMEM lane quadrant): with G = 2 set e drains sub-tile e, with G = 1 the sets
    // take alternate column blocks — the epilogue of a super tile must not outlast the main loop of the next one.
    const int quad = warp & 3, eset = warp >> 2;
    const uint32_t stg = stg0 + warp * V8_STG_BYTES;
    const int c_out = p.c_out;
    constexpr bool f32 = F32;
    const bool res_staged = p.residual != nullptr && !f32;
    const bool no_store = TSG_DBG(8) != 0;
    const char *resp = reinterpret_cast<const char *>(p.residual);
    const long long res_pitch = (long long)c_out * 2;
    const int g = G == 2 ? eset : 0;
    const int bw = f32 ? 16 : 32;                           // full block width in columns (64 B of output per row)
    const int cstep = G == 2 ? bw : 2 * bw;                 // this warp's blocks start at cfirst, cfirst + cstep, ...
    const int cfirst = G == 2 ? 0 : eset * bw;
    auto blk_cols = [&](int lc) -> int { return n_eff - lc >= bw ? bw : 16; };  // the last block may be 16 columns wide
    auto prefetch_res = [&](int rows_g, int c0, int ncols) {
      const char *src = resp + (long long)c0 * 2;
      if (ncols == 32) epi_prefetch_res<32>(stg, src, res_pitch, rows_g, lane);
      else epi_prefetch_res<16>(stg, src, res_pitch, rows_g, lane);
    };
    uint32_t it = 0;
    for (;; ++it) {
      const volatile Plan *pl = plan_wait();
      const int st = pl->tile;
      const unsigned mask_g = PAIR ? (pl->mask[0] | pl->mask[1]) : pl->mask[g];   // PAIR: one MMA writes both tiles' accumulators
      const int cbase = (pl->n >> 16) * n_eff;              // first output channel of this work item
      const int split = pl->split, pslot = pl->slot;
      plan_release_warp();
      if (st < 0) break;
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      const int my_tile = st * GP + (PAIR ? my_g : g);
      const long long r = (long long)my_tile * TC_BM + quad * 32 + lane;  // destination row: fetched before the long wait
      const int rows_g = r < n_rows ? (p.perm ? __ldg(p.perm + r) : (int)r) : -1;
      const bool live = my_tile < num_tiles && cfirst < n_eff;
      const bool is_split = ITEMS && (split >> 8) > 1;                 // one of the two work items of a K-split tile
      if (res_staged && live && !is_split) prefetch_res(rows_g, cbase + cfirst, blk_cols(cfirst));  // lands while the main loop still runs
      mbar_wait_sleep(tfull0 + 8 * buf, ph);
      tc_fence_after();
      if (threadIdx.x == 0) TSG_TRACE(10, it);
      if (live) {
        const bool have_acc = mask_g != 0u || NPH > 1;   // a folded shortcut multiplies every live tile
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (buf * G + g) * (uint32_t)n_eff;
        // K split: this warp and the warps with the same number in the tile's other work items own the same rows and
        // columns.  Every arrival but the last (counter state[0]) stores its accumulators in its own slab and reports them
        // visible (state[1]); the last one waits for nparts - 1 such reports, sums the parts in part order and runs the real
        // epilogue.  Nobody but the last arrival ever waits, and it waits only for warps that are already in their epilogue.
        int role = -1;                       // -1 whole tile, 0 contributor, 1 finisher
        SplitParts sp{nullptr, 1, 0, 0};
        float4 *slab = nullptr;
        int *state = nullptr;
        if (is_split) {
          const int nparts = split >> 8;
          state = p.split_state + ((size_t)pslot * V8_EPI_WARPS + warp) * 2;
          sp.part_stride = (size_t)(n_eff / 4) * TC_BM;
          slab = reinterpret_cast<float4 *>(p.split_scratch) + (size_t)pslot * p.split_parts * sp.part_stride + quad * 32 + lane;
          sp.nparts = nparts;
          sp.mine = split & 0xff;
          int a = 0;
          if (lane == 0) a = atomicAdd(state, 1);
          a = __shfl_sync(0xffffffffu, a, 0);
          role = a == nparts - 1 ? 1 : 0;
          if (role == 1) {
            if (lane == 0) spin_until(state + 1, nparts - 1);
            __syncwarp();
            __threadfence();
            if (res_staged) prefetch_res(rows_g, cbase + cfirst, blk_cols(cfirst));
          }
        }
        if (role == 0) {
          float4 *mine = slab + (size_t)sp.mine * sp.part_stride;
          for (int lc = cfirst; lc < n_eff; lc += cstep) {
            if (blk_cols(lc) == 32 && !f32) epi_store_partial<32>(taddr + lc, mine + (lc / 4) * TC_BM);
            else epi_store_partial<16>(taddr + lc, mine + (lc / 4) * TC_BM);
          }
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(state + 1, 1);
        } else {
          for (int lc = cfirst; lc < n_eff; lc += cstep) {     // lc: column inside the work item, c0: output channel
            const int ncols = blk_cols(lc), c0 = cbase + lc;
            if (res_staged) {
              if (lc != cfirst) prefetch_res(rows_g, c0, ncols);
              asm volatile("cp.async.wait_all;" ::: "memory");
              __syncwarp();
            }
            SplitParts spc = sp;
            if (role == 1) spc.slab_row = slab + (lc / 4) * TC_BM;
            if (f32) epi_block<16, true>(p, taddr + lc, have_acc, bias_s + c0, stg, lane, rows_g, c0, res_staged, no_store, spc);
            else if (ncols == 32) epi_block<32, false>(p, taddr + lc, have_acc, bias_s + c0, stg, lane, rows_g, c0, res_staged, no_store, spc);
            else epi_block<16, false>(p, taddr + lc, have_acc, bias_s + c0, stg, lane, rows_g, c0, res_staged, no_store, spc);
          }
          if (role == 1) {   // all parts are in: re-arm the tile's counters for the next launch
            __syncwarp();
            if (lane == 0) {
              state[0] = 0;
              state[1] = 0;
            }
          }
        }
      }
      tc_fence_before();
      if (PAIR && rank) mbar_arrive_remote(tempty0 + 8 * buf, 0);   // the leader's MMA warps wait for both CTAs' epilogues
      else mbar_arrive(tempty0 + 8 * buf);
      if (threadIdx.x == 0) TSG_TRACE(11, it);
      if (threadIdx.x == 0) TSG_LIFE(4);
    }
  } else if (warp == V8_MMA_WARP || warp == V8_MMA_WARP + 1) {
    // ================================================================= MMA issuers
    // One thread needs ~900 cycles per stage (barrier wait + proxy fence ~200, four tcgen05.mma + commit ~470, loop
    // ~230: traces in profiles/README.md) whatever the MMA's size, and that was the per-stage floor of v13.  The two
    // MMA warps therefore ALTERNATE stages (each issues the MMAs of all G sub-tiles of its stage), so waits, fences
    // and commits of consecutive stages overlap.  The MMAs stay in stage order through a hand-off: the issuer of stage s
    // executes tcgen05.fence::before_thread_sync and arrives on an mbarrier, the issuer of s + 1 waits for it and
    // executes tcgen05.fence::after_thread_sync — accumulation order, and the result, are those of a single issuer.
    const int mw = warp - V8_MMA_WARP;
    if (PAIR && rank != 0) {
      // peer CTA of a pair: no MMA issue.  Its two warps relay "my half of stage s is full" to the leader, alternating
      // stages like the leader's two issuers do: one thread needs ~800 cycles per stage for wait + proxy fence + remote
      // arrival (stage-level trace, profiles/r02/stage_trace_l4_256.txt: the pair kernel ran at exactly that rate with a
      // single relay thread, whatever the stage size).
      if (lane == 0) {
        uint32_t slot0 = 0, phase0 = 0, gpar = 0;
        for (;;) {
          const volatile Plan *pl = plan_wait();
          if (pl->tile < 0) {
            plan_release_lane();
            break;
          }
          const int n = pl->n & 0xffff;
          const int i0 = (int)((mw ^ gpar) & 1u);
          uint32_t slot = slot0 + i0, phase = phase0;
          if (slot >= nst) {
            slot -= nst;
            phase ^= 1;
          }
          for (int i = i0; i < n; i += 2) {
            mbar_wait(full0 + 8 * slot, phase);       // this CTA's rows (cp.async) and weight half (bulk copy) have landed
            fence_async_proxy();                      // ... ordered before the tensor cores' (async proxy) reads
            mbar_arrive_remote(pfull0 + 8 * slot, 0);
            slot += 2;
            if (slot >= nst) {
              slot -= nst;
              phase ^= 1;
            }
          }
          gpar ^= (uint32_t)n & 1u;
          slot0 += (uint32_t)n;
          while (slot0 >= nst) {
            slot0 -= nst;
            phase0 ^= 1;
          }
          plan_release_lane();
        }
      }
    } else if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_eff >> 3) << 17) | (((PAIR ? 2 * TC_BM : TC_BM) >> 4) << 24);
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      const uint32_t desc_lo_stage = stage_bytes >> 4;
      const uint32_t b_lo0 = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);          // LBO field = 1 (ignored for swizzled K-major)
      const uint32_t a_lo0 = b_lo0 + (b_bytes >> 4);
      const bool tracer = mw == 0;
      uint32_t slot0 = 0, phase0 = 0, it = 0;  // ring position of the super tile's first stage
      uint32_t gpar = 0;                       // parity of its global stage number
      uint32_t mine = 0;                       // stages this thread has issued: its k-th stage is global stage 2 k + mw
      int n_mma = 0;
      for (;; ++it) {
        const volatile Plan *pl = plan_wait();
        if (pl->tile < 0) {
          plan_release_lane();
          break;
        }
        const int n = pl->n & 0xffff;
        const uint32_t buf = it & 1, ph = (it >> 1) & 1;
        mbar_wait<PAIR>(tempty0 + 8 * buf, ph ^ 1);
        tc_fence_after();
        if (tracer) TSG_TRACE(8, it);
        const uint32_t d_tmem = tmem_base + buf * G * (uint32_t)n_eff;
        const int i0 = (int)((mw ^ gpar) & 1u);
        uint32_t slot = slot0 + i0, phase = phase0;
        if (slot >= nst) {
          slot -= nst;
          phase ^= 1;
        }
        unsigned d_next = i0 < n ? pl->stage[i0] : 0u;
        for (int i = i0; i < n; i += 2) {
          const unsigned d = d_next;
          if (i + 2 < n) d_next = pl->stage[i + 2];  // off the critical path: read before the wait
          if (tracer) TSG_TRACE(5, n_mma);
          TSG_STATE(pl->tile, n, i, n_mma);
          mbar_wait(full0 + 8 * slot, phase);  // the gathered rows (cp.async, generic proxy) and the weight slice have landed
          if (PAIR) mbar_wait<true>(pfull0 + 8 * slot, phase);   // ... and so has the peer's half (fenced and relayed by the peer)
          fence_async_proxy();                 // ... order them before this thread's tensor-core (async proxy) reads
          if (mine | mw) {                     // every stage but global stage 0 follows the other issuer's previous stage
            const uint32_t k = mw ? mine : mine - 1;
            mbar_wait(obar0 + 8 * (mw ^ 1), k & 1);
          }
          tc_fence_after();
          if (tracer) TSG_TRACE(2, n_mma);
          const uint32_t b_lo = b_lo0 + slot * desc_lo_stage;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            // every slice is a full 64-channel block: four K = 16 MMAs per sub-tile
            const uint32_t a_lo = a_lo0 + slot * desc_lo_stage + g * (TC_A_BYTES >> 4);
            const uint32_t dt = d_tmem + g * (uint32_t)n_eff;
            if (PAIR) {   // every stage of the list multiplies the pair (a tile that does not need it holds zero rows)
              if (TSG_DBG(4)) continue;
              umma_bf16_pair(dt, make_desc(a_lo, desc_hi), make_desc(b_lo, desc_hi), idesc, i == 0 ? 0u : 1u);
              umma_bf16_pair(dt, make_desc(a_lo + 2, desc_hi), make_desc(b_lo + 2, desc_hi), idesc, 1u);
              umma_bf16_pair(dt, make_desc(a_lo + 4, desc_hi), make_desc(b_lo + 4, desc_hi), idesc, 1u);
              umma_bf16_pair(dt, make_desc(a_lo + 6, desc_hi), make_desc(b_lo + 6, desc_hi), idesc, 1u);
              continue;
            }
            if (!((d >> (9 + g)) & 1u) || TSG_DBG(4)) continue;
            umma_bf16(dt, make_desc(a_lo, desc_hi), make_desc(b_lo, desc_hi), idesc, ((d >> (11 + g)) & 1u) ^ 1u);
            umma_bf16(dt, make_desc(a_lo + 2, desc_hi), make_desc(b_lo + 2, desc_hi), idesc, 1u);
            umma_bf16(dt, make_desc(a_lo + 4, desc_hi), make_desc(b_lo + 4, desc_hi), idesc, 1u);
            umma_bf16(dt, make_desc(a_lo + 6, desc_hi), make_desc(b_lo + 6, desc_hi), idesc, 1u);
          }
          tc_fence_before();
          mbar_arrive(obar0 + 8 * mw);         // the other issuer may issue the next stage
          if (PAIR) umma_commit_pair(empty0 + 8 * slot);   // "stage consumed", in both CTAs
          else umma_commit(empty0 + 8 * slot);             // "stage consumed" (arrives once the MMAs have read it)
          if (tracer) TSG_TRACE(3, n_mma);
          ++n_mma;
          ++mine;
          if (THR && mw == 0) *reinterpret_cast<volatile int *>(&issued_s) = 2 * (int)mine;   // progress for the planner's throttle
          slot += 2;
          if (slot >= nst) {
            slot -= nst;
            phase ^= 1;
          }
        }
        if (PAIR) umma_commit_pair(tfull0 + 8 * buf);
        else umma_commit(tfull0 + 8 * buf);  // this issuer's share of the accumulators is complete (immediately if it had no stage)
        if (tracer) TSG_TRACE(9, it);
#ifdef TSG_TC_TRACE
        if (tracer && (p.dbg & 128) && blockIdx.x == 0 && it < TRACE_N) g_trace[12][it] = n;
#endif
        // ring position and parity of the next super tile's first stage
        gpar ^= (uint32_t)n & 1u;
        slot0 += (uint32_t)n;
        while (slot0 >= nst) {
          slot0 -= nst;
          phase0 ^= 1;
        }
        plan_release_lane();
      }
      if (tracer) {
        TSG_CTA(2, (long long)n_mma);
        TSG_CTA(3, (long long)it);
      }
    }
    __syncwarp();
  } else if (warp == V8_W_WARP) {
    // ================================================================= weight loader
    if (lane == 0) {
      Ring r;
      int n_w = 0;
      for (;;) {
        const volatile Plan *pl = plan_wait();
        if (pl->tile < 0) {
          plan_release_lane();
          break;
        }
        const int n = pl->n & 0xffff;
        const size_t hoff = (size_t)((pl->n >> 16) + my_g) * b_bytes;   // this work item's (PAIR: this CTA's) rows of every [c_out][64] block
        for (int i = 0; i < n; ++i) {
          const unsigned d = pl->stage[i];
          const unsigned phi = SC ? (d >> 13) & 1u : 0u;
          const uint8_t *src = (phi ? p.ph[1].packed_w : p.ph[0].packed_w) +
                               (size_t)((d & 31u) * (phi ? Q1 : Q0) + ((d >> 5) & 15u)) * b_full + hoff;
          TSG_STATE(pl->tile, n, i, n_w);
          mbar_wait(empty0 + 8 * r.slot, r.phase ^ 1);
          TSG_TRACE(4, n_w);
          ++n_w;
          if (TSG_DBG(2)) {
            mbar_arrive(full0 + 8 * r.slot);
          } else {
            mbar_arrive_expect_tx(full0 + 8 * r.slot, b_bytes);
            bulk_g2s(smem_base + r.slot * stage_bytes, src, b_bytes, full0 + 8 * r.slot);
          }
          r.advance(nst);
        }
        plan_release_lane();
      }
    }
    __syncwarp();
  } else if (warp == V8_SCHED_WARP) {
    // ================================================================= planner
    // Draws super-tile tickets from a global counter (heaviest tiles first) and expands each into its stage list, one
    // virtual offset per lane, so that no other role walks masks or touches global memory to learn what to do next.
    // (v7/v8 let every warp enumerate the stages itself: ~2000 cycles of branchy scalar code per producer group stage,
    // the bottleneck of the whole kernel — profiles/README.md.)
    Ring w;
    int static_next = blockIdx.x;
    bool first_ticket = true;
    // work-item list (K split, G == 1): the list replaces the tile enumeration; its length is a device value
    constexpr bool use_items = ITEMS;
    const int n_work = use_items ? __ldg(p.n_items) : num_super * NS;
    // Look-ahead throttle (v18).  The plan ring lets the planner run up to four tiles ahead — and at kernel start it
    // did: every CTA drew four tickets in its first microsecond, so a launch with 2-3 tiles per SM was assigned
    // statically, in launch order, and the CTAs holding the heaviest tiles ended up with the most tiles (per-CTA
    // timeline, profiles/r02: stride-8 layers finished between 21 and 50 us).  A ticket is now drawn only when fewer
    // than LOOKAHEAD planned stages are left in front of the MMA issuers: long tiles take their next ticket shortly before
    // they finish (true heaviest-first list scheduling), short tiles still keep several plans in flight.
    // With many tiles per CTA the imbalance averages out and a stalled planner only costs (measured: +7 % on the
    // stride-1 layers), so the throttle is applied to launches with at most eight work items per CTA.
    const int LOOKAHEAD = p.lookahead * (int)nst + 4;   // progress is published by the MMA issuer, up to nst stages behind the producers
    // A template switch, chosen by the host from the tile count: the two-tile kernel's hot loops are sensitive to every
    // KB of code (75 KB of SASS against the instruction cache) — the throttle, compiled in but never taken, cost the
    // stride-1 layers 6 %.
    const bool throttle = THR && !PAIR && p.sched;
    int planned = 0;
    for (;;) {
      mbar_wait(sempty0 + 8 * w.slot, w.phase ^ 1);
      if (throttle) {
        long long t0 = 0;
        for (uint32_t spin = 1; planned - *reinterpret_cast<volatile int *>(&issued_s) >= LOOKAHEAD; ++spin) {
          __nanosleep(64);
          if ((spin & 0x3fffu) == 0) {   // same watchdog policy as mbar_wait
            const long long now = clock64();
            if (!t0) t0 = now;
            else if (now - t0 > (1ll << 33)) __trap();
          }
        }
      }
      int t = 0;
      if (lane == 0) {
        if (PAIR) {          // both CTAs of a pair must plan the same unit: static round-robin over the clusters
          t = static_next >> 1;
          static_next += gridDim.x;
        } else if (p.sched) {
          // the first ticket of a CTA is its block index (no round trip to the L2 before the first plan); the shared
          // counter hands out the tickets from gridDim.x on
          t = first_ticket ? (int)blockIdx.x : (int)gridDim.x + atomicAdd(p.sched, 1);
        } else {
          t = static_next;
          static_next += gridDim.x;
        }
      }
      first_ticket = false;
      t = __shfl_sync(0xffffffffu, t, 0);
      int st, half = 0, split = 1 << 8, pslot = 0;
      unsigned mm = 0, lv = 0;
      if (use_items) {
        int4 it = make_int4(-1, 0, 1 << 8, 0);
        if (t < n_work) it = __ldg(p.items + t);
        st = it.x;
        split = it.z;
        pslot = it.w;
        if (st >= 0 && lane == 0) {
          lv = st < num_tiles ? 1u : 0u;
          mm = lv ? ((unsigned)it.y & kmask) : 0u;
        }
      } else {
        st = t < n_work ? num_super - 1 - t / NS : -1;  // heavy (high-key) tiles first; NS column blocks each
        half = t < n_work ? t % NS : 0;
        if (st >= 0 && lane < GP) {
          const int tile = st * GP + lane;
          lv = tile < num_tiles ? 1u : 0u;
          mm = lv ? ((p.tile_mask ? __ldg(p.tile_mask + tile) : 0xffffffffu) & kmask) : 0u;
        }
      }
      const unsigned m0 = __shfl_sync(0xffffffffu, mm, 0), m1 = GP > 1 ? __shfl_sync(0xffffffffu, mm, 1) : 0u;
      // a folded shortcut (phase 1) is summed by part 0 of a split tile only
      const unsigned l0 = (split & 0xff) ? 0u : __shfl_sync(0xffffffffu, lv, 0), l1 = GP > 1 ? __shfl_sync(0xffffffffu, lv, 1) : 0u;
      Plan *pl = &plans[w.slot];
      // per phase: the slices of virtual offset `lane` that some tile of the super tile needs, and their positions in
      // the stage list (phase 0 first); f0 / f1 = first stage that multiplies sub-tile 0 / 1 (its first MMA overwrites
      // the accumulator)
      unsigned sl_ph[2] = {0u, 0u}, g0_ph[2] = {0u, 0u}, g1_ph[2] = {0u, 0u};
      int pos_ph[2] = {0, 0};
      int base = 0, f0 = 0x7fffffff, f1 = 0x7fffffff;
#pragma unroll
      for (int phi = 0; phi < 2; ++phi) {
        if (phi >= NPH) break;
        const int Pp = phi ? P1 : P0, Qp = phi ? Q1 : Q0, KVp = phi ? KV1 : KV0;
        const unsigned subP = (1u << Pp) - 1u;
        const unsigned a0 = phi ? l0 : m0, a1 = phi ? l1 : m1;   // the shortcut phase has one offset, present in every live tile
        const unsigned g0 = lane < KVp ? (a0 >> (lane * Pp)) & subP : 0u, g1 = lane < KVp ? (a1 >> (lane * Pp)) & subP : 0u;
        const unsigned gu = g0 | g1;
        unsigned sl = 0;
        for (int j = 0; j < Qp; ++j) sl |= (gu & need_of(phi, j)) ? (1u << j) : 0u;
        int incl = __popc(sl);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        const int pos = base + incl - __popc(sl);
        {
          int q = pos;
          for (unsigned rest = sl; rest; rest &= rest - 1, ++q) {
            const unsigned need = need_of(phi, __ffs(rest) - 1);
            if ((g0 & need) && f0 == 0x7fffffff) f0 = q;
            if ((g1 & need) && f1 == 0x7fffffff) f1 = q;
          }
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
        sl_ph[phi] = sl;
        g0_ph[phi] = g0;
        g1_ph[phi] = g1;
        pos_ph[phi] = pos;
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) {
        f0 = min(f0, __shfl_xor_sync(0xffffffffu, f0, d));
        f1 = min(f1, __shfl_xor_sync(0xffffffffu, f1, d));
      }
#pragma unroll
      for (int phi = 0; phi < 2; ++phi) {
        if (phi >= NPH) break;
        int pos = pos_ph[phi];
        for (unsigned rest = sl_ph[phi]; rest; rest &= rest - 1) {
          const int j = __ffs(rest) - 1;
          const unsigned need = need_of(phi, j);
          pl->stage[pos] = (unsigned short)(lane | (j << 5) | ((g0_ph[phi] & need) ? 1u << 9 : 0u) | ((g1_ph[phi] & need) ? 1u << 10 : 0u) |
                                            (pos == f0 ? 1u << 11 : 0u) | (pos == f1 ? 1u << 12 : 0u) | ((unsigned)phi << 13));
          ++pos;
        }
      }
      const int n_total = base;
      planned += n_total;
      TSG_STATE(st, n_total, t, (int)w.slot);
      if (lane == 0) {
        pl->tile = st;
        pl->n = n_total | (half << 16);
        pl->mask[0] = m0;
        pl->mask[1] = m1;
        pl->split = split;
        pl->slot = pslot;
      }
      __threadfence_block();  // every lane's stage entries are performed before lane 0 publishes the plan
      __syncwarp();
      if (lane == 0) mbar_arrive(sfull0 + 8 * w.slot);  // release: the plan is visible to the waiters
      if (lane == 0 && w.slot == 0 && w.phase == 0) TSG_LIFE(2);
      w.advance(V8_PLAN_SLOTS);
      if (st < 0) break;
    }
    if (lane == 0 && p.sched && !PAIR) {  // every CTA draws exactly one terminal ticket: the last one re-arms the counters for the next launch
      __threadfence();
      if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
        p.sched[0] = 0;
        p.sched[1] = 0;
        __threadfence();
      }
    }
    __syncwarp();
  } else {
    // ================================================================= producers
    // V8_GROUPS independent groups of V8_GROUP_WARPS warps; group g owns the global stages s with s % NG == g, so a
    // hand-shake (empty wait, landing wait, proxy fence, one arrival per warp) is paid once per 8 G copies per thread,
    // and while one group waits for its rows to land the others are issuing theirs.  A warp covers 32 consecutive tile
    // rows (four 8-lane row groups x 8 rows).  While the rows of stage n land, the warp requests the 128-byte index
    // lines of ITS next stage with cp.async into a private shared-memory buffer; they are complete at the same
    // cp.async.wait_all and are read back with two LDS.128 per sub-tile.
    const int pw = warp - V8_PROD_WARP0;
    const int grp = pw / V8_GROUP_WARPS, wg = pw % V8_GROUP_WARPS;
    const int NG = (int)nst < V8_GROUPS ? (int)nst : V8_GROUPS;
    const int chunk = lane & 7, rsl = lane >> 3;  // 8 lanes cover one 128-byte tile row; row group rsl of the warp
    const int rsub = wg * 4 + rsl;                // thread owns tile rows 8 rsub .. 8 rsub + 7
    uint32_t dst_off[V8_Q];                       // row 8 rsub + q sits at chunk position chunk ^ q of its 128-byte line
#pragma unroll
    for (int q = 0; q < V8_Q; ++q) dst_off[q] = b_bytes + (uint32_t)(rsub * V8_Q + q) * 128u + (uint32_t)((chunk ^ q) << 4);
    const uint32_t ibuf = idx0 + (uint32_t)pw * 2u * idx_warp_bytes;  // double buffered
    const long long n_out = n_rows;
    // ---- cursor over the global stage sequence: plan of the current super tile + index of the next stage in it
    int st = -2, n = 0, i = 0, gs = 0;  // gs = (global stage number of plan index i) mod NG
    unsigned masks[2] = {0u, 0u};       // offset masks of the super tile's sub-tiles (phase 0)
    unsigned lives = 0;                 // bit g: sub-tile g exists (the one offset of the shortcut phase)
    unsigned d_cur = 0;                 // descriptor of the stage the cursor delivered last
    // Moves the cursor to this group's next stage: 1 = delivered (d_cur), 0 = no more work, 2 = the next plan is not
    // published yet and `block` is false.  The non-blocking form exists because a producer must never wait for a plan
    // while it owes an arrival on a full barrier: the planner waits for the MMA issuers to release old plans, and they
    // wait for that arrival (a deadlock v9 hit as soon as tiles became one or two stages long).
    auto advance_mine = [&](bool block) -> int {
      for (;;) {
        if (st >= 0) {
          int skip = grp - gs;
          if (skip < 0) skip += NG;
          if (i + skip < n) {
            i += skip;
            d_cur = plans[pr.slot].stage[i];
            ++i;
            gs = grp + 1 == NG ? 0 : grp + 1;
            return 1;
          }
          gs = (gs + (n - i)) % NG;
          plan_release_warp();
          st = -2;  // between plans
        }
        if (st == -1) return 0;
        if (!block) {
          const int ready = __shfl_sync(0xffffffffu, (int)mbar_test(sfull0 + 8 * pr.slot, pr.phase), 0);
          if (!ready) return 2;
        }
        const volatile Plan *pl = plan_wait();
        st = pl->tile;
        if (st < 0) {
          plan_release_warp();
          st = -1;
          return 0;
        }
        n = pl->n & 0xffff;
        masks[0] = pl->mask[0];
        masks[1] = pl->mask[1];
        lives = (st * GP < num_tiles ? 1u : 0u) | ((GP > 1 && st * GP + 1 < num_tiles) ? 2u : 0u);
        if (PAIR) {   // this CTA's tile only
          masks[0] = masks[my_g];
          lives = (lives >> my_g) & 1u;
        }
        i = 0;
        TSG_STATE(st, n, gs, -1);
      }
    };
    if (grp >= NG) {
      for (;;) {  // fewer stages than groups: this group only keeps the plan ring moving
        const int t = plan_wait()->tile;
        plan_release_warp();
        if (t < 0) break;
      }
    } else {
      // ---- this lane's view of a stage (locate() fills it from the delivered descriptor)
      struct View {
        long long m0;       // first tile row of the super tile
        const char *src;    // source tensor + byte offset of the lane's chunk inside a source row
        uint32_t rb;        // source row pitch
        uint32_t ioff;      // this lane's 32 bytes inside the warp's index buffer (sub-tile 0)
        unsigned kbits;     // per sub-tile: the lane's offset is present
        unsigned act;       // per sub-tile: the slice is needed
        bool indexed;       // neighbour indices come from an index line (false: identity map)
      };
      auto locate = [&](View &v) {
        const unsigned phi = SC ? (d_cur >> 13) & 1u : 0u;
        const int kv = d_cur & 31u, j = (d_cur >> 5) & 15u;
        v.act = PAIR ? 1u : (d_cur >> 9) & 3u;   // PAIR: a tile that does not need the stage still presents (zero) rows
        const uint32_t e = lut[phi * 128 + j * 8 + chunk];
        const int k = kv * (phi ? P1 : P0) + (int)((e >> 2) & 3u);
        const bool second = (e >> 4) & 1u;
        const __nv_bfloat16 *base = phi ? (second ? p.ph[1].in1 : p.ph[1].in0) : (second ? p.ph[0].in1 : p.ph[0].in0);
        v.src = reinterpret_cast<const char *>(base) + (e >> 5) * 16u;
        v.rb = (uint32_t)(phi ? (second ? p.ph[1].c1 : p.ph[1].c0) : (second ? p.ph[0].c1 : p.ph[0].c0)) * 2u;
        v.ioff = (e & 3u) * (uint32_t)(G * 128) + (uint32_t)rsl * 32u;
        v.m0 = ((long long)st * GP + my_g) * TC_BM;
        v.kbits = phi ? (k == 0 ? lives : 0u) : (((masks[0] >> k) & 1u) | (((masks[1] >> k) & 1u) << 1));
        v.indexed = (phi ? p.ph[1].nbr : p.ph[0].nbr) != nullptr;
      };
      // request the index lines of the delivered stage: per (offset of the slice, sub-tile) the warp's 32 rows = 128 B
      auto prefetch_idx = [&](uint32_t buf) {
        const unsigned phi = SC ? (d_cur >> 13) & 1u : 0u;
        const int *nbr = phi ? p.ph[1].nbr : p.ph[0].nbr;
        if (!nbr) return;
        const long long nbr_stride = phi ? p.ph[1].nbr_stride : p.ph[0].nbr_stride;
        const int kv = d_cur & 31u, j = (d_cur >> 5) & 15u;
        const unsigned need = need_of(phi, j);
        const int k0 = kv * (phi ? P1 : P0) + __ffs(need) - 1, nks = __popc(need);
        for (int t = lane; t < nks * G * 8; t += 32) {
          const int ks = t / (G * 8), g = (t >> 3) % G, piece = t & 7;
          const int k = k0 + ks;
          const bool present = phi ? (k == 0 && ((lives >> g) & 1u)) : (((g ? masks[1] : masks[0]) >> k) & 1u);
          if (present)
            cp_async16(buf + (uint32_t)((ks * G + g) * 128 + piece * 16),
                       nbr + (long long)k * nbr_stride + ((long long)st * GP + my_g) * TC_BM + g * TC_BM + wg * 32 + piece * 4, 16u);
        }
      };
      // Nothing in this loop waits for gathered rows to land: a slot stays occupied only from the first copy to the
      // MMA's commit, and the (long, latency-bound) search for the next stage overlaps the landing without delaying the
      // hand-off.  Each thread's copies of a stage are followed by cp.async.mbarrier.arrive.noinc on the stage's full
      // barrier (the arrival fires when they have landed; the MMA thread issues the generic->async proxy fence after
      // its wait).  Commit groups per thread, in order: idx(n), rows(n-1), idx(n+1), rows(n); cp.async.wait_group 2 at
      // the top of iteration n guarantees idx(n).  The look-ahead never blocks on the plan ring (see advance_mine): if
      // the next plan is not published yet, the stage in hand is issued first and the look-ahead is redone, blocking.
      uint32_t slot = (uint32_t)grp, phase = 0;  // ring position of this group's next stage (always NG stages further)
      int n_issued = 0;
      uint32_t ib = 0;                           // which half of the warp's index buffer holds the current stage
      View cur, nxt;
      bool have = advance_mine(true) == 1;
      if (pw == 0 && lane == 0) TSG_LIFE(3);
      if (have) {
        locate(cur);
        prefetch_idx(ibuf);
      }
      nxt = cur;
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");  // stands in for rows(-1)
      while (have) {
        __syncwarp();  // every lane has read the indices that lived in the other half of the buffer
        int r = advance_mine(false);
        if (r == 1) {
          locate(nxt);
          prefetch_idx(ibuf + (ib ^ 1u) * idx_warp_bytes);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (pw == 0 && lane == 0) TSG_TRACE(7, n_issued);
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();
        int idx[G][V8_Q];  // neighbour row of tile row 8 rsub + q, or -1
        const uint32_t ibase = ibuf + ib * idx_warp_bytes + cur.ioff;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if ((cur.kbits >> g) & 1u) {
            if (cur.indexed) {
              const uint4 a = lds128(ibase + g * 128), b = lds128(ibase + g * 128 + 16);
              idx[g][0] = (int)a.x; idx[g][1] = (int)a.y; idx[g][2] = (int)a.z; idx[g][3] = (int)a.w;
              idx[g][4] = (int)b.x; idx[g][5] = (int)b.y; idx[g][6] = (int)b.z; idx[g][7] = (int)b.w;
            } else {  // identity map (1x1x1 convolutions, point MLPs)
              const long long r0 = cur.m0 + g * TC_BM + rsub * V8_Q;
#pragma unroll
              for (int q = 0; q < V8_Q; ++q) idx[g][q] = r0 + q < n_out ? (int)(r0 + q) : -1;
            }
          } else {
#pragma unroll
            for (int q = 0; q < V8_Q; ++q) idx[g][q] = -1;
          }
        }
        TSG_STATE(st, n, i, n_issued);
        mbar_wait(empty0 + 8 * slot, phase ^ 1);
        if (pw == 0 && lane == 0) TSG_TRACE(0, n_issued);
        const uint32_t dst = smem_base + slot * stage_bytes;
        if (!TSG_DBG(1)) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (!((cur.act >> g) & 1u)) continue;
#pragma unroll
            for (int q = 0; q < V8_Q; ++q) {
              const int v = idx[g][q];
              cp_async16(dst + g * TC_A_BYTES + dst_off[q], cur.src + (unsigned long long)(unsigned)max(v, 0) * cur.rb, v >= 0 ? 16u : 0u);
            }
          }
        }
        cp_async_arrive(full0 + 8 * slot);  // fires once this thread's copies (all of them so far) have landed
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (pw == 0 && lane == 0) TSG_TRACE(6, n_issued);
        if (pw == 0 && lane == 0) TSG_TRACE(1, n_issued);
        ++n_issued;
        slot += NG;
        if (slot >= nst) {
          slot -= nst;
          phase ^= 1;
        }
        if (r == 2) {  // the next plan was not there yet: nothing is owed any more, so wait for it now
          r = advance_mine(true);
          if (r == 1) {
            locate(nxt);
            prefetch_idx(ibuf + (ib ^ 1u) * idx_warp_bytes);
          }
          asm volatile("cp.async.wait_all;" ::: "memory");
        }
        cur = nxt;
        ib ^= 1u;
        have = r == 1;
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TSG_LIFE(5);
  if (threadIdx.x == 0) TSG_CTA(1, gtime());
  if (PAIR) {   // neither CTA leaves (or frees TMEM) while the other may still read its shared memory or arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == V8_MMA_WARP) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// K-slice packing of a layer with C = c0 + c1 input channels (both multiples of 16): the flat chunk stream repeats every
// P offsets with Q slices; need[j] = which of the P offsets slice j touches.
struct SlicePlan {
  int P, Q, cpo, ksmax;  // ksmax = most kernel offsets any one slice touches
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
  sp.ksmax = 1;
  for (int j = 0; j < sp.Q && j < 16; ++j) {
    const int lo = (8 * j) / sp.cpo, hi = (8 * j + 7) / sp.cpo;
    if (hi - lo + 1 > sp.ksmax) sp.ksmax = hi - lo + 1;
    unsigned long long bits = 0;
    for (int k = lo; k <= hi; ++k) bits |= 1ull << k;
    sp.need |= bits << (4 * j);
  }
  return sp;
}

// W (K, c_in, c_out) fp32 -> per (virtual offset kv, slice j) a [c_out][64] bf16 block, K-major, 128B-swizzled: the byte
// image the MMA reads, so a linear bulk copy stages it.  Column 8 c + e of block (kv, j) is channel 8 cc + e of offset
// kv P + ksub with 8 j + c = ksub cpo + cc.  Optional per-output-channel scale (folded BatchNorm).
__global__ void pack_weights_kernel(const float *__restrict__ w, int K, int c_in, int c_out, long long sk, long long sci,
                                    long long sco, int P, int Q, int cpo, const float *__restrict__ out_scale,
                                    __nv_bfloat16 *__restrict__ packed) {   // sk / sci / sco: element strides of W (a transposed or sliced view packs without a copy)
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
    if (k < K && ch < c_in) v = w[k * sk + ch * sci + n * sco];
    if (out_scale) v *= out_scale[n];
    const long long blk = ((long long)kv * Q + j) * c_out * TC_KB;
    const int sw = (((col >> 3) ^ (n & 7)) << 3) | (col & 7);
    packed[blk + (long long)n * TC_KB + sw] = __float2bfloat16_rn(v);
  }
}

// Work-item list of a K-split launch.  One block; tiles are listed heaviest first (descending tile index of a mask-sorted
// map).  A tile with more than `cap` active offsets becomes ceil(active / cap) work items (at most max_parts) over disjoint,
// equally sized runs of its active offsets: the serial stage chain of the heaviest tile is what a launch with fewer tiles than
// SMs waits for (stride-16 level of the benchmark: 120 tiles, 9.2 active offsets per SM on average but 27 on the critical path).
// At most max_slots tiles are split.
__global__ void __launch_bounds__(1024, 1) conv_split_items_kernel(const unsigned *__restrict__ tile_mask, long long n_out,
                                                                 const int *__restrict__ n_out_dev, int K, int cap, int max_parts,
                                                                 int max_slots, int4 *__restrict__ items, int *__restrict__ n_items) {
  __shared__ int scan_a[1024], scan_b[1024];
  const long long n_rows = dev_count(n_out_dev, n_out);
  const int num_tiles = (int)((n_rows + TC_BM - 1) / TC_BM);
  const unsigned kmask = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
  const int per = (num_tiles + 1023) / 1024;
  const int i0 = threadIdx.x * per, i1 = min(num_tiles, i0 + per);   // positions in heavy-first order: tile = num_tiles - 1 - i
  auto parts_of = [&](int x) -> int { return x > cap ? min(max_parts, (x + cap - 1) / cap) : 1; };
  int nsplit = 0;
  for (int i = i0; i < i1; ++i) nsplit += parts_of(__popc(tile_mask[num_tiles - 1 - i] & kmask)) > 1 ? 1 : 0;
  // exclusive block scans (Hillis-Steele over 1024 partial sums): split tiles before this thread, then items
  auto block_scan = [&](int *buf, int v) -> int {
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      const int o = threadIdx.x >= d ? buf[threadIdx.x - d] : 0;
      __syncthreads();
      buf[threadIdx.x] += o;
      __syncthreads();
    }
    return buf[threadIdx.x] - v;
  };
  int sbase = block_scan(scan_a, nsplit);
  int nitems = 0;
  {
    int sb = sbase;
    for (int i = i0; i < i1; ++i) {
      const int np = parts_of(__popc(tile_mask[num_tiles - 1 - i] & kmask));
      nitems += (np > 1 && sb < max_slots) ? np : 1;
      if (np > 1) ++sb;
    }
  }
  int ibase = block_scan(scan_b, nitems);
  for (int i = i0; i < i1; ++i) {
    const int tile = num_tiles - 1 - i;
    const unsigned m = tile_mask[tile] & kmask;
    const int x = __popc(m), np = parts_of(x);
    if (np > 1 && sbase < max_slots) {
      unsigned rest = m;
      for (int q = 0; q < np; ++q) {            // part q: active offsets [q x / np, (q + 1) x / np) in ascending order
        unsigned mine = 0;
        for (int b = q * x / np; b < (q + 1) * x / np; ++b) {
          mine |= rest & (0u - rest);
          rest &= rest - 1;
        }
        items[ibase++] = make_int4(tile, (int)mine, q | (np << 8), sbase);
      }
    } else {
      items[ibase++] = make_int4(tile, (int)m, 0 | (1 << 8), 0);
    }
    if (np > 1) ++sbase;
  }
  if (threadIdx.x == 1023) *n_items = scan_b[1023];
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
  return tsg_conv_pack_weights2(weight, k, c_in, c_out, (int64_t)c_in * c_out, c_out, 1, c0, c1, out_scale, packed, stream);
}

int tsg_conv_pack_weights2(const float *weight, int k, int c_in, int c_out, int64_t stride_k, int64_t stride_cin,
                           int64_t stride_cout, int c0, int c1, const float *out_scale, void *packed, tsg_stream_t stream) {
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
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, stream>>>(weight, k, c_in, c_out, stride_k, stride_cin, stride_cout, sp.P, sp.Q,
                                                                sp.cpo, out_scale, (__nv_bfloat16 *)packed);
  return check_launch("tsg_conv_pack_weights");
}

/* profiling aid, not part of the public header: copies the trace of the last TSG_TC_DEBUG&128 launch to `host` (13 x 96 int64) */
int tsg_debug_conv_trace(long long *host) {
  TSG_CUDA(cudaDeviceSynchronize());
  TSG_CUDA(cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * 13 * TRACE_N));
  return TSG_OK;
}

/* ... per-CTA {entry ns, exit ns, stages of MMA warp 0, tiles} of the same launch (160 x 4 int64) */
int tsg_debug_conv_ctas(long long *host) {
  TSG_CUDA(cudaDeviceSynchronize());
  TSG_CUDA(cudaMemcpyFromSymbol(host, g_cta, sizeof(long long) * 160 * 4));
  return TSG_OK;
}

/* ... and CTA 0's life-cycle stamps of the same launch (8 int64) */
int tsg_debug_conv_life(long long *host) {
  TSG_CUDA(cudaDeviceSynchronize());
  TSG_CUDA(cudaMemcpyFromSymbol(host, g_life, sizeof(long long) * 8));
  return TSG_OK;
}

static int fill_phase(TcPhase &h, const void *in0, int c0, const void *in1, int c1, const void *packed_w, int k,
                      const int32_t *nbr, int64_t nbr_stride, int *ksmax, const char *what) {
  const SlicePlan sp = slice_plan(c0, c1);
  if (sp.Q > 16) {
    set_error("%s: unsupported channel count (more than 16 slices per offset group)", what);
    return TSG_ERR_UNSUPPORTED;
  }
  h.in0 = (const __nv_bfloat16 *)in0;
  h.in1 = (const __nv_bfloat16 *)in1;
  h.packed_w = (const uint8_t *)packed_w;
  h.nbr = nbr;
  h.nbr_stride = nbr_stride;
  h.slice_need = sp.need;
  h.c0 = c0;
  h.c1 = c1;
  h.pk = sp.P;
  h.kq = sp.Q;
  h.cpo = sp.cpo;
  h.K = k;
  if (sp.ksmax > *ksmax) *ksmax = sp.ksmax;
  return TSG_OK;
}

int tsg_conv_split_items(const uint32_t *tile_mask, int64_t n_out, const int32_t *n_out_dev, int k, int cap, int max_parts,
                         int max_slots, int32_t *items, int32_t *n_items, tsg_stream_t stream) {
  if (!tile_mask || !items || !n_items || k <= 0 || k > 32 || cap < 1 || max_parts < 2 || max_parts > 32 || max_slots < 0 ||
      n_out <= 0 || (n_out + TC_BM - 1) / TC_BM > 4096) {
    set_error("tsg_conv_split_items: need tile_mask, items, n_items, 0 < K <= 32, cap >= 1, 2 <= max_parts <= 32 and at most 4096 tiles");
    return TSG_ERR_INVALID;
  }
  conv_split_items_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(tile_mask, n_out, n_out_dev, k, cap, max_parts, max_slots,
                                                                (int4 *)items, n_items);
  return check_launch("tsg_conv_split_items");
}

int tsg_conv_fwd_tc4(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                     int c_out, const int32_t *nbr, int64_t nbr_stride, const uint32_t *tile_mask, const int32_t *perm,
                     int64_t n_out, const int32_t *n_out_dev, const void *sc_in0, int sc_c0, const void *sc_in1, int sc_c1,
                     const void *sc_packed_w, const int32_t *sc_idx, void *out, int out_dtype, const float *bias,
                     const void *residual, int relu, int num_sms_hint, int32_t *sched, const int32_t *items,
                     const int32_t *n_items, int max_slots, int max_parts, float *split_scratch, int32_t *split_state,
                     tsg_stream_t stream) {
  if (c0 % 16 || c1 % 16 || c_out % 16 || c_out > 256 || c_out <= 0 || c0 <= 0 || k <= 0 || k > 32 ||
      (out_dtype != TSG_BF16 && out_dtype != TSG_F32)) {
    set_error("tsg_conv_fwd_tc: need c0,c1,c_out multiples of 16, c_out<=256, K<=32, out bf16/f32");
    return TSG_ERR_UNSUPPORTED;
  }
  if (sc_in0 && (sc_c0 % 16 || sc_c1 % 16 || sc_c0 <= 0 || !sc_packed_w || (perm && !sc_idx))) {
    set_error("tsg_conv_fwd_tc: shortcut needs sc_c0,sc_c1 multiples of 16, packed weights and (with perm) its index line");
    return TSG_ERR_INVALID;
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
  TcParams p;
  memset(&p, 0, sizeof(p));
  int ksmax = 1;
  int rc = fill_phase(p.ph[0], in0, c0, in1, c1, packed_w, k, nbr, nbr_stride, &ksmax, "tsg_conv_fwd_tc");
  if (rc != TSG_OK) return rc;
  p.n_phases = 1;
  p.ph[1] = p.ph[0];
  if (sc_in0) {
    rc = fill_phase(p.ph[1], sc_in0, sc_c0, sc_in1, sc_c1, sc_packed_w, 1, sc_idx, nbr_stride, &ksmax, "tsg_conv_fwd_tc (shortcut)");
    if (rc != TSG_OK) return rc;
    p.n_phases = 2;
    const int stages0 = (k + p.ph[0].pk - 1) / p.ph[0].pk * p.ph[0].kq;
    if (stages0 + p.ph[1].kq > V8_PLAN_MAX) {
      set_error("tsg_conv_fwd_tc: too many pipeline stages per tile");
      return TSG_ERR_UNSUPPORTED;
    }
  }
  p.c_out = c_out;
  p.tile_mask = tile_mask;
  p.perm = perm;
  p.n_out = n_out;
  p.n_out_dev = n_out_dev;
  p.out = out;
  p.out_f32 = out_dtype == TSG_F32;
  p.bias = bias;
  p.residual = (const __nv_bfloat16 *)residual;
  p.relu = relu;
  p.sched = sched;
  if (items) {
    if (!n_items || !sched || (max_slots > 0 && (!split_scratch || !split_state || max_parts < 2))) {
      set_error("tsg_conv_fwd_tc: a work-item list needs n_items, the scheduler counters and (max_slots > 0) scratch + state + max_parts >= 2");
      return TSG_ERR_INVALID;
    }
    p.items = (const int4 *)items;
    p.n_items = n_items;
    p.split_scratch = split_scratch;
    p.split_state = split_state;
    p.split_parts = max_parts;
  }
#ifdef TSG_TC_TRACE  // profiling knock-outs (trace build only, wrong results): 1 no gathers, 2 no weight copies,
  const char *dbg_env = getenv("TSG_TC_DEBUG");  // 4 no MMAs, 8 no epilogue stores, 128 trace; re-read on every launch
  p.dbg = dbg_env ? atoi(dbg_env) : 0;
#else
  p.dbg = 0;
#endif
  int dev = 0;
  TSG_CUDA(cudaGetDevice(&dev));
  int sms = num_sms_hint;
  if (sms <= 0) TSG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long num_tiles = (n_out + TC_BM - 1) / TC_BM;
  // N split (TSG_TC_NSPLIT=2, off by default): launches with fewer tiles than SMs hand out (tile, column half) work
  // items, halving the serial stage chain of the heaviest tile.  Measured on the 256->256 layers at stride 16 (120 tiles):
  // 52 -> 63 us — a stage costs about the same whatever its width (the pipeline is hand-shake bound, profiles/README.md), so
  // twice the stages on 148 instead of 120 SMs is a loss.  Kept as a tested switch for wider / shallower launches.
  const char *ns_env = getenv("TSG_TC_NSPLIT");
  const int ns = (!items && ns_env && atoi(ns_env) == 2 && c_out >= 64 && c_out % 32 == 0 && num_tiles <= sms) ? 2 : 1;
  p.ns = ns;
  p.n_eff = c_out / ns;
  // G sub-tiles share every weight slice; bounded by TMEM (2 buffers x G x n_eff fp32 columns <= 512) and by the
  // number of super tiles needed to keep every SM busy
  int G = p.n_eff <= 128 ? 2 : 1;  // (G = 4 spills the producers' index registers at 768 threads per CTA)
  while (G > 1 && (num_tiles + G - 1) / G * ns < 2LL * sms) G >>= 1;
  if (items) G = 1;   // work items address single tiles
  // CTA pairs (cta_group::2, M = 256 over two SMs, half a weight slice per CTA): TSG_TC_PAIR = 0 never (default: measured no faster, see profiles/README.md), 1 the
  // launches that would otherwise run one tile per CTA with c_out >= 128 (L2 -> SM weight traffic bound), 2 every launch
  // with c_out >= 64.  Not combined with work-item lists or the column split.
  static const int pair_mode = getenv("TSG_TC_PAIR") ? atoi(getenv("TSG_TC_PAIR")) : 0;
  const bool pair = ns == 1 && num_tiles >= 2 && sms >= 2 &&
                    ((pair_mode == 1 && G == 1 && c_out >= 128) || (pair_mode >= 2 && c_out >= 64));
  if (pair) {   // takes precedence over a work-item list (the K split bought 4 % where pairs buy far more)
    G = 1;
    items = nullptr;
    p.items = nullptr;
  }
#ifdef TSG_TC_TRACE
  if (getenv("TSG_TC_G1")) G = 1;
#endif
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)G * (uint32_t)p.n_eff) cols <<= 1;
  p.tmem_cols = cols;
  p.ksmax = ksmax;
  // dynamic shared memory: 1 KB alignment slack + stages + epilogue staging + the producers' index buffers
  const size_t b_bytes = (size_t)p.n_eff * (pair ? 64 : 128), stage_bytes = b_bytes + (size_t)G * TC_A_BYTES;
  const size_t fixed = 1024 + (size_t)V8_EPI_WARPS * V8_STG_BYTES + (size_t)V8_PROD_WARPS * 2 * ksmax * G * 128;
  int stages = (int)((V8_DYN_SMEM - fixed) / stage_bytes);
  if (stages > V8_MAX_STAGES) stages = V8_MAX_STAGES;
#ifdef TSG_TC_TRACE
  if (const char *e = getenv("TSG_TC_STAGES")) stages = atoi(e) < stages ? atoi(e) : stages;
#endif
  if (stages < 2) {
    set_error("tsg_conv_fwd_tc: not enough shared memory for the pipeline");
    return TSG_ERR_UNSUPPORTED;
  }
  p.na = stages;
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  static bool configured[64] = {false};   // the >48 KB shared-memory opt-in is per device
  if (dev < 0 || dev >= 64 || !configured[dev]) {
#define TSG_TC_SMEM(...) TSG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, V8_DYN_SMEM))
#define TSG_TC_SMEM4(g, pair, thr, items)                                                                                   \
  TSG_TC_SMEM(g, pair, thr, false, false, items); TSG_TC_SMEM(g, pair, thr, true, false, items);                            \
  TSG_TC_SMEM(g, pair, thr, false, true, items); TSG_TC_SMEM(g, pair, thr, true, true, items)
    TSG_TC_SMEM4(1, false, false, false); TSG_TC_SMEM4(1, false, true, false); TSG_TC_SMEM4(2, false, false, false);
    TSG_TC_SMEM4(2, false, true, false); TSG_TC_SMEM4(1, true, false, false); TSG_TC_SMEM4(1, false, true, true);
    TSG_TC_SMEM4(1, false, false, true);
#undef TSG_TC_SMEM4
#undef TSG_TC_SMEM
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const long long work = items ? num_tiles + (long long)max_slots * (max_parts - 1) : pair ? (num_tiles + 1) / 2 : (num_tiles + G - 1) / G * ns;
  // programmatic dependent launch (TSG_TC_PDL=1, off by default): this grid's CTAs may be scheduled while the previous
  // kernel of the stream drains; the kernel executes griddepcontrol.wait before it reads anything.  Measured: +1 % with one
  // batch in flight, -8 % with two (early CTAs of one stream's next convolution hold the SMs the other stream's small
  // kernels would have used), so the benchmark configuration leaves it off.
  static const bool pdl = getenv("TSG_TC_PDL") && atoi(getenv("TSG_TC_PDL")) == 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = pair ? dim3(2u * (unsigned)(work < sms / 2 ? work : sms / 2)) : dim3((unsigned)(work < sms ? work : sms));
  cfg.blockDim = dim3(V8_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  // look-ahead throttle (see the planner): launches with at most `thr_max` work items per CTA
  static const int thr_max = getenv("TSG_TC_THROTTLE") ? atoi(getenv("TSG_TC_THROTTLE")) : 8;
  static const int lookahead = getenv("TSG_TC_LOOKAHEAD") ? atoi(getenv("TSG_TC_LOOKAHEAD")) : 6;
  p.lookahead = lookahead;
  const bool thr = sched && work <= (long long)thr_max * cfg.gridDim.x;
  const bool sc = p.n_phases > 1, f32 = p.out_f32 != 0;
#define TSG_TC_GO(...) TSG_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<__VA_ARGS__>, p))
#define TSG_TC_GO4(g, pair, thr, items)                                                  \
  do {                                                                                   \
    if (sc && f32) TSG_TC_GO(g, pair, thr, true, true, items);                           \
    else if (sc) TSG_TC_GO(g, pair, thr, true, false, items);                            \
    else if (f32) TSG_TC_GO(g, pair, thr, false, true, items);                           \
    else TSG_TC_GO(g, pair, thr, false, false, items);                                   \
  } while (0)
  if (pair) TSG_TC_GO4(1, true, false, false);
  else if (items && thr) TSG_TC_GO4(1, false, true, true);
  else if (items) TSG_TC_GO4(1, false, false, true);
  else if (G == 2 && thr) TSG_TC_GO4(2, false, true, false);
  else if (G == 2) TSG_TC_GO4(2, false, false, false);
  else if (thr) TSG_TC_GO4(1, false, true, false);
  else TSG_TC_GO4(1, false, false, false);
#undef TSG_TC_GO4
#undef TSG_TC_GO
  return check_launch("tsg_conv_fwd_tc");
}

int tsg_conv_fwd_tc3(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                     int c_out, const int32_t *nbr, int64_t nbr_stride, const uint32_t *tile_mask, const int32_t *perm,
                     int64_t n_out, const int32_t *n_out_dev, const void *sc_in0, int sc_c0, const void *sc_in1, int sc_c1,
                     const void *sc_packed_w, const int32_t *sc_idx, void *out, int out_dtype, const float *bias,
                     const void *residual, int relu, int num_sms_hint, int32_t *sched, tsg_stream_t stream) {
  return tsg_conv_fwd_tc4(in0, c0, in1, c1, n_in, packed_w, k, c_out, nbr, nbr_stride, tile_mask, perm, n_out, n_out_dev, sc_in0,
                          sc_c0, sc_in1, sc_c1, sc_packed_w, sc_idx, out, out_dtype, bias, residual, relu, num_sms_hint, sched,
                          nullptr, nullptr, 0, 0, nullptr, nullptr, stream);
}

int tsg_conv_fwd_tc2(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                     int c_out, const int32_t *nbr, int64_t nbr_stride, const uint32_t *tile_mask, const int32_t *perm,
                     int64_t n_out, const void *sc_in0, int sc_c0, const void *sc_in1, int sc_c1,
                     const void *sc_packed_w, const int32_t *sc_idx, void *out, int out_dtype, const float *bias,
                     const void *residual, int relu, int num_sms_hint, int32_t *sched, tsg_stream_t stream) {
  return tsg_conv_fwd_tc3(in0, c0, in1, c1, n_in, packed_w, k, c_out, nbr, nbr_stride, tile_mask, perm, n_out, nullptr, sc_in0,
                          sc_c0, sc_in1, sc_c1, sc_packed_w, sc_idx, out, out_dtype, bias, residual, relu, num_sms_hint, sched,
                          stream);
}

int tsg_conv_fwd_tc(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                    int c_out, const int32_t *nbr, int64_t nbr_stride, const uint32_t *tile_mask, const int32_t *perm,
                    int64_t n_out, void *out, int out_dtype, const float *bias, const void *residual, int relu,
                    int num_sms_hint, int32_t *sched, tsg_stream_t stream) {
  return tsg_conv_fwd_tc2(in0, c0, in1, c1, n_in, packed_w, k, c_out, nbr, nbr_stride, tile_mask, perm, n_out, nullptr, 0,
                          nullptr, 0, nullptr, nullptr, out, out_dtype, bias, residual, relu, num_sms_hint, sched, stream);
}

}  // extern "C"
