// Weight gradient of the sparse convolution on the 5th-generation tensor cores (sm_100a):
//   grad_w[k] (c_in x c_out, fp32) = sum over the pairs (i, o) of kernel offset k of  x[i,:]^T  gy[o,:],
// bf16 operands, fp32 accumulation in TMEM.  Replaces the gather + `torch::mm_out(grad_kernel[k], in_buffer^T, grad_out_buffer)`
// of the reference's backward (TS/backend/convolution/convolution_cuda.cu:167-278) for the autocast training path.
//
// The GEMM of one offset has M = c_in, N = c_out and the PAIR LIST as its K dimension.  A gathered row in shared memory —
// 64 channels = 128 bytes per pair, eight pairs per 1024-byte swizzle atom, exactly what the forward kernel's producers
// write — IS a canonical MN-major operand tile of tcgen05 (element (channel, pair): channels contiguous, pairs strided):
// x rows are the A operand (M-major), gy rows the B operand (N-major); no transposition anywhere.  The instruction
// descriptor selects MN-major for both operands (bits 15, 16); the shared-memory descriptors carry SBO = 1024 B (next eight
// pairs) and LBO = 8192 B (next 64-channel block: blocks are stored [block][64 pairs][128 B]).
//
// grid = (work units, ceil(c_in / 128)); 288 threads.  The pairs of every offset come as a compact list (tsg_kmap_pair_list,
// built once per kernel map): a work unit is up to WT_UNIT 64-pair chunks of ONE offset (TMEM accumulates one grad_w[k] tile),
// so heavy offsets (the centre has a pair for every row) simply get more units.  Eight gather warps walk the unit's
// chunks: lane j of warp w loads pair w + 8 j, the indices are broadcast by shuffles, and both operand rows are gathered
// with 16-byte cp.async into one of WT_STAGES stages, each thread's copies followed by cp.async.mbarrier.arrive.noinc on the
// stage's full barrier; nothing else synchronises the gatherers.  The ninth warp issues four K = 16 MMAs per stage
// (M = 128 input channels = TMEM lanes, N = all output channels <= 256 fp32 columns) and releases the stage with
// tcgen05.commit.  The epilogue adds the unit's tile to grad_w with 16-byte vector atomics straight from TMEM.
#include <cstdlib>

#include "tc_common.cuh"

namespace tsg {

constexpr int WT_PAIRS = 64;                 // pairs per stage = four K = 16 MMAs
constexpr int WT_STAGES = 3;
constexpr int WT_BLOCK = WT_PAIRS * 128;     // one 64-channel block of a stage: 8 KB
constexpr int WT_GATHER_WARPS = 8;
constexpr int WT_THREADS = 32 * (WT_GATHER_WARPS + 1);

struct WgradParams {
  const __nv_bfloat16 *in, *gy;
  const int2 *pairs;        // {in row, out row}, offset by offset
  const int *start;         // K + 1: first pair of every offset
  float *gw;
  int K, c_in, c_out, n_mma;   // n_mma: c_out rounded up to a multiple of 16 (the MMA's N)
  int nb;                   // 64-channel blocks of gy per stage
  int target;               // work units aimed at (the grid holds target + K)
  uint32_t tmem_cols;
  int dbg;   // TSG_WG_DEBUG knock-outs for profiling (wrong results): 1 no gathers, 2 no MMAs, 4 no epilogue stores
};

__device__ long long g_wg_prof[8];   // trace build, TSG_WG_DEBUG & 64: cycle totals of CTA 0 (gather: wait empty, issue; MMA: wait full, issue; chunks; total)
#ifdef TSG_TC_TRACE
#define WG_CLOCK() clock64()
#else
#define WG_CLOCK() 0ll
#endif

__global__ void __launch_bounds__(WT_THREADS) conv_wgrad_tc_kernel(const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * WT_STAGES + 1];
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- which unit.  A unit is up to U 64-pair chunks of one offset, U = ceil(chunks of the whole map / target): about
  // `target` units whatever the (device-side) pair count, at most target + K.  Every warp resolves it for itself: lane kk
  // holds offset kk's bounds, a warp scan of the unit counts finds the offset of unit blockIdx.x.
  int k, p0, p1;
  {
    const int lo = lane < p.K ? __ldg(p.start + lane) : 0, hi = lane < p.K ? __ldg(p.start + lane + 1) : 0;
    const int total = __shfl_sync(0xffffffffu, hi, p.K - 1);
    const int U = max(1, ((total + WT_PAIRS - 1) / WT_PAIRS + p.target - 1) / p.target);
    const int chunks = (hi - lo + WT_PAIRS - 1) / WT_PAIRS, units = (chunks + U - 1) / U;
    int incl = units;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const unsigned owner = __ballot_sync(0xffffffffu, (int)blockIdx.x >= incl - units && (int)blockIdx.x < incl);
    if (!owner) return;   // beyond the last unit (uniform: every warp computes the same)
    k = __ffs(owner) - 1;
    const int u = (int)blockIdx.x - (__shfl_sync(0xffffffffu, incl, k) - __shfl_sync(0xffffffffu, units, k));
    const int klo = __shfl_sync(0xffffffffu, lo, k), khi = __shfl_sync(0xffffffffu, hi, k);
    p0 = klo + u * U * WT_PAIRS;
    p1 = min(khi, p0 + U * WT_PAIRS);
  }
  const int n_chunks = (p1 - p0 + WT_PAIRS - 1) / WT_PAIRS;
  const int ci0 = blockIdx.y * 128;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)(2 + p.nb) * WT_BLOCK;     // [A block 0][A block 1][B block 0..nb-1]
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = full0 + 8 * WT_STAGES, done = empty0 + 8 * WT_STAGES;
  if (tid == 0) {
    for (int s = 0; s < WT_STAGES; ++s) {
      mbar_init(full0 + 8 * s, WT_GATHER_WARPS * 32);   // every gather thread, when its copies have landed
      mbar_init(empty0 + 8 * s, 1);                     // one tcgen05.commit
    }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int c_in = p.c_in, c_out = p.c_out;

  if (warp == WT_GATHER_WARPS) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      // M = 128, N = n_mma, bf16 x bf16 -> fp32, A and B MN-major
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.n_mma >> 3) << 17) |
                             ((128u >> 4) << 24);
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      const uint32_t lbo = (uint32_t)(WT_BLOCK >> 4) << 16;                       // LBO = one 64-channel block
      Ring r;
      long long t_wait = 0, t_issue = 0;
      for (int c = 0; c < n_chunks; ++c) {
        const long long t0 = WG_CLOCK();
        mbar_wait(full0 + 8 * r.slot, r.phase);
        const long long t1 = WG_CLOCK();
        t_wait += t1 - t0;
        fence_async_proxy();
        tc_fence_after();
        const uint32_t a0 = smem_base + r.slot * stage_bytes, b0 = a0 + 2 * WT_BLOCK;
#pragma unroll
        for (int ks = 0; ks < WT_PAIRS / 16; ++ks) {
          const uint32_t a_lo = (((a0 + ks * 2048u) & 0x3FFFFu) >> 4) | lbo, b_lo = (((b0 + ks * 2048u) & 0x3FFFFu) >> 4) | lbo;
          uint64_t da, db;
          asm("mov.b64 %0, {%1, %2};" : "=l"(da) : "r"(a_lo), "r"(desc_hi));
          asm("mov.b64 %0, {%1, %2};" : "=l"(db) : "r"(b_lo), "r"(desc_hi));
          if (!(p.dbg & 2)) umma_bf16(tmem_base, da, db, idesc, (c | ks) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * r.slot);
        r.advance(WT_STAGES);
        t_issue += WG_CLOCK() - t1;
      }
      umma_commit(done);   // every MMA of this unit has completed
      if ((p.dbg & 64) && blockIdx.x == 0 && blockIdx.y == 0) {
        g_wg_prof[2] = t_wait;
        g_wg_prof[3] = t_issue;
        g_wg_prof[4] = n_chunks;
      }
    }
    __syncwarp();
  } else {
    // ================================================================= gather warps
    // geometry fixed per thread: warp w copies pairs w, w + 8, ..., w + 56 of a chunk (their swizzle phase, r & 7, is the
    // warp's own), lane l the 16-byte piece l of a pair (and piece l + 32 when a pair has more than 32: 16 pieces of x,
    // 8 per 64-channel block of gy).  Nothing in the per-chunk loop divides or branches on the piece.
    const int per_pair = 16 + p.nb * 8;
    const char *g_base[2];
    uint32_t g_pitch[2], g_dst[2];
    bool g_on[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pi = lane + 32 * h;
      const bool isa = pi < 16;
      const int ch = isa ? ci0 + pi * 8 : (pi - 16) * 8;
      g_on[h] = pi < per_pair && (isa ? ch < c_in : ch < c_out) && !(p.dbg & 1);
      g_base[h] = reinterpret_cast<const char *>(isa ? p.in : p.gy) + (g_on[h] ? (size_t)ch * 2 : 0);
      g_pitch[h] = (uint32_t)(isa ? c_in : c_out) * 2u;
      const int blk = isa ? (pi >> 3) : 2 + ((pi - 16) >> 3);
      g_dst[h] = (uint32_t)blk * WT_BLOCK + (uint32_t)(((pi & 7) ^ (warp & 7)) << 4);
    }
    const bool isa0 = lane < 16, two = per_pair > 32;   // (isa0 is declared before the loader below uses it)
    // lane j < 8 holds the INPUT row of pair w + 8 j of the chunk, lane 8 + j its OUTPUT row (-1 beyond the unit): a piece
    // of x (lanes 0-15) fetches its row index from lane j, a piece of gy from lane 8 + j — one shuffle per pair (shuffles
    // share the load/store unit's issue port with the copies, and the copies' issue rate is what bounds this kernel)
    auto load_pair = [&](int c) -> int {
      const int e = p0 + c * WT_PAIRS + warp + 8 * (lane & 7);
      if (!(c < n_chunks && e < p1)) return -1;
      const int *q = reinterpret_cast<const int *>(p.pairs + e);
      return __ldg(q + ((lane >> 3) & 1));
    };
    const int src_lane = isa0 ? 0 : 8;
    int pr_next = load_pair(0), pr_next2 = load_pair(1);
    Ring r;
    long long t_wait = 0, t_issue = 0;
    const long long t_begin = WG_CLOCK();
    for (int c = 0; c < n_chunks; ++c) {
      const int pr = pr_next;
      pr_next = pr_next2;
      pr_next2 = load_pair(c + 2);         // indices travel two chunks ahead of the gather that uses them
      const long long t0 = WG_CLOCK();
      mbar_wait(empty0 + 8 * r.slot, r.phase ^ 1);
      const long long t1 = WG_CLOCK();
      t_wait += t1 - t0;
      const uint32_t st0 = smem_base + r.slot * stage_bytes + (uint32_t)warp * 128u;
#pragma unroll
      for (int j = 0; j < WT_PAIRS / 8; ++j) {
        const int idx = __shfl_sync(0xffffffffu, pr, j + src_lane);
        const uint32_t drow = st0 + (uint32_t)j * 1024u;      // pair w + 8 j: row (w + 8 j) * 128 B
        {
          const bool ok = idx >= 0 && g_on[0];
          if (lane < per_pair) cp_async16(drow + g_dst[0], g_base[0] + (ok ? (size_t)idx * g_pitch[0] : 0), ok ? 16u : 0u);
        }
        if (two) {   // pieces 32-47 are all gy
          const int i_out = __shfl_sync(0xffffffffu, pr, j + 8);
          const bool ok = i_out >= 0 && g_on[1];
          if (lane + 32 < per_pair) cp_async16(drow + g_dst[1], g_base[1] + (ok ? (size_t)i_out * g_pitch[1] : 0), ok ? 16u : 0u);
        }
      }
      cp_async_arrive(full0 + 8 * r.slot);   // fires once this thread's copies have landed
      r.advance(WT_STAGES);
      t_issue += WG_CLOCK() - t1;
    }
    if ((p.dbg & 64) && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
      g_wg_prof[0] = t_wait;
      g_wg_prof[1] = t_issue;
      g_wg_prof[5] = WG_CLOCK() - t_begin;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");

    // ---- epilogue: TMEM lane = input channel ci0 + lane, column = output channel.  Warps 0-3 and 4-7 split the columns.
    mbar_wait(done, 0);
    tc_fence_after();
    if (!(p.dbg & 4)) {
      const int quad = warp & 3, half = warp >> 2;
      const int ci = ci0 + quad * 32 + lane;
      const int n16 = p.n_mma / 16;
      for (int cb = half; cb < n16; cb += 2) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb * 16u, v);
        if (ci < c_in) {
          float *dst = p.gw + ((long long)k * c_in + ci) * c_out + cb * 16;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int co = cb * 16 + 4 * j;
            if (co + 3 < c_out) {
              atomicAdd(reinterpret_cast<float4 *>(dst + 4 * j), make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                              __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
            } else {
              for (int e = 0; e < 4; ++e)
                if (co + e < c_out) atomicAdd(dst + 4 * j + e, __uint_as_float(v[4 * j + e]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
}

}  // namespace tsg

using namespace tsg;

extern "C" {

/* bf16 operands (c_in, c_out multiples of 8, c_out <= 256), fp32 grad_w (k, c_in, c_out): the autocast training path on
 * tcgen05 (see the header of this file).  Unsupported shapes return TSG_ERR_UNSUPPORTED (the caller picks another kernel). */
int tsg_conv_wgrad_tc(const void *in, int64_t n_in, int c_in, const void *grad_out, int64_t n_out, int c_out,
                      const int32_t *pairs, const int32_t *start, int64_t pair_cap, int k, float *grad_w, tsg_stream_t stream) {
  (void)n_in;
  (void)n_out;
  if (c_in % 8 || c_out % 8 || c_in <= 0 || c_out <= 0 || c_out > 256 || k <= 0 || k > 32 || !pairs || !start) {
    set_error("tsg_conv_wgrad_tc: c_in and c_out must be multiples of 8, c_out <= 256, K <= 32, and a pair list is required");
    return TSG_ERR_UNSUPPORTED;
  }
  TSG_CUDA(cudaMemsetAsync(grad_w, 0, (size_t)k * c_in * c_out * sizeof(float), stream));
  if (pair_cap <= 0) return TSG_OK;
  WgradParams p;
  p.in = (const __nv_bfloat16 *)in;
  p.gy = (const __nv_bfloat16 *)grad_out;
  p.pairs = (const int2 *)pairs;
  p.start = start;
  p.gw = grad_w;
  p.K = k;
  p.c_in = c_in;
  p.c_out = c_out;
  p.n_mma = (c_out + 15) / 16 * 16;
  p.nb = (c_out + 63) / 64;
  uint32_t cols = 32;
  while (cols < (uint32_t)p.n_mma) cols <<= 1;
  p.tmem_cols = cols;
  const char *dbg_env = getenv("TSG_WG_DEBUG");
  p.dbg = dbg_env ? atoi(dbg_env) : 0;
  const int ci_tiles = (c_in + 127) / 128;
  // about four units per SM (two CTAs of <= 97 KB fit an SM; the widest layers run one), never more than the chunks there can be
  long long target = 4LL * num_sms();
  const long long max_chunks = (pair_cap + WT_PAIRS - 1) / WT_PAIRS;
  if (target > max_chunks) target = max_chunks;
  if (target < 1) target = 1;
  p.target = (int)target;
  const long long units = target + k;
  const size_t smem = 1024 + (size_t)WT_STAGES * (2 + p.nb) * WT_BLOCK;
  static bool configured[64] = {false};
  int dev = 0;
  TSG_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    TSG_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + WT_STAGES * 6 * WT_BLOCK));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  dim3 grid((unsigned)units, (unsigned)ci_tiles);
  conv_wgrad_tc_kernel<<<grid, WT_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("tsg_conv_wgrad_tc");
}

/* profiling aid, not part of the public header */
int tsg_debug_wgrad_prof(long long *host) {
  TSG_CUDA(cudaDeviceSynchronize());
  TSG_CUDA(cudaMemcpyFromSymbol(host, g_wg_prof, sizeof(long long) * 8));
  return TSG_OK;
}

}  // extern "C"
