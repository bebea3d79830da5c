// Sparse convolution on the 5th-generation tensor cores (sm_100a): output-stationary implicit GEMM,
//   out[row(r),:] = epilogue( sum_k in[nbr[k,r],:] @ W[k] ),   bf16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM walks "super tiles" of G x 128 tile rows (G in {1,2,4}).  The unit of the shared-memory
// pipeline is a STAGE = one (kernel offset k, 64-channel slice j): the pre-swizzled weight slice W[k][j] (c_out x 128 B)
// plus the G gathered 128 x 64 A tiles that multiply it, behind ONE full / ONE empty mbarrier — so the weight slice is
// loaded once per super tile and the per-stage fixed costs (barrier waits, proxy fence, commit) are paid once per
// 4G MMAs.  Warp roles (448 threads):
//   warps 0-3   epilogue   tcgen05.ld the G 128 x c_out fp32 accumulators (one TMEM lane quadrant per warp),
//                          + bias + residual, ReLU, store bf16/fp32 rows (to row perm[r] when tiles are mask-sorted)
//   warp  4     MMA        converged warp; one elected lane issues tcgen05.mma (M=128, N=c_out, K=16) per 16 input
//                          channels and commits the stage; owns the TMEM allocation (2 buffers x G accumulators)
//   warp  5     weights    one thread streams W[k][j] into the stage with cp.async.bulk (UBLKCP, complete_tx)
//   warps 6-13  producers  gather the neighbour rows with 16-byte cp.async (LDGSTS, zero-fill for missing neighbours)
//                          into the 128B-swizzled K-major A tiles and arrive on the stage's full barrier with
//                          cp.async.mbarrier.arrive.noinc (no wait in the producer: the ring depth is the only limit on
//                          loads in flight).  Neighbour indices of offset k+1 are prefetched while k is being issued.
// History (profiles/ncu_conv_r01.md): v1 paid ~200 producer instructions per 16 KB slot (runtime modulo, per-row
// address arithmetic) and was bound by producer issue; v3 made the producers lean and became bound by the single MMA
// thread's own instruction latency (two try_waits, MEMBAR + proxy fence, descriptor rebuilds and ELECT loops per slot).
// Offsets for which a sub-tile has no neighbour at all are skipped by every role (tile_mask); with rows sorted by
// their neighbour bit mask (tsg_kmap_sort_rows) that removes more than half of the (tile, offset) work.
// Accumulators are double buffered in TMEM so the epilogue of super tile t overlaps the mainloop of t+1.
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace tsg {

constexpr int V4_MMA_WARP = TC_EPI_WARPS;        // 4
constexpr int V4_W_WARP = TC_EPI_WARPS + 1;      // 5
constexpr int V4_PROD_WARP0 = TC_EPI_WARPS + 2;  // 6
constexpr int V4_PROD_WARPS = 8;
constexpr int V4_PROD_THREADS = 32 * V4_PROD_WARPS;
constexpr int V4_THREADS = 32 * (V4_PROD_WARP0 + V4_PROD_WARPS);  // 448
constexpr int V4_MAX_STAGES = 8;

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

template <int G>
__global__ void __launch_bounds__(V4_THREADS, 1) conv_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * V4_MAX_STAGES + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.kb0 + p.kb1;
  const uint32_t b_bytes = (uint32_t)p.c_out * 128u;                 // multiple of 2048
  const uint32_t stage_bytes = b_bytes + (uint32_t)G * TC_A_BYTES;   // [W slice][A tile 0]..[A tile G-1]
  const uint32_t nst = (uint32_t)p.na;                               // stages
  const long long num_tiles = (p.n_out + TC_BM - 1) / TC_BM;
  const long long num_super = (num_tiles + G - 1) / G;
  const unsigned kmask = p.K >= 32 ? 0xffffffffu : ((1u << p.K) - 1u);
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[V4_MAX_STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * V4_MAX_STAGES]), tempty0 = tfull0 + 16;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < nst; ++s) {
      mbar_init(full0 + 8 * s, V4_PROD_THREADS + 1);  // every producer thread (noinc arrive) + the weight thread
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, TC_EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == V4_MMA_WARP) {  // TMEM allocation by the MMA warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  auto tile_masks = [&](long long st, unsigned (&masks)[G]) -> unsigned {
    unsigned um = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long tile = st * G + g;
      masks[g] = tile < num_tiles ? ((p.tile_mask ? __ldg(p.tile_mask + tile) : 0xffffffffu) & kmask) : 0u;
      um |= masks[g];
    }
    return um;
  };

  if (warp < TC_EPI_WARPS) {
    // ================================================================= epilogue
    uint32_t it = 0;
    for (long long st = blockIdx.x; st < num_super; st += gridDim.x, ++it) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      unsigned masks[G];
      long long rows[G];
      tile_masks(st, masks);
#pragma unroll
      for (int g = 0; g < G; ++g) {  // destination rows are fetched before the (long) wait for the accumulators
        const long long r = (st * G + g) * TC_BM + warp * 32 + lane;
        rows[g] = r < p.n_out ? (p.perm ? (long long)__ldg(p.perm + r) : r) : -1;
      }
      mbar_wait_sleep(tfull0 + 8 * buf, ph);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (st * G + g >= num_tiles) break;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (buf * G + g) * (uint32_t)p.c_out;
        for (int c = 0; c < p.c_out; c += 16) {
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
    }
  } else if (warp == V4_MMA_WARP) {
    // ================================================================= MMA issuer (whole warp converged)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.c_out >> 3) << 17) | ((TC_BM >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
    const uint32_t desc_lo_stage = stage_bytes >> 4, desc_lo_a = b_bytes >> 4;
    const uint32_t desc_lo0 = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);       // LBO field = 1 (ignored for swizzled K-major)
    Ring r;
    uint32_t it = 0;
    for (long long st = blockIdx.x; st < num_super; st += gridDim.x, ++it) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      unsigned masks[G];
      const unsigned umask = tile_masks(st, masks);
      mbar_wait(tempty0 + 8 * buf, ph ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + buf * G * (uint32_t)p.c_out;
      unsigned started = 0;
      for (int k = next_bit(umask, -1); k < 32; k = next_bit(umask, k)) {
        unsigned act = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) act |= ((masks[g] >> k) & 1u) << g;
        for (int j = 0; j < KB; ++j) {
          const int kc = j < p.kb0 ? min(TC_KB, p.c0 - j * TC_KB) : min(TC_KB, p.c1 - (j - p.kb0) * TC_KB);
          const int nk = kc >> 4;
          mbar_wait(full0 + 8 * r.slot, r.phase);
          fence_async_proxy();  // cp.async wrote through the generic proxy; the MMA reads through the async proxy
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_lo = desc_lo0 + r.slot * desc_lo_stage;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              if (!((act >> g) & 1u)) continue;
              const uint32_t a_lo = b_lo + desc_lo_a + g * (TC_A_BYTES >> 4);
              const uint32_t d_tmem = d0 + g * (uint32_t)p.c_out;
              umma_bf16(d_tmem, make_desc(a_lo, desc_hi), make_desc(b_lo, desc_hi), idesc, (started >> g) & 1u);
              if (nk > 1) umma_bf16(d_tmem, make_desc(a_lo + 2, desc_hi), make_desc(b_lo + 2, desc_hi), idesc, 1u);
              if (nk > 2) umma_bf16(d_tmem, make_desc(a_lo + 4, desc_hi), make_desc(b_lo + 4, desc_hi), idesc, 1u);
              if (nk > 3) umma_bf16(d_tmem, make_desc(a_lo + 6, desc_hi), make_desc(b_lo + 6, desc_hi), idesc, 1u);
            }
            umma_commit(empty0 + 8 * r.slot);  // frees the stage once these MMAs have read it
          }
          __syncwarp();
          started |= act;
          r.advance(nst);
        }
      }
      if (elect_one()) umma_commit(tfull0 + 8 * buf);  // accumulators complete (immediately if there was no work)
      __syncwarp();
    }
  } else if (warp == V4_W_WARP) {
    // ================================================================= weight loader
    if (lane == 0) {
      Ring r;
      for (long long st = blockIdx.x; st < num_super; st += gridDim.x) {
        unsigned masks[G];
        const unsigned umask = tile_masks(st, masks);
        for (int k = next_bit(umask, -1); k < 32; k = next_bit(umask, k)) {
          const uint8_t *wk = p.packed_w + (size_t)k * KB * b_bytes;
          for (int j = 0; j < KB; ++j) {
            mbar_wait(empty0 + 8 * r.slot, r.phase ^ 1);
            mbar_arrive_expect_tx(full0 + 8 * r.slot, b_bytes);
            bulk_g2s(smem_base + r.slot * stage_bytes, wk + (size_t)j * b_bytes, b_bytes, full0 + 8 * r.slot);
            r.advance(nst);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================= producers
    const int pt = threadIdx.x - 32 * V4_PROD_WARP0;  // 0..255
    const int chunk = pt & 7, rsub = pt >> 3;         // 8 lanes cover one 128-byte row; 32 rows per pass, 4 passes
    const uint32_t dst_off = b_bytes + (uint32_t)rsub * 128u + (uint32_t)((chunk ^ (rsub & 7)) << 4);
    const uint32_t col_bytes = (uint32_t)chunk * 16u;
    const uint32_t rb0 = (uint32_t)p.c0 * 2u, rb1 = (uint32_t)p.c1 * 2u;
    const char *in0 = reinterpret_cast<const char *>(p.in0) + col_bytes;
    const char *in1 = reinterpret_cast<const char *>(p.in1) + col_bytes;
    Ring r;
    for (long long st = blockIdx.x; st < num_super; st += gridDim.x) {
      unsigned masks[G];
      const unsigned umask = tile_masks(st, masks);
      const long long m0 = st * G * TC_BM + rsub;
      int cur[G][4], nxt[G][4];
      auto load_idx = [&](int (&dst)[G][4], int k) {
        const int *src = p.nbr + (long long)k * p.n_out;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const bool act = (masks[g] >> k) & 1u;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const long long o = m0 + g * TC_BM + 32 * q;
            dst[g][q] = (act && o < p.n_out) ? (p.nbr ? __ldg(src + o) : (int)o) : -1;
          }
        }
      };
      int k = next_bit(umask, -1);
      if (k < 32) load_idx(cur, k);
      while (k < 32) {
        const int kn = next_bit(umask, k);
        if (kn < 32) load_idx(nxt, kn);  // prefetch the next offset's neighbour rows while this one is issued
        for (int j = 0; j < KB; ++j) {
          const bool second = j >= p.kb0;
          const uint32_t rb = second ? rb1 : rb0;
          const uint32_t ch0b = (uint32_t)(second ? j - p.kb0 : j) * 128u;
          const char *bp = (second ? in1 : in0) + ch0b;
          const bool chunk_ok = ch0b + col_bytes < rb;   // this 16-byte chunk exists in the (possibly partial) slice
          mbar_wait(empty0 + 8 * r.slot, r.phase ^ 1);
          const uint32_t dst = smem_base + r.slot * stage_bytes + dst_off;
          if (chunk_ok) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
              if (!((masks[g] >> k) & 1u)) continue;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int idx = cur[g][q];
                const bool ok = idx >= 0;
                cp_async16(dst + g * TC_A_BYTES + q * 4096u, bp + (size_t)(ok ? idx : 0) * rb, ok ? 16u : 0u);
              }
            }
          }
          cp_async_arrive(full0 + 8 * r.slot);
          r.advance(nst);
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int q = 0; q < 4; ++q) cur[g][q] = nxt[g][q];
        k = kn;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == V4_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// W (K, c_in, c_out) fp32 -> per (k, 64-channel slice) [c_out][64] bf16, K-major, 128B-swizzled: the byte image the
// MMA reads, so a linear bulk copy stages it.  Optional per-output-channel scale (folded BatchNorm).
__global__ void pack_weights_kernel(const float *__restrict__ w, int K, int c_in, int c_out, int c0, int c1, int kb0,
                                    int KB, const float *__restrict__ out_scale, __nv_bfloat16 *__restrict__ packed) {
  const long long total = (long long)K * KB * c_out * TC_KB;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % TC_KB);
    const int n = (int)((t / TC_KB) % c_out);
    const int j = (int)((t / ((long long)TC_KB * c_out)) % KB);
    const int k = (int)(t / ((long long)TC_KB * c_out * KB));
    const bool second = j >= kb0;
    const int ch = (second ? j - kb0 : j) * TC_KB + c;          // channel within its source tensor
    const int g = second ? c0 + ch : ch;                        // row of W[k]
    float v = 0.f;
    if (ch < (second ? c1 : c0) && g < c_in) {
      v = w[((long long)k * c_in + g) * c_out + n];
      if (out_scale) v *= out_scale[n];
    }
    const long long blk = ((long long)k * KB + j) * c_out * TC_KB;
    const int sw = (((c >> 3) ^ (n & 7)) << 3) | (c & 7);
    packed[blk + (long long)n * TC_KB + sw] = __float2bfloat16_rn(v);
  }
}

}  // namespace tsg

using namespace tsg;

extern "C" {

size_t tsg_conv_pack_bytes(int k, int c0, int c1, int c_out) {
  const int KB = (c0 + TC_KB - 1) / TC_KB + (c1 + TC_KB - 1) / TC_KB;
  return (size_t)k * KB * c_out * TC_KB * 2;
}

int tsg_conv_pack_weights(const float *weight, int k, int c_in, int c_out, int c0, int c1, const float *out_scale,
                          void *packed, tsg_stream_t stream) {
  if (c0 % 16 || c1 % 16 || c_out % 16 || c_out > 256 || c_out <= 0 || c0 <= 0 || c0 + c1 < c_in) {
    set_error("tsg_conv_pack_weights: need c0,c1,c_out multiples of 16, c_out<=256, c0+c1>=c_in");
    return TSG_ERR_UNSUPPORTED;
  }
  const int kb0 = (c0 + TC_KB - 1) / TC_KB, KB = kb0 + (c1 + TC_KB - 1) / TC_KB;
  const long long total = (long long)k * KB * c_out * TC_KB;
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, stream>>>(weight, k, c_in, c_out, c0, c1, kb0, KB, out_scale,
                                                                (__nv_bfloat16 *)packed);
  return check_launch("tsg_conv_pack_weights");
}

}  // extern "C"

namespace tsg {
// TMA gather4 producer variant (conv_tc_tma.cu): opt-in with TSG_TC_IMPL=tma, kept for A/B timing
int conv_fwd_tc_tma(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                    int c_out, const int32_t *nbr, const uint32_t *tile_mask, int64_t n_out, void *out, int out_dtype,
                    const float *bias, const void *residual, int relu, int num_sms_hint, tsg_stream_t stream);
}  // namespace tsg

extern "C" {

int tsg_conv_fwd_tc(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                    int c_out, const int32_t *nbr, const uint32_t *tile_mask, const int32_t *perm, int64_t n_out,
                    void *out, int out_dtype, const float *bias, const void *residual, int relu, int num_sms_hint,
                    tsg_stream_t stream) {
  if (c0 % 16 || c1 % 16 || c_out % 16 || c_out > 256 || c_out <= 0 || c0 <= 0 || k <= 0 || k > 32 ||
      (out_dtype != TSG_BF16 && out_dtype != TSG_F32)) {
    set_error("tsg_conv_fwd_tc: need c0,c1,c_out multiples of 16, c_out<=256, K<=32, out bf16/f32");
    return TSG_ERR_UNSUPPORTED;
  }
  if (n_out <= 0) return TSG_OK;
  static const char *impl = getenv("TSG_TC_IMPL");
  if (impl && strcmp(impl, "tma") == 0 && !perm) {
    const int rc = conv_fwd_tc_tma(in0, c0, in1, c1, n_in, packed_w, k, c_out, nbr, tile_mask, n_out, out, out_dtype,
                                   bias, residual, relu, num_sms_hint, stream);
    if (rc >= 0) return rc;  // < 0: no tensor map could be encoded, use the default producer
  }
  TcParams p;
  p.in0 = (const __nv_bfloat16 *)in0;
  p.in1 = (const __nv_bfloat16 *)in1;
  p.c0 = c0;
  p.c1 = c1;
  p.kb0 = (c0 + TC_KB - 1) / TC_KB;
  p.kb1 = (c1 + TC_KB - 1) / TC_KB;
  p.packed_w = (const uint8_t *)packed_w;
  p.K = k;
  p.c_out = c_out;
  p.nbr = nbr;
  p.tile_mask = tile_mask;
  p.perm = perm;
  p.n_out = n_out;
  p.out = out;
  p.out_f32 = out_dtype == TSG_F32;
  p.bias = bias;
  p.residual = (const __nv_bfloat16 *)residual;
  p.relu = relu;
  const int sms = num_sms_hint > 0 ? num_sms_hint : num_sms();
  const long long num_tiles = (n_out + TC_BM - 1) / TC_BM;
  // G sub-tiles share every weight slice; bounded by TMEM (2 buffers x G x c_out fp32 columns <= 512) and by the
  // number of super tiles needed to keep every SM busy
  int G = c_out <= 64 ? 4 : (c_out <= 128 ? 2 : 1);
  while (G > 1 && (num_tiles + G - 1) / G < 2LL * sms) G >>= 1;
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)G * (uint32_t)c_out) cols <<= 1;
  p.tmem_cols = cols;
  const size_t b_bytes = (size_t)c_out * 128, stage_bytes = b_bytes + (size_t)G * TC_A_BYTES, budget = 222 * 1024;
  int stages = (int)(budget / stage_bytes);
  if (stages > V4_MAX_STAGES) stages = V4_MAX_STAGES;
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
    TSG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured = true;
  }
  const long long num_super = (num_tiles + G - 1) / G;
  const unsigned grid = (unsigned)(num_super < sms ? num_super : sms);
  if (G == 4) conv_tc_kernel<4><<<grid, V4_THREADS, smem, stream>>>(p);
  else if (G == 2) conv_tc_kernel<2><<<grid, V4_THREADS, smem, stream>>>(p);
  else conv_tc_kernel<1><<<grid, V4_THREADS, smem, stream>>>(p);
  return check_launch("tsg_conv_fwd_tc");
}

}  // extern "C"
