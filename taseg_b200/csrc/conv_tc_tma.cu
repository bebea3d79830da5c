// Sparse convolution on tcgen05 with a TMA-gather producer (sm_100a) — opt-in variant (TSG_TC_IMPL=tma).
// Measured on B200 (profiles/conv_layers_r01_*.txt): 2x SLOWER than the cp.async producer of conv_tc.cu — the TMA
// unit serialises the 128-byte rows of a gather4 (~18 cycles per row), so the default path gathers with LDGSTS.
//
// Same output-stationary implicit GEMM as conv_tc.cu (G sub-tiles of 128 rows share every weight slice, fp32
// accumulators double buffered in TMEM, fused bias/residual/ReLU epilogue), but the A operand is staged by the TMA
// unit instead of by threads: one warp issues `cp.async.bulk.tensor.2d...tile::gather4` (UTMALDG), each instruction
// fetching FOUR neighbour rows of 64 channels straight into the 128B-swizzled K-major slot the MMA reads.  A missing
// neighbour is row index -1: the TMA unit zero-fills out-of-bounds rows, so there is no predication, no zero-fill
// store and no proxy fence, and the whole 128-row tile costs 32 instructions instead of ~300 per thread.
// Warp roles (224 threads): 0-3 epilogue, 4 MMA issuer + TMEM owner, 5 gather (indices prefetched 4 offsets ahead
// through a small cp.async ring in shared memory), 6 weight loader (cp.async.bulk of the pre-swizzled W[k] slices).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace tsg {

constexpr int T3_THREADS = 7 * 32;
constexpr int T3_NI = 8;  // index ring slots (offsets)
constexpr int T3_D = 4;   // offsets of index prefetch in flight

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int4 lds128(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

template <int G>
__global__ void __launch_bounds__(T3_THREADS, 1) conv_tma_kernel(const __grid_constant__ CUtensorMap tmap0,
                                                                 const __grid_constant__ CUtensorMap tmap1,
                                                                 const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_MAX_A + 2 * TC_MAX_B + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.kb0 + p.kb1;
  const uint32_t b_bytes = (uint32_t)p.c_out * 128u;
  const uint32_t a_base = smem_base + (uint32_t)p.nb * b_bytes;
  const uint32_t idx_base = a_base + (uint32_t)p.na * TC_A_BYTES;
  const long long num_tiles = (p.n_out + TC_BM - 1) / TC_BM;
  const long long num_super = (num_tiles + G - 1) / G;
  const unsigned kmask = p.K >= 32 ? 0xffffffffu : ((1u << p.K) - 1u);
  const uint32_t afull0 = smem_u32(&bars[0]), aempty0 = smem_u32(&bars[TC_MAX_A]);
  const uint32_t bfull0 = smem_u32(&bars[2 * TC_MAX_A]), bempty0 = smem_u32(&bars[2 * TC_MAX_A + TC_MAX_B]);
  const uint32_t tfull0 = smem_u32(&bars[2 * TC_MAX_A + 2 * TC_MAX_B]), tempty0 = tfull0 + 16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.na; ++s) {
      mbar_init(afull0 + 8 * s, 1);
      mbar_init(aempty0 + 8 * s, 1);
    }
    for (int s = 0; s < p.nb; ++s) {
      mbar_init(bfull0 + 8 * s, 1);
      mbar_init(bempty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, TC_EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS) {
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
      masks[g] = tile < num_tiles ? ((p.tile_mask ? p.tile_mask[tile] : 0xffffffffu) & kmask) : 0u;
      um |= masks[g];
    }
    return um;
  };

  if (warp < TC_EPI_WARPS) {
    // ================================================================= epilogue
    uint32_t it = 0;
    for (long long st = blockIdx.x; st < num_super; st += gridDim.x, ++it) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      mbar_wait_sleep(tfull0 + 8 * buf, ph);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const long long tile = st * G + g;
        if (tile >= num_tiles) break;
        const unsigned mask = (p.tile_mask ? p.tile_mask[tile] : 0xffffffffu) & kmask;
        const long long row = tile * TC_BM + warp * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (buf * G + g) * (uint32_t)p.c_out;
        for (int c = 0; c < p.c_out; c += 16) {
          uint32_t v[16];
          if (mask) {
            tmem_ld16(taddr + c, v);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
          }
          if (row < p.n_out) epilogue_store16(p, row, c, v);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty0 + 8 * buf);
    }
  } else if (warp == TC_EPI_WARPS) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.c_out >> 3) << 17) | ((TC_BM >> 4) << 24);
      Ring a, b;
      uint32_t it = 0;
      for (long long st = blockIdx.x; st < num_super; st += gridDim.x, ++it) {
        const uint32_t buf = it & 1, ph = (it >> 1) & 1;
        unsigned masks[G];
        const unsigned umask = tile_masks(st, masks);
        mbar_wait(tempty0 + 8 * buf, ph ^ 1);
        tc_fence_after();
        unsigned started = 0;
        for (int k = next_bit(umask, -1); k < 32; k = next_bit(umask, k)) {
          for (int j = 0; j < KB; ++j) {
            const int kc = j < p.kb0 ? min(TC_KB, p.c0 - j * TC_KB) : min(TC_KB, p.c1 - (j - p.kb0) * TC_KB);
            mbar_wait(bfull0 + 8 * b.slot, b.phase);
            const uint32_t b_addr = smem_base + b.slot * b_bytes;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              if (!((masks[g] >> k) & 1u)) continue;
              mbar_wait(afull0 + 8 * a.slot, a.phase);
              tc_fence_after();
              const uint32_t a_addr = a_base + a.slot * TC_A_BYTES;
              const uint32_t d_tmem = tmem_base + (buf * G + g) * (uint32_t)p.c_out;
              for (int ks = 0; ks < kc; ks += 16) {
                umma_bf16(d_tmem, umma_desc(a_addr + ks * 2), umma_desc(b_addr + ks * 2), idesc, (started >> g) & 1u);
                started |= 1u << g;
              }
              umma_commit(aempty0 + 8 * a.slot);
              a.advance(p.na);
            }
            umma_commit(bempty0 + 8 * b.slot);
            b.advance(p.nb);
          }
        }
        umma_commit(tfull0 + 8 * buf);
      }
    }
    __syncwarp();
  } else if (warp == TC_EPI_WARPS + 1) {
    // ================================================================= gather producer (TMA gather4)
    Ring a;
    for (long long st = blockIdx.x; st < num_super; st += gridDim.x) {
      unsigned masks[G];
      const unsigned umask = tile_masks(st, masks);
      const long long m0 = st * G * TC_BM;
      int kp = next_bit(umask, -1);   // next offset whose indices will be prefetched
      uint32_t sp = 0, sc = 0;        // index-ring slots: prefetch / consume
      auto prefetch = [&]() {
        if (p.nbr && kp < 32) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (!((masks[g] >> kp) & 1u)) continue;
            const long long o = m0 + g * TC_BM + 4 * lane;
            const int *src = p.nbr + (long long)kp * p.n_out + o;
            const uint32_t dst = idx_base + ((sp * G + g) * TC_BM + 4 * lane) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) cp_async4(dst + 4 * q, o + q < p.n_out ? src + q : p.nbr, o + q < p.n_out ? 4u : 0u);
          }
          kp = next_bit(umask, kp);
        }
        cp_async_commit();
        if (++sp == T3_NI) sp = 0;
      };
#pragma unroll
      for (int d = 0; d < T3_D; ++d) prefetch();
      for (int k = next_bit(umask, -1); k < 32; k = next_bit(umask, k)) {
        int4 id[G];
        cp_async_wait<T3_D - 1>();
        __syncwarp();
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int o = (int)(m0 + g * TC_BM + 4 * lane);
          if (p.nbr) id[g] = lds128(idx_base + ((sc * G + g) * TC_BM + 4 * lane) * 4);
          else id[g] = make_int4(o, o + 1, o + 2, o + 3);   // identity map; rows >= n_in are zero-filled by the TMA unit
          if (p.nbr && o + 3 >= p.n_out) {                  // rows past the end were zero-filled as index 0: mark missing
            if (o + 0 >= p.n_out) id[g].x = -1;
            if (o + 1 >= p.n_out) id[g].y = -1;
            if (o + 2 >= p.n_out) id[g].z = -1;
            if (o + 3 >= p.n_out) id[g].w = -1;
          }
        }
        if (++sc == T3_NI) sc = 0;
        __syncwarp();
        prefetch();
        for (int j = 0; j < KB; ++j) {
          const bool second = j >= p.kb0;
          const CUtensorMap *map = second ? &tmap1 : &tmap0;
          const int ch0 = (second ? j - p.kb0 : j) * TC_KB;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (!((masks[g] >> k) & 1u)) continue;
            mbar_wait(aempty0 + 8 * a.slot, a.phase ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(afull0 + 8 * a.slot, TC_A_BYTES);
            __syncwarp();
            tma_gather4(a_base + a.slot * TC_A_BYTES + lane * 512, map, ch0, id[g].x, id[g].y, id[g].z, id[g].w,
                        afull0 + 8 * a.slot);
            a.advance(p.na);
          }
        }
      }
      cp_async_wait<0>();
    }
  } else {
    // ================================================================= weight loader
    if (lane == 0) {
      Ring b;
      for (long long st = blockIdx.x; st < num_super; st += gridDim.x) {
        unsigned masks[G];
        const unsigned umask = tile_masks(st, masks);
        for (int k = next_bit(umask, -1); k < 32; k = next_bit(umask, k)) {
          for (int j = 0; j < KB; ++j) {
            mbar_wait(bempty0 + 8 * b.slot, b.phase ^ 1);
            mbar_arrive_expect_tx(bfull0 + 8 * b.slot, b_bytes);
            bulk_g2s(smem_base + b.slot * b_bytes, p.packed_w + ((size_t)k * KB + j) * b_bytes, b_bytes,
                     bfull0 + 8 * b.slot);
            b.advance(p.nb);
          }
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

// (n rows, c channels) bf16 row-major feature matrix; box = 64 channels x 1 row (gather4 moves four of them),
// 128B swizzle to match the UMMA K-major descriptor; out-of-bounds rows/channels read as zero.
static bool make_feature_map(CUtensorMap *m, const void *base, int c, int64_t n) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)n};
  cuuint64_t strides[1] = {(cuuint64_t)c * 2};
  cuuint32_t box[2] = {TC_KB, 1};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Returns a tsg_status, or -1 when no tensor map could be encoded (the caller then uses the cp.async producer).
int conv_fwd_tc_tma(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                    int c_out, const int32_t *nbr, const uint32_t *tile_mask, int64_t n_out, void *out, int out_dtype,
                    const float *bias, const void *residual, int relu, int num_sms_hint, tsg_stream_t stream) {
  CUtensorMap tm0, tm1;
  bool tma = n_in > 0 && make_feature_map(&tm0, in0, c0, n_in);
  if (tma && c1 > 0) tma = make_feature_map(&tm1, in1, c1, n_in);
  if (!tma) return -1;
  if (c1 == 0) tm1 = tm0;

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
  p.perm = nullptr;
  p.n_out = n_out;
  p.out = out;
  p.out_f32 = out_dtype == TSG_F32;
  p.bias = bias;
  p.residual = (const __nv_bfloat16 *)residual;
  p.relu = relu;
  const int sms = num_sms_hint > 0 ? num_sms_hint : num_sms();
  const long long num_tiles = (n_out + TC_BM - 1) / TC_BM;
  int G = c_out <= 64 ? 4 : (c_out <= 128 ? 2 : 1);
  while (G > 1 && (num_tiles + G - 1) / G < 2LL * sms) G >>= 1;
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)G * (uint32_t)c_out) cols <<= 1;
  p.tmem_cols = cols;
  const size_t b_bytes = (size_t)c_out * 128, idx_bytes = (size_t)T3_NI * G * TC_BM * 4, budget = 200 * 1024;
  int nb = (int)((72 * 1024) / b_bytes);
  if (nb > TC_MAX_B) nb = TC_MAX_B;
  if (nb < 2) nb = 2;
  if (c_out > 128) nb = 3;
  int na = (int)((budget - nb * b_bytes - idx_bytes) / TC_A_BYTES);
  if (na > TC_MAX_A) na = TC_MAX_A;
  if (na < 2) {
    set_error("tsg_conv_fwd_tc: not enough shared memory for the pipeline");
    return TSG_ERR_UNSUPPORTED;
  }
  p.na = na;
  p.nb = nb;
  const size_t smem = nb * b_bytes + (size_t)na * TC_A_BYTES + idx_bytes + 1024;
  static bool configured = false;
  if (!configured) {
    TSG_CUDA(cudaFuncSetAttribute(conv_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    TSG_CUDA(cudaFuncSetAttribute(conv_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    TSG_CUDA(cudaFuncSetAttribute(conv_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  const long long num_super = (num_tiles + G - 1) / G;
  const unsigned grid = (unsigned)(num_super < sms ? num_super : sms);
  if (G == 4) conv_tma_kernel<4><<<grid, T3_THREADS, smem, stream>>>(tm0, tm1, p);
  else if (G == 2) conv_tma_kernel<2><<<grid, T3_THREADS, smem, stream>>>(tm0, tm1, p);
  else conv_tma_kernel<1><<<grid, T3_THREADS, smem, stream>>>(tm0, tm1, p);
  return check_launch("tsg_conv_fwd_tc(tma)");
}

}  // namespace tsg
