// Sparse convolution on the 5th-generation tensor cores (sm_100a): output-stationary implicit GEMM,
//   out[o,:] = epilogue( sum_k in[nbr[k,o],:] @ W[k] ),   bf16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM walks 128-row output tiles.  Warp roles (288 threads):
//   warps 0-3  epilogue   tcgen05.ld the 128 x c_out fp32 accumulator (one TMEM lane quadrant per warp),
//                         + bias + residual, ReLU, store bf16/fp32 rows
//   warp  4    MMA        one elected thread issues tcgen05.mma (M=128, N=c_out, K=16) per 16 input channels and
//                         commits to mbarriers; also owns the TMEM allocation (2 accumulator buffers)
//   warps 5-8  producers  gather the 128 neighbour rows of (tile, offset k) with 16-byte cp.async (zero-fill for
//                         missing neighbours) into a 128B-swizzled K-major tile; one thread starts the bulk copy
//                         (cp.async.bulk -> UBLKCP) of the pre-swizzled W[k] slice, completing on the same mbarrier
// A "unit" is one (offset k, 64-channel slice): 16 KB of A + c_out*128 B of B per pipeline stage.  Offsets for
// which no row of the tile has a neighbour are skipped by all roles (tile_mask).  The accumulator is double
// buffered in TMEM so the epilogue of tile t overlaps the mainloop of tile t+1.
#include "common.cuh"

namespace tsg {

constexpr int TC_BM = 128;
constexpr int TC_KB = 64;                       // channels per unit (128 B of bf16)
constexpr int TC_A_BYTES = TC_BM * TC_KB * 2;   // 16 KB
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_EPI_WARPS = 4, TC_PROD_WARPS = 4;
constexpr int TC_THREADS = 32 * (TC_EPI_WARPS + 1 + TC_PROD_WARPS);
constexpr int TC_LAG = 2;                       // cp.async groups kept in flight per producer thread

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    const uint64_t now = global_ns();
    if (!t0) t0 = now;
    else if (now - t0 > 4000000000ull) __trap();  // 4 s: a protocol bug must fail loudly, never hang the GPU
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 B, 8-row groups
// 1024 B apart (SBO), version 1, layout type 2.  `addr` may be advanced by 32 B per K=16 step inside the row.
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct TcParams {
  const __nv_bfloat16 *in0, *in1;
  int c0, c1, kb0, kb1;
  const uint8_t *packed_w;
  int K, c_out, stages;
  const int *nbr;
  const unsigned *tile_mask;
  long long n_out;
  void *out;
  int out_f32;
  const float *bias;
  const __nv_bfloat16 *residual;
  int relu;
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_MAX_STAGES + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.kb0 + p.kb1;
  const uint32_t b_bytes = (uint32_t)p.c_out * 128u;
  const uint32_t stage_bytes = TC_A_BYTES + b_bytes;
  const long long num_tiles = (p.n_out + TC_BM - 1) / TC_BM;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TC_MAX_STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * TC_MAX_STAGES]), tempty0 = smem_u32(&bars[2 * TC_MAX_STAGES + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, TC_PROD_WARPS * 32);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, TC_EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS) {  // TMEM allocation by the MMA warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < TC_EPI_WARPS) {
    // ================================================================= epilogue
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      const unsigned mask = p.tile_mask ? p.tile_mask[tile] : 0xffffffffu;
      mbar_wait(tfull0 + 8 * buf, ph);
      tc_fence_after();
      const long long row = tile * TC_BM + warp * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * (uint32_t)p.c_out;
      for (int c = 0; c < p.c_out; c += 16) {
        uint32_t v[16];
        if (mask) {
          tmem_ld16(taddr + c, v);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (row < p.n_out) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += __ldg(p.bias + c + j);
          }
          if (p.residual) {
            const uint4 *rp = reinterpret_cast<const uint4 *>(p.residual + row * p.c_out + c);
            const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
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
          if (p.out_f32) {
            float4 *op = reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out) + row * p.c_out + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              w[j] = *reinterpret_cast<const uint32_t *>(&h);
            }
            uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + row * p.c_out + c);
            op[0] = make_uint4(w[0], w[1], w[2], w[3]);
            op[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty0 + 8 * buf);
    }
  } else if (warp == TC_EPI_WARPS) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.c_out >> 3) << 17) | ((TC_BM >> 4) << 24);
      uint32_t it = 0, unit = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1, ph = (it >> 1) & 1;
        const unsigned mask = p.tile_mask ? p.tile_mask[tile] : 0xffffffffu;
        mbar_wait(tempty0 + 8 * buf, ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)p.c_out;
        uint32_t first = 1;
        for (int k = 0; k < p.K; ++k) {
          if (!((mask >> k) & 1u)) continue;
          for (int j = 0; j < KB; ++j, ++unit) {
            const uint32_t s = unit % p.stages, sph = (unit / p.stages) & 1;
            const int kc = j < p.kb0 ? min(TC_KB, p.c0 - j * TC_KB) : min(TC_KB, p.c1 - (j - p.kb0) * TC_KB);
            mbar_wait(full0 + 8 * s, sph);
            tc_fence_after();
            const uint32_t a_addr = smem_base + s * stage_bytes, b_addr = a_addr + TC_A_BYTES;
            for (int ks = 0; ks < kc; ks += 16) {
              umma_bf16(d_tmem, umma_desc(a_addr + ks * 2), umma_desc(b_addr + ks * 2), idesc, first ? 0u : 1u);
              first = 0;
            }
            umma_commit(empty0 + 8 * s);  // frees the stage once these MMAs have read it
          }
        }
        umma_commit(tfull0 + 8 * buf);  // accumulator complete (arrives immediately if the tile had no work)
      }
    }
    __syncwarp();
  } else {
    // ================================================================= producers
    const int pt = threadIdx.x - 32 * (TC_EPI_WARPS + 1);  // 0..127
    const int chunk = pt & 7, rsub = pt >> 3;              // 8 lanes cover one 128-byte row; 16 rows per pass
    uint32_t unit = 0, arrived = 0;                        // units issued / units signalled on their full barrier
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const unsigned mask = p.tile_mask ? p.tile_mask[tile] : 0xffffffffu;
      const long long m0 = tile * TC_BM;
      for (int k = 0; k < p.K; ++k) {
        if (!((mask >> k) & 1u)) continue;
        int src[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const long long o = m0 + rsub + 16 * q;
          src[q] = o < p.n_out ? (p.nbr ? __ldg(p.nbr + (long long)k * p.n_out + o) : (int)o) : -1;
        }
        for (int j = 0; j < KB; ++j, ++unit) {
          const uint32_t s = unit % p.stages, sph = (unit / p.stages) & 1;
          mbar_wait(empty0 + 8 * s, sph ^ 1);
          const uint32_t a_addr = smem_base + s * stage_bytes;
          const bool second = j >= p.kb0;
          const __nv_bfloat16 *base = second ? p.in1 : p.in0;
          const int cs = second ? p.c1 : p.c0;
          const int ch0 = (second ? j - p.kb0 : j) * TC_KB;
          if (pt == 0) {
            mbar_expect_tx(full0 + 8 * s, b_bytes);
            bulk_g2s(a_addr + TC_A_BYTES, p.packed_w + ((size_t)k * KB + j) * b_bytes, b_bytes, full0 + 8 * s);
          }
          if (ch0 + chunk * 8 < cs) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int r = rsub + 16 * q;
              const uint32_t dst = a_addr + r * 128 + ((chunk ^ (r & 7)) << 4);
              const bool ok = src[q] >= 0;
              const __nv_bfloat16 *g = ok ? base + (size_t)src[q] * cs + ch0 + chunk * 8 : base;
              cp_async16(dst, g, ok ? 16u : 0u);
            }
          }
          cp_async_commit();
          if (unit >= TC_LAG) {  // the group issued TC_LAG units ago has landed: publish it
            cp_async_wait<TC_LAG>();
            fence_async_proxy();
            mbar_arrive(full0 + 8 * (arrived % p.stages));
            ++arrived;
          }
        }
      }
    }
    cp_async_wait<0>();
    fence_async_proxy();
    for (; arrived < unit; ++arrived) mbar_arrive(full0 + 8 * (arrived % p.stages));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS) {
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

int tsg_conv_fwd_tc(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                    int c_out, const int32_t *nbr, const uint32_t *tile_mask, int64_t n_out, void *out, int out_dtype,
                    const float *bias, const void *residual, int relu, int num_sms_hint, tsg_stream_t stream) {
  (void)n_in;
  if (c0 % 16 || c1 % 16 || c_out % 16 || c_out > 256 || c_out <= 0 || c0 <= 0 || k <= 0 || k > 32 ||
      (out_dtype != TSG_BF16 && out_dtype != TSG_F32)) {
    set_error("tsg_conv_fwd_tc: need c0,c1,c_out multiples of 16, c_out<=256, K<=32, out bf16/f32");
    return TSG_ERR_UNSUPPORTED;
  }
  if (n_out <= 0) return TSG_OK;
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
  p.n_out = n_out;
  p.out = out;
  p.out_f32 = out_dtype == TSG_F32;
  p.bias = bias;
  p.residual = (const __nv_bfloat16 *)residual;
  p.relu = relu;
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)c_out) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage_bytes = TC_A_BYTES + (size_t)c_out * 128;
  const size_t budget = 200 * 1024;
  int stages = (int)(budget / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < TC_LAG + 1) {
    set_error("tsg_conv_fwd_tc: not enough shared memory for the pipeline");
    return TSG_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const size_t smem = stages * stage_bytes + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    TSG_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    configured = 220 * 1024;
  }
  const long long num_tiles = (n_out + TC_BM - 1) / TC_BM;
  int sms = num_sms_hint > 0 ? num_sms_hint : num_sms();
  const unsigned grid = (unsigned)(num_tiles < sms ? num_tiles : sms);
  conv_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(p);
  return check_launch("tsg_conv_fwd_tc");
}

}  // extern "C"
