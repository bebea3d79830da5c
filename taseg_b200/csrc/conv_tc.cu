// Sparse convolution on the 5th-generation tensor cores (sm_100a): output-stationary implicit GEMM,
//   out[o,:] = epilogue( sum_k in[nbr[k,o],:] @ W[k] ),   bf16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM walks "super tiles" of G x 128 output rows (G in {1,2,4}).  Warp roles (288 threads):
//   warps 0-3  epilogue   tcgen05.ld the G 128 x c_out fp32 accumulators (one TMEM lane quadrant per warp),
//                         + bias + residual, ReLU, store bf16/fp32 rows
//   warp  4    MMA        one elected thread issues tcgen05.mma (M=128, N=c_out, K=16) per 16 input channels and
//                         commits to mbarriers; also owns the TMEM allocation (2 buffers x G accumulators)
//   warps 5-8  producers  gather the 128 neighbour rows of (sub-tile g, offset k) with 16-byte cp.async (zero-fill
//                         for missing neighbours) into a 128B-swizzled K-major A slot and hand it to the MMA warp
//                         with cp.async.mbarrier.arrive (no wait in the producer: the ring depth is the only limit
//                         on loads in flight); thread 0 also streams the pre-swizzled W[k] slices through a
//                         separate B ring with cp.async.bulk (UBLKCP).  Neighbour indices of offset k+1 are
//                         prefetched while offset k is being issued.
// W[k] (c_out*128 B per 64-channel slice) is loaded ONCE per super tile and reused by its G sub-tiles, which divides
// the dominant L2->SMEM stream of the narrow layers by G.  Offsets for which a sub-tile has no neighbour at all are
// skipped by every role (tile_mask).  Accumulators are double buffered in TMEM so the epilogue of super tile t
// overlaps the mainloop of t+1.
#include "tc_common.cuh"

namespace tsg {

template <int G>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_MAX_A + 2 * TC_MAX_B + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.kb0 + p.kb1;
  const uint32_t b_bytes = (uint32_t)p.c_out * 128u;
  const uint32_t a_base = smem_base + (uint32_t)p.nb * b_bytes;  // B ring first (b_bytes is a multiple of 2048)
  const long long num_tiles = (p.n_out + TC_BM - 1) / TC_BM;
  const long long num_super = (num_tiles + G - 1) / G;
  const unsigned kmask = p.K >= 32 ? 0xffffffffu : ((1u << p.K) - 1u);
  const uint32_t afull0 = smem_u32(&bars[0]), aempty0 = smem_u32(&bars[TC_MAX_A]);
  const uint32_t bfull0 = smem_u32(&bars[2 * TC_MAX_A]), bempty0 = smem_u32(&bars[2 * TC_MAX_A + TC_MAX_B]);
  const uint32_t tfull0 = smem_u32(&bars[2 * TC_MAX_A + 2 * TC_MAX_B]), tempty0 = tfull0 + 16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.na; ++s) {
      mbar_init(afull0 + 8 * s, TC_PROD_WARPS * 32);
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
    for (long long st = blockIdx.x; st < num_super; st += gridDim.x, ++it) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      mbar_wait(tfull0 + 8 * buf, ph);
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
      }
      tc_fence_before();
      mbar_arrive(tempty0 + 8 * buf);
    }
  } else if (warp == TC_EPI_WARPS) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.c_out >> 3) << 17) | ((TC_BM >> 4) << 24);
      uint32_t it = 0, a_unit = 0, b_unit = 0;
      for (long long st = blockIdx.x; st < num_super; st += gridDim.x, ++it) {
        const uint32_t buf = it & 1, ph = (it >> 1) & 1;
        unsigned masks[G], umask = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const long long tile = st * G + g;
          masks[g] = tile < num_tiles ? ((p.tile_mask ? p.tile_mask[tile] : 0xffffffffu) & kmask) : 0u;
          umask |= masks[g];
        }
        mbar_wait(tempty0 + 8 * buf, ph ^ 1);
        tc_fence_after();
        unsigned started = 0;
        for (int k = next_bit(umask, -1); k < 32; k = next_bit(umask, k)) {
          for (int j = 0; j < KB; ++j, ++b_unit) {
            const uint32_t bs = b_unit % p.nb, bph = (b_unit / p.nb) & 1;
            const int kc = j < p.kb0 ? min(TC_KB, p.c0 - j * TC_KB) : min(TC_KB, p.c1 - (j - p.kb0) * TC_KB);
            mbar_wait(bfull0 + 8 * bs, bph);
            const uint32_t b_addr = smem_base + bs * b_bytes;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              if (!((masks[g] >> k) & 1u)) continue;
              const uint32_t as = a_unit % p.na, aph = (a_unit / p.na) & 1;
              ++a_unit;
              mbar_wait(afull0 + 8 * as, aph);
              fence_async_proxy();  // cp.async wrote through the generic proxy; the MMA reads through the async proxy
              tc_fence_after();
              const uint32_t a_addr = a_base + as * TC_A_BYTES;
              const uint32_t d_tmem = tmem_base + (buf * G + g) * (uint32_t)p.c_out;
              for (int ks = 0; ks < kc; ks += 16) {
                umma_bf16(d_tmem, umma_desc(a_addr + ks * 2), umma_desc(b_addr + ks * 2), idesc, (started >> g) & 1u);
                started |= 1u << g;
              }
              umma_commit(aempty0 + 8 * as);  // frees the A slot once these MMAs have read it
            }
            umma_commit(bempty0 + 8 * bs);
          }
        }
        umma_commit(tfull0 + 8 * buf);  // accumulators complete (arrives immediately if the super tile had no work)
      }
    }
    __syncwarp();
  } else {
    // ================================================================= producers
    const int pt = threadIdx.x - 32 * (TC_EPI_WARPS + 1);  // 0..127
    const int chunk = pt & 7, rsub = pt >> 3;              // 8 lanes cover one 128-byte row; 16 rows per pass
    uint32_t a_unit = 0, b_unit = 0;
    for (long long st = blockIdx.x; st < num_super; st += gridDim.x) {
      unsigned masks[G], umask = 0;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const long long tile = st * G + g;
        masks[g] = tile < num_tiles ? ((p.tile_mask ? p.tile_mask[tile] : 0xffffffffu) & kmask) : 0u;
        umask |= masks[g];
      }
      const long long m0 = st * G * TC_BM;
      int cur[G][8], nxt[G][8];
      auto load_idx = [&](int (&dst)[G][8], int k) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const bool act = (masks[g] >> k) & 1u;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const long long o = m0 + g * TC_BM + rsub + 16 * q;
            dst[g][q] = (act && o < p.n_out) ? (p.nbr ? __ldg(p.nbr + (long long)k * p.n_out + o) : (int)o) : -1;
          }
        }
      };
      int k = next_bit(umask, -1);
      if (k < 32) load_idx(cur, k);
      while (k < 32) {
        const int kn = next_bit(umask, k);
        if (kn < 32) load_idx(nxt, kn);  // prefetch the next offset's neighbour rows while this one is issued
        for (int j = 0; j < KB; ++j, ++b_unit) {
          const bool second = j >= p.kb0;
          const __nv_bfloat16 *base = second ? p.in1 : p.in0;
          const int cs = second ? p.c1 : p.c0;
          const int ch0 = (second ? j - p.kb0 : j) * TC_KB;
          if (pt == 0) {
            const uint32_t bs = b_unit % p.nb, bph = (b_unit / p.nb) & 1;
            mbar_wait(bempty0 + 8 * bs, bph ^ 1);
            mbar_arrive_expect_tx(bfull0 + 8 * bs, b_bytes);
            bulk_g2s(smem_base + bs * b_bytes, p.packed_w + ((size_t)k * KB + j) * b_bytes, b_bytes, bfull0 + 8 * bs);
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (!((masks[g] >> k) & 1u)) continue;
            const uint32_t as = a_unit % p.na, aph = (a_unit / p.na) & 1;
            ++a_unit;
            mbar_wait(aempty0 + 8 * as, aph ^ 1);
            const uint32_t a_addr = a_base + as * TC_A_BYTES;
            if (ch0 + chunk * 8 < cs) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int r = rsub + 16 * q;
                const uint32_t dst = a_addr + r * 128 + ((chunk ^ (r & 7)) << 4);
                const bool ok = cur[g][q] >= 0;
                const __nv_bfloat16 *gp = ok ? base + (size_t)cur[g][q] * cs + ch0 + chunk * 8 : base;
                cp_async16(dst, gp, ok ? 16u : 0u);
              }
            }
            cp_async_arrive(afull0 + 8 * as);
          }
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int q = 0; q < 8; ++q) cur[g][q] = nxt[g][q];
        k = kn;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
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

}  // extern "C"

namespace tsg {
// cp.async producer variant (kept as the fallback when a TMA tensor map cannot be encoded, and for A/B timing)
int conv_fwd_tc_cpasync(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
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
  const int sms = num_sms_hint > 0 ? num_sms_hint : num_sms();
  const long long num_tiles = (n_out + TC_BM - 1) / TC_BM;
  // G sub-tiles share every weight slice; bounded by TMEM (2 buffers x G x c_out fp32 columns <= 512) and by the
  // number of super tiles needed to keep every SM busy
  int G = c_out <= 64 ? 4 : (c_out <= 128 ? 2 : 1);
  while (G > 1 && (num_tiles + G - 1) / G < 2LL * sms) G >>= 1;
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)G * (uint32_t)c_out) cols <<= 1;
  p.tmem_cols = cols;
  const size_t b_bytes = (size_t)c_out * 128, budget = 200 * 1024;
  p.nb = b_bytes * 3 + 4 * TC_A_BYTES <= budget ? 3 : 2;
  int na = (int)((budget - p.nb * b_bytes) / TC_A_BYTES);
  if (na > TC_MAX_A) na = TC_MAX_A;
  if (na < 2) {
    set_error("tsg_conv_fwd_tc: not enough shared memory for the pipeline");
    return TSG_ERR_UNSUPPORTED;
  }
  p.na = na;
  const size_t smem = p.nb * b_bytes + (size_t)na * TC_A_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    TSG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    TSG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    TSG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  const long long num_super = (num_tiles + G - 1) / G;
  const unsigned grid = (unsigned)(num_super < sms ? num_super : sms);
  if (G == 4) conv_tc_kernel<4><<<grid, TC_THREADS, smem, stream>>>(p);
  else if (G == 2) conv_tc_kernel<2><<<grid, TC_THREADS, smem, stream>>>(p);
  else conv_tc_kernel<1><<<grid, TC_THREADS, smem, stream>>>(p);
  return check_launch("tsg_conv_fwd_tc(cp.async)");
}
}  // namespace tsg
