// Sparse convolution, CUDA-core path: output-stationary implicit GEMM over the kernel-map table.
// This is the exact-fp32 path used for parity runs, odd channel counts and training (dgrad/wgrad); the bf16
// tcgen05/TMEM path lives in conv_tc.cu.  One launch per convolution: rows are gathered straight into shared
// memory (no gather buffer), all K offsets accumulate in registers (no scatter, no atomics, deterministic).
#include <mma.h>

#include "common.cuh"

namespace tsg {

constexpr int BM = 64, BN = 64, BK = 16;

// out[o, n0:n0+64] = epilogue( sum_k in[nbr[k,o], :] @ W[k][:, n0:n0+64] )
// WT: weights are addressed transposed (W[k] is (c_out_w, c_in_w) row-major and we need its transpose) -> dgrad.
template <typename T, bool WT>
__global__ void __launch_bounds__(256) conv_fwd_kernel(const T *__restrict__ in, int c_in, const T *__restrict__ w,
                                                       int K, int c_out, const int *__restrict__ nbr, int64_t n_out,
                                                       T *__restrict__ out, const float *__restrict__ scale,
                                                       const float *__restrict__ bias, const T *__restrict__ residual,
                                                       int relu) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int s_idx[BM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_row = tid >> 2, a_col = (tid & 3) * 4;   // A tile: 64 rows x 16 cols, 4 consecutive cols per thread
  const int b_row = tid >> 4, b_col = (tid & 15) * 4;  // B tile: 16 rows x 64 cols

  for (int k = 0; k < K; ++k) {
    int my = -1;
    if (tid < BM) {
      const int64_t o = m0 + tid;
      my = o < n_out ? nbr[(int64_t)k * n_out + o] : -1;
      s_idx[tid] = my;
    }
    if (!__syncthreads_or(my >= 0)) continue;  // nobody in this tile has a neighbour at offset k
    const int src = s_idx[a_row];
    const T *wk = w + (int64_t)k * c_in * c_out;
    for (int kb = 0; kb < c_in; kb += BK) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = kb + a_col + j;
        As[a_col + j][a_row] = (src >= 0 && c < c_in) ? to_f32(in[(int64_t)src * c_in + c]) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = kb + b_row, co = n0 + b_col + j;
        float v = 0.f;
        if (ci < c_in && co < c_out) v = to_f32(WT ? wk[(int64_t)co * c_in + ci] : wk[(int64_t)ci * c_out + co]);
        Bs[b_row][b_col + j] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t o = m0 + ty * 4 + i;
    if (o >= n_out) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= c_out) continue;
      float v = acc[i][j];
      if (scale) v *= scale[co];
      if (bias) v += bias[co];
      if (residual) v += to_f32(residual[o * c_out + co]);
      if (relu) v = fmaxf(v, 0.f);
      out[o * c_out + co] = from_f32<T>(v);
    }
  }
}

// grad_w[k][ci][co] += sum over a slab of output rows of in[nbr[k,o]][ci] * gy[o][co]
// grid = (K * splits, ceil(c_in/32), ceil(c_out/32)); 32x32 tile per CTA, 2x2 per thread, fp32 atomics across slabs.
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float *__restrict__ in, int c_in,
                                                         const float *__restrict__ gy, int c_out,
                                                         const int *__restrict__ nbr, int64_t n_out, int splits,
                                                         float *__restrict__ gw) {
  __shared__ float Xs[32][33];
  __shared__ float Gs[32][33];
  __shared__ int s_idx[32];
  const int k = blockIdx.x / splits, sp = blockIdx.x % splits;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.z * 32;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t rows_per = ((n_out + splits - 1) / splits + 31) / 32 * 32;
  const int64_t r_begin = sp * rows_per, r_end = min(n_out, r_begin + rows_per);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  const int lr = tid >> 3, lc = (tid & 7) * 4;  // loader: 32 rows x 32 cols, 4 cols per thread
  for (int64_t r0 = r_begin; r0 < r_end; r0 += 32) {
    int my = -1;
    if (tid < 32) {
      const int64_t o = r0 + tid;
      my = o < r_end ? nbr[(int64_t)k * n_out + o] : -1;
      s_idx[tid] = my;
    }
    if (!__syncthreads_or(my >= 0)) continue;
    const int src = s_idx[lr];
    const int64_t o = r0 + lr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + lc + j, co = co0 + lc + j;
      Xs[lr][lc + j] = (src >= 0 && ci < c_in) ? in[(int64_t)src * c_in + ci] : 0.f;
      Gs[lr][lc + j] = (src >= 0 && co < c_out) ? gy[o * c_out + co] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float x0 = Xs[r][ty * 2], x1 = Xs[r][ty * 2 + 1], g0 = Gs[r][tx * 2], g1 = Gs[r][tx * 2 + 1];
      acc[0][0] = fmaf(x0, g0, acc[0][0]); acc[0][1] = fmaf(x0, g1, acc[0][1]);
      acc[1][0] = fmaf(x1, g0, acc[1][0]); acc[1][1] = fmaf(x1, g1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ci = ci0 + ty * 2 + i, co = co0 + tx * 2 + j;
      if (ci < c_in && co < c_out && acc[i][j] != 0.f)
        atomicAdd(&gw[((int64_t)k * c_in + ci) * c_out + co], acc[i][j]);
    }
}

// ---------------------------------------------------------------- weight gradient on the tensor cores (bf16 operands)
// grad_w[k] (c_in x c_out, fp32) = sum over the pairs (i, o) of offset k of in[i,:]^T gy[o,:].
// grid = (K, splits, ceil(c_in/128) * ceil(c_out/128)), 256 threads = 8 warps as 4 (ci) x 2 (co), each warp owns a
// 32 x 64 block of the CTA's 128 x 128 tile = 2 x 4 accumulator fragments (fp32).  The CTA scans its slab of output
// rows 256 at a time, COMPACTS the rows that have a neighbour at offset k into a pair list in shared memory (warp
// ballots; order preserved, so sums are deterministic per split) and multiplies 64 pairs at a time:
//   A = X^T  (ci x pair)  read column-major from Xs[pair][ci],   B = gY (pair x co) read row-major from Gs[pair][co],
// warp-level bf16 MMAs (mma.sync through the WMMA API — a GEMM whose K dimension is the gathered pair list does not
// fit the tile-row pipeline of conv_tc.cu; tcgen05 for this kernel is future work).  Partial tiles of the splits are
// added to grad_w with fp32 atomics.  The fp32 kernel above stays the exact path for parity runs.
constexpr int WG_TILE = 128, WG_PAIRS = 64, WG_PITCH = WG_TILE + 8;  // +8 bf16: rows 272 B apart, conflict-free fragment loads

__global__ void __launch_bounds__(256) conv_wgrad_bf16_kernel(const __nv_bfloat16 *__restrict__ in, int c_in,
                                                              const __nv_bfloat16 *__restrict__ gy, int c_out,
                                                              const int *__restrict__ nbr, int64_t n_out, int splits,
                                                              int co_tiles, float *__restrict__ gw) {
  using namespace nvcuda;
  __shared__ __align__(32) __nv_bfloat16 Xs[WG_PAIRS][WG_PITCH];
  __shared__ __align__(32) __nv_bfloat16 Gs[WG_PAIRS][WG_PITCH];
  __shared__ int p_src[WG_PAIRS + 256], p_out[WG_PAIRS + 256];
  __shared__ int warp_cnt[8];
  __shared__ __align__(32) float scratch[8][16 * 16];
  const int k = blockIdx.x, sp = blockIdx.y;
  const int ci0 = (blockIdx.z / co_tiles) * WG_TILE, co0 = (blockIdx.z % co_tiles) * WG_TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wi = warp >> 1, wj = warp & 1;  // warp's 32 x 64 block: ci tiles 2 wi, 2 wi + 1; co tiles 4 wj .. 4 wj + 3
  const int64_t rows_per = ((n_out + splits - 1) / splits + 255) / 256 * 256;
  const int64_t r_begin = sp * rows_per, r_end = min(n_out, r_begin + rows_per);
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) wmma::fill_fragment(acc[a][b], 0.f);

  // multiply the first `cnt` (<= 64) pairs of the list; rows beyond cnt are zero
  auto multiply = [&](int cnt) {
    // gather: 64 pairs x 128 channels of each operand, 16-byte chunks (8 channels), 4 + 4 chunks per thread
    for (int c = tid; c < WG_PAIRS * (WG_TILE / 8); c += 256) {
      const int r = c >> 4, q = c & 15;
      uint4 x = make_uint4(0u, 0u, 0u, 0u), g = make_uint4(0u, 0u, 0u, 0u);
      if (r < cnt) {
        if (ci0 + q * 8 < c_in) x = __ldg(reinterpret_cast<const uint4 *>(in + (int64_t)p_src[r] * c_in + ci0 + q * 8));
        if (co0 + q * 8 < c_out) g = __ldg(reinterpret_cast<const uint4 *>(gy + (int64_t)p_out[r] * c_out + co0 + q * 8));
      }
      *reinterpret_cast<uint4 *>(&Xs[r][q * 8]) = x;
      *reinterpret_cast<uint4 *>(&Gs[r][q * 8]) = g;
    }
    __syncthreads();
    if (ci0 + wi * 32 < c_in && co0 + wj * 64 < c_out) {  // warps whose whole block lies outside the layer skip the math
#pragma unroll
      for (int ks = 0; ks < WG_PAIRS; ks += 16) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> fa[2];
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fb[4];
#pragma unroll
        for (int a = 0; a < 2; ++a) wmma::load_matrix_sync(fa[a], &Xs[ks][wi * 32 + a * 16], WG_PITCH);
#pragma unroll
        for (int b = 0; b < 4; ++b) wmma::load_matrix_sync(fb[b], &Gs[ks][wj * 64 + b * 16], WG_PITCH);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) wmma::mma_sync(acc[a][b], fa[a], fb[b], acc[a][b]);
      }
    }
    __syncthreads();
  };

  int count = 0;  // pairs waiting in the list (uniform across the CTA)
  for (int64_t r0 = r_begin; r0 < r_end; r0 += 256) {
    const int64_t o = r0 + tid;
    const int src = o < r_end ? __ldg(nbr + (int64_t)k * n_out + o) : -1;
    const unsigned bal = __ballot_sync(0xffffffffu, src >= 0);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int base = count, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) base += warp_cnt[w];
      total += warp_cnt[w];
    }
    if (src >= 0) {
      const int pos = base + __popc(bal & ((1u << lane) - 1u));
      p_src[pos] = src;
      p_out[pos] = (int)o;
    }
    count += total;
    __syncthreads();
    while (count >= WG_PAIRS) {
      multiply(WG_PAIRS);
      // shift the rest of the list down (multiply() ended with a barrier: nobody reads the old positions any more)
      const int rest = count - WG_PAIRS;
      int s0 = -1, s1 = -1;
      if (tid < rest) {
        s0 = p_src[WG_PAIRS + tid];
        s1 = p_out[WG_PAIRS + tid];
      }
      __syncthreads();
      if (tid < rest) {
        p_src[tid] = s0;
        p_out[tid] = s1;
      }
      count = rest;
      __syncthreads();
    }
  }
  if (count > 0) multiply(count);

  // add this split's tile to grad_w
  if (ci0 + wi * 32 < c_in && co0 + wj * 64 < c_out) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        wmma::store_matrix_sync(scratch[warp], acc[a][b], 16, wmma::mem_row_major);
        __syncwarp();
        const int ci_base = ci0 + wi * 32 + a * 16, co_base = co0 + wj * 64 + b * 16;
        for (int e = lane; e < 256; e += 32) {
          const int ci = ci_base + (e >> 4), co = co_base + (e & 15);
          const float v = scratch[warp][e];
          if (ci < c_in && co < c_out && v != 0.f) atomicAdd(&gw[((int64_t)k * c_in + ci) * c_out + co], v);
        }
        __syncwarp();
      }
  }
}

}  // namespace tsg

using namespace tsg;

extern "C" {

int tsg_conv_fwd(const void *in, int dtype, int64_t n_in, int c_in, const void *weight, int k, int c_in_w, int c_out,
                 const int32_t *nbr, int64_t n_out, void *out, const float *scale, const float *bias,
                 const void *residual, int relu, tsg_stream_t stream) {
  (void)n_in;
  if (c_in != c_in_w) {
    set_error("Input feature size and kernel size mismatch");
    return TSG_ERR_INVALID;
  }
  if (n_out <= 0 || c_out <= 0) return TSG_OK;
  dim3 grid((unsigned)((n_out + BM - 1) / BM), (unsigned)((c_out + BN - 1) / BN));
  switch (dtype) {
    case TSG_F32:
      conv_fwd_kernel<float, false><<<grid, 256, 0, stream>>>((const float *)in, c_in, (const float *)weight, k, c_out,
                                                              nbr, n_out, (float *)out, scale, bias,
                                                              (const float *)residual, relu);
      break;
    case TSG_BF16:
      conv_fwd_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(
          (const __nv_bfloat16 *)in, c_in, (const __nv_bfloat16 *)weight, k, c_out, nbr, n_out, (__nv_bfloat16 *)out,
          scale, bias, (const __nv_bfloat16 *)residual, relu);
      break;
    case TSG_F16:
      conv_fwd_kernel<__half, false><<<grid, 256, 0, stream>>>((const __half *)in, c_in, (const __half *)weight, k,
                                                               c_out, nbr, n_out, (__half *)out, scale, bias,
                                                               (const __half *)residual, relu);
      break;
    default:
      set_error("tsg_conv_fwd: unknown dtype %d", dtype);
      return TSG_ERR_INVALID;
  }
  return check_launch("tsg_conv_fwd");
}

int tsg_conv_dgrad(const float *grad_out, int64_t n_out, int c_out, const float *weight, int k, int c_in,
                   const int32_t *nbr_t, int64_t n_in, float *grad_in, tsg_stream_t stream) {
  (void)n_out;
  if (n_in <= 0 || c_in <= 0) return TSG_OK;
  // grad_in[i,:] = sum_k grad_out[nbr_t[k,i],:] @ W[k]^T : "input" channels = c_out, "output" channels = c_in
  dim3 grid((unsigned)((n_in + BM - 1) / BM), (unsigned)((c_in + BN - 1) / BN));
  conv_fwd_kernel<float, true><<<grid, 256, 0, stream>>>(grad_out, c_out, weight, k, c_in, nbr_t, n_in, grad_in,
                                                         nullptr, nullptr, nullptr, 0);
  return check_launch("tsg_conv_dgrad");
}

int tsg_conv_wgrad(const float *in, int64_t n_in, int c_in, const float *grad_out, int64_t n_out, int c_out,
                   const int32_t *nbr, int k, float *grad_w, tsg_stream_t stream) {
  (void)n_in;
  TSG_CUDA(cudaMemsetAsync(grad_w, 0, (size_t)k * c_in * c_out * sizeof(float), stream));
  if (n_out <= 0 || k <= 0) return TSG_OK;
  const int tiles = ((c_in + 31) / 32) * ((c_out + 31) / 32);
  int splits = (int)((4LL * num_sms() + (int64_t)k * tiles - 1) / ((int64_t)k * tiles));
  const int64_t max_splits = (n_out + 511) / 512;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)(k * splits), (unsigned)((c_in + 31) / 32), (unsigned)((c_out + 31) / 32));
  conv_wgrad_kernel<<<grid, 256, 0, stream>>>(in, c_in, grad_out, c_out, nbr, n_out, splits, grad_w);
  return check_launch("tsg_conv_wgrad");
}

/* bf16 operands (c_in, c_out multiples of 8), fp32 grad_w: the autocast training path. */
int tsg_conv_wgrad_bf16(const void *in, int64_t n_in, int c_in, const void *grad_out, int64_t n_out, int c_out,
                        const int32_t *nbr, int k, float *grad_w, tsg_stream_t stream) {
  (void)n_in;
  if (c_in % 8 || c_out % 8 || c_in <= 0 || c_out <= 0) {
    set_error("tsg_conv_wgrad_bf16: c_in and c_out must be multiples of 8");
    return TSG_ERR_UNSUPPORTED;
  }
  TSG_CUDA(cudaMemsetAsync(grad_w, 0, (size_t)k * c_in * c_out * sizeof(float), stream));
  if (n_out <= 0 || k <= 0) return TSG_OK;
  const int ci_tiles = (c_in + WG_TILE - 1) / WG_TILE, co_tiles = (c_out + WG_TILE - 1) / WG_TILE;
  int splits = (int)((3LL * num_sms() + (int64_t)k * ci_tiles * co_tiles - 1) / ((int64_t)k * ci_tiles * co_tiles));
  const int64_t max_splits = (n_out + 2047) / 2048;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  dim3 grid((unsigned)k, (unsigned)splits, (unsigned)(ci_tiles * co_tiles));
  conv_wgrad_bf16_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16 *)in, c_in, (const __nv_bfloat16 *)grad_out, c_out,
                                                   nbr, n_out, splits, co_tiles, grad_w);
  return check_launch("tsg_conv_wgrad_bf16");
}

}  // extern "C"
