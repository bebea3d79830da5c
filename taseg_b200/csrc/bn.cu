// BatchNorm over the (N, C) feature matrix of a sparse tensor (spnn.BatchNorm = nn.BatchNorm1d on SparseTensor.F,
// TS/torchsparse/nn/modules/norm.py:10-13), training and evaluation, forward and backward.
//
// The reference runs ATen's batch_norm; on (N, C) rows with C = 32..256 its channels-last kernels reach about a tenth of
// the HBM bandwidth (16 ms of a 57 ms training step of the benchmark network: profiles/README.md).  Here every pass is a
// streaming kernel over 16-byte vectors of 8 channels: a thread owns one 8-channel group of a row stripe, so a warp reads
// consecutive bytes, accumulates in fp32 registers and meets the other threads of its channel group once per CTA in shared
// memory.  Per-CTA partials (count, mean, M2 — or plain sums for the gradient) are combined in double precision, in CTA
// order, by a one-CTA finalize kernel: results are deterministic and do not suffer the E[x^2] - E[x]^2 cancellation.
//   forward :  bn_stats -> bn_finalize (mean, invstd, running statistics) -> bn_apply  [y = (x - mean) invstd gamma + beta]
//   backward:  bn_bwd_stats [s1 = sum dy, s2 = sum dy xhat] -> bn_bwd_finalize [dgamma = s2, dbeta = s1] -> bn_bwd_apply
//              [dx = gamma invstd (dy - s1 / N - xhat s2 / N)]
#include <cuda_bf16.h>

#include "common.cuh"

namespace tsg {

constexpr int BN_THREADS = 256;
constexpr int BN_MAXC = 1024;

template <typename T>
__device__ __forceinline__ void load8(const T *p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float *p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16 *p, float (&v)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = __uint_as_float(w[j] << 16);
    v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
template <typename T>
__device__ __forceinline__ void store8(T *p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float *p, const float (&v)[8]) {
  reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16 *p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t *>(&h);
  }
  *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// Rows are dealt to CTAs in stripes: CTA b owns rows [b * rows_per_cta, (b + 1) * rows_per_cta).  Inside a CTA thread t
// handles channel group t % tpr of rows (t / tpr) + i * rpp.  partial[b] = {count, then per channel: a, b}.
// MODE 0: a = mean of the stripe, b = M2 (sum of squared deviations).  MODE 1: a = sum dy, b = sum dy * xhat.
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS) bn_reduce_kernel(const T *__restrict__ x, const T *__restrict__ dy, int64_t n, int c,
                                                               int64_t rows_per_cta, const float *__restrict__ mean,
                                                               const float *__restrict__ invstd, float *__restrict__ partial) {
  extern __shared__ float s_acc[];   // [rpp][2][c]: every thread's sums, added up in a fixed order below
  const int tpr = c / 8, rpp = BN_THREADS / tpr;
  const int rt = threadIdx.x / tpr, cg = threadIdx.x - rt * tpr;
  const int64_t r0 = blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
  float a[8], b[8], mu[8], is[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
  if (rt < rpp) {
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mu[j] = mean[cg * 8 + j];
        is[j] = invstd[cg * 8 + j];
      }
    }
    for (int64_t r = r0 + rt; r < r1; r += rpp) {
      float v[8];
      load8<T>(x + r * c + cg * 8, v);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[j] += v[j];
          b[j] = fmaf(v[j], v[j], b[j]);
        }
      } else {
        float g[8];
        load8<T>(dy + r * c + cg * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[j] += g[j];
          b[j] = fmaf(g[j], (v[j] - mu[j]) * is[j], b[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_acc[(rt * 2) * c + cg * 8 + j] = a[j];
      s_acc[(rt * 2 + 1) * c + cg * 8 + j] = b[j];
    }
  }
  __syncthreads();
  float *out = partial + (int64_t)blockIdx.x * (1 + 2 * c);
  const float cnt = (float)max((int64_t)0, r1 - r0);
  if (threadIdx.x == 0) out[0] = cnt;
  for (int i = threadIdx.x; i < c; i += BN_THREADS) {
    float s = 0.f, q = 0.f;
    for (int t = 0; t < rpp; ++t) {
      s += s_acc[(t * 2) * c + i];
      q += s_acc[(t * 2 + 1) * c + i];
    }
    if (MODE == 0) {
      const float m = cnt > 0.f ? s / cnt : 0.f;
      out[1 + i] = m;
      out[1 + c + i] = fmaxf(q - s * m, 0.f);   // M2 of the stripe (a few thousand rows: fp32 is enough here)
    } else {
      out[1 + i] = s;
      out[1 + c + i] = q;
    }
  }
}

// One warp per channel: lane l combines stripes l, l + 32, ... in order (Chan's update, double precision), then the 32
// lane results are combined by a shuffle tree — a fixed order, so the statistics are deterministic.
__device__ __forceinline__ void chan_merge(double &n, double &m, double &m2, double nb, double mb, double m2b) {
  if (nb <= 0.0) return;
  const double nt = n + nb, d = mb - m;
  m += d * nb / nt;
  m2 += m2b + d * d * n * nb / nt;
  n = nt;
}

__global__ void __launch_bounds__(256) bn_finalize_kernel(const float *__restrict__ partial, int nb, int c, float eps, float momentum,
                                                          float *__restrict__ running_mean, float *__restrict__ running_var,
                                                          float *__restrict__ mean, float *__restrict__ invstd) {
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (ch >= c) return;
  double n = 0.0, m = 0.0, m2 = 0.0;
  for (int b = lane; b < nb; b += 32) {
    const float *p = partial + (int64_t)b * (1 + 2 * c);
    chan_merge(n, m, m2, (double)p[0], (double)p[1 + ch], (double)p[1 + c + ch]);
  }
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {   // lane l absorbs lane l + d: stripes stay in ascending groups
    const double on = __shfl_down_sync(0xffffffffu, n, d), om = __shfl_down_sync(0xffffffffu, m, d),
                 om2 = __shfl_down_sync(0xffffffffu, m2, d);
    if ((lane & (2 * d - 1)) == 0) chan_merge(n, m, m2, on, om, om2);
  }
  if (lane == 0) {
    const double var = n > 0.0 ? m2 / n : 0.0;
    mean[ch] = (float)m;
    invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
    if (running_var) running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)(n > 1.0 ? m2 / (n - 1.0) : var);
  }
}

// column sums of the gradient partials (one warp per channel, fixed order): sums[0][c] = sum dy (= dbeta),
// sums[1][c] = sum dy xhat (= dgamma)
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float *__restrict__ partial, int nb, int c, float *__restrict__ sums) {
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (ch >= c) return;
  double s1 = 0.0, s2 = 0.0;
  for (int b = lane; b < nb; b += 32) {
    const float *p = partial + (int64_t)b * (1 + 2 * c);
    s1 += p[1 + ch];
    s2 += p[1 + c + ch];
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, d);
    s2 += __shfl_down_sync(0xffffffffu, s2, d);
  }
  if (lane == 0) {
    sums[ch] = (float)s1;
    sums[c + ch] = (float)s2;
  }
}

// MODE 0: y = (x - mean) invstd gamma + beta (eval: mean / invstd from the running statistics).
// MODE 1: dx = gamma invstd (dy - s1 / n - xhat s2 / n)   (training backward);  MODE 2: dx = gamma invstd dy (eval backward)
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const T *__restrict__ x, const T *__restrict__ dy, int64_t n, int c,
                                                              const float *__restrict__ mean, const float *__restrict__ invstd,
                                                              const float *__restrict__ gamma, const float *__restrict__ beta,
                                                              const float *__restrict__ sums, T *__restrict__ out) {
  const int tpr = c / 8;
  const int64_t total = n * tpr;
  const float inv_n = 1.f / (float)n;
  for (int64_t t = blockIdx.x * (int64_t)BN_THREADS + threadIdx.x; t < total; t += (int64_t)gridDim.x * BN_THREADS) {
    const int cg = (int)(t % tpr);
    float v[8], o[8];
    if (MODE != 2) load8<T>(x + t * 8, v);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = cg * 8 + j;
        const float sc = invstd[ch] * (gamma ? gamma[ch] : 1.f);
        o[j] = fmaf(v[j] - mean[ch], sc, beta ? beta[ch] : 0.f);
      }
    } else {
      float g[8];
      load8<T>(dy + t * 8, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = cg * 8 + j;
        const float sc = invstd[ch] * (gamma ? gamma[ch] : 1.f);
        if (MODE == 1) {
          const float xh = (v[j] - mean[ch]) * invstd[ch];
          o[j] = sc * (g[j] - sums[ch] * inv_n - xh * sums[c + ch] * inv_n);
        } else {
          o[j] = sc * g[j];
        }
      }
    }
    store8<T>(out + t * 8, o);
  }
}

static int bn_blocks(int64_t n, int c, int64_t *rows_per_cta) {
  const int rpp = BN_THREADS / (c / 8);
  int64_t nb = 2LL * num_sms();
  int64_t rows = (n + nb - 1) / nb;
  rows = (rows + rpp - 1) / rpp * rpp;
  if (rows < 4 * rpp) rows = 4 * rpp;
  *rows_per_cta = rows;
  return (int)((n + rows - 1) / rows);
}

}  // namespace tsg

using namespace tsg;

extern "C" {

/* workspace of the forward / backward reductions: one {count, 2 x C} record per CTA */
size_t tsg_bn_ws_bytes(int64_t n, int c) {
  if (c <= 0 || c % 8 || c > BN_MAXC || n <= 0) return 0;
  int64_t rows;
  return (size_t)bn_blocks(n, c, &rows) * (1 + 2 * (size_t)c) * sizeof(float);
}

/* Training forward statistics: mean / invstd (C fp32 each) of x (n, c), c a multiple of 8 <= 1024, dtype fp32 or bf16;
 * running_mean / running_var (may be NULL) are updated as nn.BatchNorm1d does (momentum, unbiased variance). */
int tsg_bn_stats(const void *x, int dtype, int64_t n, int c, float eps, float momentum, float *running_mean, float *running_var,
                 float *mean, float *invstd, void *ws, size_t ws_bytes, tsg_stream_t stream) {
  if (c <= 0 || c % 8 || c > BN_MAXC || n <= 0 || (dtype != TSG_F32 && dtype != TSG_BF16)) {
    set_error("tsg_bn_stats: need n > 0, c a multiple of 8 <= 1024, fp32 or bf16 rows");
    return TSG_ERR_UNSUPPORTED;
  }
  if (ws_bytes < tsg_bn_ws_bytes(n, c)) {
    set_error("tsg_bn_stats: workspace too small");
    return TSG_ERR_INVALID;
  }
  int64_t rows;
  const int nb = bn_blocks(n, c, &rows);
  const size_t smem = 2 * (size_t)c * (BN_THREADS / (c / 8)) * sizeof(float);
  if (dtype == TSG_F32)
    bn_reduce_kernel<float, 0><<<nb, BN_THREADS, smem, stream>>>((const float *)x, nullptr, n, c, rows, nullptr, nullptr, (float *)ws);
  else
    bn_reduce_kernel<__nv_bfloat16, 0><<<nb, BN_THREADS, smem, stream>>>((const __nv_bfloat16 *)x, nullptr, n, c, rows, nullptr,
                                                                        nullptr, (float *)ws);
  bn_finalize_kernel<<<(c + 7) / 8, 256, 0, stream>>>((const float *)ws, nb, c, eps, momentum, running_mean, running_var, mean,
                                                         invstd);
  return check_launch("tsg_bn_stats");
}

/* y = (x - mean) * invstd * gamma + beta; gamma / beta may be NULL (affine=False) */
int tsg_bn_apply(const void *x, int dtype, int64_t n, int c, const float *mean, const float *invstd, const float *gamma,
                 const float *beta, void *y, tsg_stream_t stream) {
  if (c <= 0 || c % 8 || n <= 0 || (dtype != TSG_F32 && dtype != TSG_BF16)) {
    set_error("tsg_bn_apply: need n > 0, c a multiple of 8, fp32 or bf16 rows");
    return TSG_ERR_UNSUPPORTED;
  }
  const int grid = grid_for(n * (c / 8), BN_THREADS);
  if (dtype == TSG_F32)
    bn_apply_kernel<float, 0><<<grid, BN_THREADS, 0, stream>>>((const float *)x, nullptr, n, c, mean, invstd, gamma, beta, nullptr,
                                                             (float *)y);
  else
    bn_apply_kernel<__nv_bfloat16, 0><<<grid, BN_THREADS, 0, stream>>>((const __nv_bfloat16 *)x, nullptr, n, c, mean, invstd, gamma,
                                                                     beta, nullptr, (__nv_bfloat16 *)y);
  return check_launch("tsg_bn_apply");
}

/* Backward.  training != 0: sums (2 x C fp32) receives {dbeta = sum dy, dgamma = sum dy * xhat} and
 * dx = gamma invstd (dy - sums[0] / n - xhat sums[1] / n);  training == 0 (running statistics were used): the same sums,
 * dx = gamma invstd dy.  dx may be NULL (only the parameter gradients are wanted). */
int tsg_bn_backward(const void *x, const void *dy, int dtype, int64_t n, int c, const float *mean, const float *invstd,
                    const float *gamma, int training, float *sums, void *dx, void *ws, size_t ws_bytes, tsg_stream_t stream) {
  if (c <= 0 || c % 8 || c > BN_MAXC || n <= 0 || (dtype != TSG_F32 && dtype != TSG_BF16)) {
    set_error("tsg_bn_backward: need n > 0, c a multiple of 8 <= 1024, fp32 or bf16 rows");
    return TSG_ERR_UNSUPPORTED;
  }
  if (ws_bytes < tsg_bn_ws_bytes(n, c)) {
    set_error("tsg_bn_backward: workspace too small");
    return TSG_ERR_INVALID;
  }
  int64_t rows;
  const int nb = bn_blocks(n, c, &rows);
  const size_t smem = 2 * (size_t)c * (BN_THREADS / (c / 8)) * sizeof(float);
  const int grid = grid_for(n * (c / 8), BN_THREADS);
  if (dtype == TSG_F32) {
    bn_reduce_kernel<float, 1><<<nb, BN_THREADS, smem, stream>>>((const float *)x, (const float *)dy, n, c, rows, mean, invstd, (float *)ws);
    bn_bwd_finalize_kernel<<<(c + 7) / 8, 256, 0, stream>>>((const float *)ws, nb, c, sums);
    if (dx && training)
      bn_apply_kernel<float, 1><<<grid, BN_THREADS, 0, stream>>>((const float *)x, (const float *)dy, n, c, mean, invstd, gamma, nullptr,
                                                               sums, (float *)dx);
    else if (dx)
      bn_apply_kernel<float, 2><<<grid, BN_THREADS, 0, stream>>>((const float *)x, (const float *)dy, n, c, mean, invstd, gamma, nullptr,
                                                               sums, (float *)dx);
  } else {
    bn_reduce_kernel<__nv_bfloat16, 1><<<nb, BN_THREADS, smem, stream>>>((const __nv_bfloat16 *)x, (const __nv_bfloat16 *)dy, n, c, rows,
                                                                        mean, invstd, (float *)ws);
    bn_bwd_finalize_kernel<<<(c + 7) / 8, 256, 0, stream>>>((const float *)ws, nb, c, sums);
    if (dx && training)
      bn_apply_kernel<__nv_bfloat16, 1><<<grid, BN_THREADS, 0, stream>>>((const __nv_bfloat16 *)x, (const __nv_bfloat16 *)dy, n, c, mean,
                                                                       invstd, gamma, nullptr, sums, (__nv_bfloat16 *)dx);
    else if (dx)
      bn_apply_kernel<__nv_bfloat16, 2><<<grid, BN_THREADS, 0, stream>>>((const __nv_bfloat16 *)x, (const __nv_bfloat16 *)dy, n, c, mean,
                                                                       invstd, gamma, nullptr, sums, (__nv_bfloat16 *)dx);
  }
  return check_launch("tsg_bn_backward");
}

}  // extern "C"
