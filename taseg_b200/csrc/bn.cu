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

#include <algorithm>

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
                                                               const float *__restrict__ invstd, const float *__restrict__ gamma,
                                                               const float *__restrict__ beta, int relu, float *__restrict__ partial) {
  extern __shared__ float s_acc[];   // [rpp][2][c]: every thread's sums, added up in a fixed order below
  const int tpr = c / 8, rpp = BN_THREADS / tpr;
  const int rt = threadIdx.x / tpr, cg = threadIdx.x - rt * tpr;
  const int64_t r0 = blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
  float a[8], b[8], mu[8], is[8], ga[8], be[8];   // ga / be: only for the ReLU mask of a fused BatchNorm + ReLU (MODE 1)
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
  if (rt < rpp) {
    if (MODE == 0) {   // pivot = the stripe's first row: sums of (x - K) and (x - K)^2 do not cancel however far the mean is from zero
      if (r0 < r1) load8<T>(x + r0 * c + cg * 8, mu);
      else
#pragma unroll
        for (int j = 0; j < 8; ++j) mu[j] = 0.f;
    }
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mu[j] = mean[cg * 8 + j];
        is[j] = invstd[cg * 8 + j];
        ga[j] = (relu && gamma) ? gamma[cg * 8 + j] : 1.f;
        be[j] = (relu && beta) ? beta[cg * 8 + j] : 0.f;
      }
    }
    // four rows per trip: the loads are independent, so a thread keeps 4 (8 with dy) 16-byte requests in flight — with one
    // load per trip the pass ran at 1.8 TB/s, bound by latency x bytes in flight
    constexpr int U = 4;
    for (int64_t r = r0 + rt; r < r1; r += (int64_t)U * rpp) {
      float v[U][8], g[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t ru = r + (int64_t)u * rpp;
        if (ru < r1) {
          load8<T>(x + ru * c + cg * 8, v[u]);
          if (MODE == 1) load8<T>(dy + ru * c + cg * 8, g[u]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[u][j] = mu[j];   // contributes nothing (MODE 0: x - K = 0, MODE 1: xhat = 0 and dy = 0)
            g[u][j] = 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (MODE == 0) {
            const float d = v[u][j] - mu[j];
            a[j] += d;
            b[j] = fmaf(d, d, b[j]);
          } else {
            const float xh = (v[u][j] - mu[j]) * is[j];
            const float gj = (relu && fmaf(xh, ga[j], be[j]) <= 0.f) ? 0.f : g[u][j];   // dy through the fused ReLU
            a[j] += gj;
            b[j] = fmaf(gj, xh, b[j]);
          }
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
      float kk[8];
      if (r0 < r1) load8<T>(x + r0 * c + (i & ~7), kk);
      const float piv = r0 < r1 ? kk[i & 7] : 0.f;
      const float m = cnt > 0.f ? s / cnt : 0.f;          // mean of (x - K)
      out[1 + i] = piv + m;
      out[1 + c + i] = fmaxf(q - s * m, 0.f);             // M2 of the stripe, from pivoted sums
    } else {
      out[1 + i] = s;
      out[1 + c + i] = q;
    }
  }
}

constexpr int BN_FW = 32;   // warps of a finalize CTA
// Lane = channel (32 consecutive channels per CTA: coalesced reads of the partial records), warp w combines stripes
// w, w + 32, ... in order; the warp results are merged in warp order through shared memory.  Fixed order = deterministic.
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float *__restrict__ partial, int nb, int c, float eps, float momentum,
                                                          float *__restrict__ running_mean, float *__restrict__ running_var,
                                                          float *__restrict__ mean, float *__restrict__ invstd,
                                                          long long *__restrict__ batches_tracked) {
  if (batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *batches_tracked += 1;   // nn.BatchNorm1d's num_batches_tracked
  // ONE plain pass over the stripe records {n_b, mean_b[c], M2_b[c]} with the first stripe's mean as pivot K (every stripe mean is
  // within a few sigma / sqrt(n_b) of the global one, so nothing cancels): N = sum n_b, S1 = sum n_b (m_b - K),
  // S2 = sum [M2_b + n_b (m_b - K)^2] in double precision; mean = K + S1 / N, M2 = S2 - S1^2 / N.  (History: a chain of Chan
  // updates with a double-precision division per stripe and lane was 90 us per call, two plain passes 18 us.)  Fixed order =
  // deterministic.
  __shared__ double sh[BN_FW][3][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, ch = blockIdx.x * 32 + lane;
  const bool on = ch < c;
  double n = 0.0, s1 = 0.0, s2 = 0.0;
  const double K = on ? (double)__ldg(partial + 1 + ch) : 0.0;
  #pragma unroll 4
  for (int b = warp; b < nb; b += BN_FW) {
    const float *p = partial + (int64_t)b * (1 + 2 * c);
    const double nb_ = (double)__ldg(p);
    n += nb_;
    if (on) {
      const double m = (double)__ldg(p + 1 + ch) - K;
      s1 += nb_ * m;
      s2 += (double)__ldg(p + 1 + c + ch) + nb_ * m * m;
    }
  }
  sh[warp][0][lane] = n;
  sh[warp][1][lane] = s1;
  sh[warp][2][lane] = s2;
  __syncthreads();
  if (warp == 0 && on) {
    for (int w = 1; w < BN_FW; ++w) {
      n += sh[w][0][lane];
      s1 += sh[w][1][lane];
      s2 += sh[w][2][lane];
    }
    const double N = n, mu = N > 0.0 ? K + s1 / N : 0.0;
    double m2 = N > 0.0 ? s2 - s1 * s1 / N : 0.0;
    if (m2 < 0.0) m2 = 0.0;
    const double var = N > 0.0 ? m2 / N : 0.0;
    mean[ch] = (float)mu;
    invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mu;
    if (running_var) running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)(N > 1.0 ? m2 / (N - 1.0) : var);
  }
}

// column sums of the gradient partials, same geometry: sums[0][c] = sum dy (= dbeta), sums[1][c] = sum dy xhat (= dgamma)
__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float *__restrict__ partial, int nb, int c, float *__restrict__ sums) {
  __shared__ double sh[BN_FW][2][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, ch = blockIdx.x * 32 + lane;
  double s1 = 0.0, s2 = 0.0;
  if (ch < c) {
    #pragma unroll 4
    for (int b = warp; b < nb; b += BN_FW) {
      const float *p = partial + (int64_t)b * (1 + 2 * c);
      s1 += (double)__ldg(p + 1 + ch);
      s2 += (double)__ldg(p + 1 + c + ch);
    }
  }
  sh[warp][0][lane] = s1;
  sh[warp][1][lane] = s2;
  __syncthreads();
  if (warp == 0 && ch < c) {
    for (int w = 1; w < BN_FW; ++w) {
      s1 += sh[w][0][lane];
      s2 += sh[w][1][lane];
    }
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
                                                              const float *__restrict__ sums, int relu, T *__restrict__ out) {
  // A thread keeps ONE channel group for the whole pass (its 8 scales / shifts live in registers) and walks rows
  // rt, rt + rows_in_flight, ...: consecutive threads still read consecutive 16-byte vectors of a row.
  const int tpr = c / 8, rpp = BN_THREADS / tpr;
  const int rt = threadIdx.x / tpr, cg = threadIdx.x - rt * tpr;
  if (rt >= rpp) return;
  const float inv_n = 1.f / (float)n;
  float sc[8], sh[8], k2[8], sh0[8];   // MODE 0: y = x sc + sh.  MODE 1: dx = g sc - sh - (x - mu) k2 with mu folded in.  MODE 2: dx = g sc
                                       // sh0: forward shift, for the mask of a fused ReLU in MODE 1 (z = x sc + sh0 > 0)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = cg * 8 + j;
    const float is = invstd[ch], ga = gamma ? gamma[ch] : 1.f, mu = mean[ch];
    sh0[j] = (beta ? beta[ch] : 0.f) - mu * is * ga;
    if (MODE == 0) {
      sc[j] = is * ga;
      sh[j] = sh0[j];
      k2[j] = 0.f;
    } else if (MODE == 1) {
      sc[j] = is * ga;
      k2[j] = is * ga * is * sums[c + ch] * inv_n;            // coefficient of (x - mu)
      sh[j] = is * ga * sums[ch] * inv_n - mu * k2[j];        // so that dx = g sc - sh - x k2
    } else {
      sc[j] = is * ga;
      sh[j] = k2[j] = 0.f;
    }
  }
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * rpp;
  for (int64_t r = (int64_t)blockIdx.x * rpp + rt; r < n; r += U * stride) {
    float v[U][8], g[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t ru = r + u * stride;
      if (ru < n) {
        if (MODE != 2) load8<T>(x + ru * c + cg * 8, v[u]);
        if (MODE != 0) load8<T>(dy + ru * c + cg * 8, g[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t ru = r + u * stride;
      if (ru >= n) break;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (MODE == 0) {
          o[j] = fmaf(v[u][j], sc[j], sh[j]);
          if (relu) o[j] = fmaxf(o[j], 0.f);
        } else if (MODE == 1) {
          const float gj = (relu && fmaf(v[u][j], sc[j], sh0[j]) <= 0.f) ? 0.f : g[u][j];
          o[j] = fmaf(gj, sc[j], -fmaf(v[u][j], k2[j], sh[j]));
        } else {
          o[j] = g[u][j] * sc[j];
        }
      }
      store8<T>(out + ru * c + cg * 8, o);
    }
  }
}

static int bn_blocks(int64_t n, int c, int64_t *rows_per_cta) {
  const int rpp = BN_THREADS / (c / 8);
  int64_t nb = 4LL * num_sms();
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
  return tsg_bn_stats2(x, dtype, n, c, eps, momentum, running_mean, running_var, nullptr, mean, invstd, ws, ws_bytes, stream);
}

/* ... and num_batches_tracked (int64 device scalar, may be NULL) incremented in the same launch */
int tsg_bn_stats2(const void *x, int dtype, int64_t n, int c, float eps, float momentum, float *running_mean, float *running_var,
                  int64_t *num_batches_tracked, float *mean, float *invstd, void *ws, size_t ws_bytes, tsg_stream_t stream) {
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
    bn_reduce_kernel<float, 0><<<nb, BN_THREADS, smem, stream>>>((const float *)x, nullptr, n, c, rows, nullptr, nullptr, nullptr, nullptr, 0, (float *)ws);
  else
    bn_reduce_kernel<__nv_bfloat16, 0><<<nb, BN_THREADS, smem, stream>>>((const __nv_bfloat16 *)x, nullptr, n, c, rows, nullptr,
                                                                        nullptr, nullptr, nullptr, 0, (float *)ws);
  bn_finalize_kernel<<<(c + 31) / 32, 1024, 0, stream>>>((const float *)ws, nb, c, eps, momentum, running_mean, running_var, mean,
                                                         invstd, (long long *)num_batches_tracked);
  return check_launch("tsg_bn_stats");
}

/* y = (x - mean) * invstd * gamma + beta, then max(y, 0) when relu != 0 (the BatchNorm + ReLU pair of every convolution block
 * in one pass); gamma / beta may be NULL (affine=False) */
int tsg_bn_apply(const void *x, int dtype, int64_t n, int c, const float *mean, const float *invstd, const float *gamma,
                 const float *beta, int relu, void *y, tsg_stream_t stream) {
  if (c <= 0 || c % 8 || n <= 0 || (dtype != TSG_F32 && dtype != TSG_BF16)) {
    set_error("tsg_bn_apply: need n > 0, c a multiple of 8, fp32 or bf16 rows");
    return TSG_ERR_UNSUPPORTED;
  }
  const int grid = (int)std::min<int64_t>((n + (BN_THREADS / (c / 8)) - 1) / (BN_THREADS / (c / 8)), 16LL * num_sms());
  if (dtype == TSG_F32)
    bn_apply_kernel<float, 0><<<grid, BN_THREADS, 0, stream>>>((const float *)x, nullptr, n, c, mean, invstd, gamma, beta, nullptr,
                                                             relu, (float *)y);
  else
    bn_apply_kernel<__nv_bfloat16, 0><<<grid, BN_THREADS, 0, stream>>>((const __nv_bfloat16 *)x, nullptr, n, c, mean, invstd, gamma,
                                                                     beta, nullptr, relu, (__nv_bfloat16 *)y);
  return check_launch("tsg_bn_apply");
}

/* Backward.  training != 0: sums (2 x C fp32) receives {dbeta = sum dy, dgamma = sum dy * xhat} and
 * dx = gamma invstd (dy - sums[0] / n - xhat sums[1] / n);  training == 0 (running statistics were used): the same sums,
 * dx = gamma invstd dy.  dx may be NULL (only the parameter gradients are wanted).  relu != 0 (training only): dy is the
 * gradient of relu(bn(x)); it is masked by bn(x) > 0, recomputed from x, in both passes. */
int tsg_bn_backward(const void *x, const void *dy, int dtype, int64_t n, int c, const float *mean, const float *invstd,
                    const float *gamma, const float *beta, int relu, int training, float *sums, void *dx, void *ws, size_t ws_bytes,
                    tsg_stream_t stream) {
  if (relu && !training) {
    set_error("tsg_bn_backward: the fused ReLU is supported in training mode only");
    return TSG_ERR_UNSUPPORTED;
  }
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
  const int grid = (int)std::min<int64_t>((n + (BN_THREADS / (c / 8)) - 1) / (BN_THREADS / (c / 8)), 16LL * num_sms());
  if (dtype == TSG_F32) {
    bn_reduce_kernel<float, 1><<<nb, BN_THREADS, smem, stream>>>((const float *)x, (const float *)dy, n, c, rows, mean, invstd, gamma, beta, relu, (float *)ws);
    bn_bwd_finalize_kernel<<<(c + 31) / 32, 1024, 0, stream>>>((const float *)ws, nb, c, sums);
    if (dx && training)
      bn_apply_kernel<float, 1><<<grid, BN_THREADS, 0, stream>>>((const float *)x, (const float *)dy, n, c, mean, invstd, gamma, beta,
                                                               sums, relu, (float *)dx);
    else if (dx)
      bn_apply_kernel<float, 2><<<grid, BN_THREADS, 0, stream>>>((const float *)x, (const float *)dy, n, c, mean, invstd, gamma, nullptr,
                                                               sums, 0, (float *)dx);
  } else {
    bn_reduce_kernel<__nv_bfloat16, 1><<<nb, BN_THREADS, smem, stream>>>((const __nv_bfloat16 *)x, (const __nv_bfloat16 *)dy, n, c, rows,
                                                                        mean, invstd, gamma, beta, relu, (float *)ws);
    bn_bwd_finalize_kernel<<<(c + 31) / 32, 1024, 0, stream>>>((const float *)ws, nb, c, sums);
    if (dx && training)
      bn_apply_kernel<__nv_bfloat16, 1><<<grid, BN_THREADS, 0, stream>>>((const __nv_bfloat16 *)x, (const __nv_bfloat16 *)dy, n, c, mean,
                                                                       invstd, gamma, beta, sums, relu, (__nv_bfloat16 *)dx);
    else if (dx)
      bn_apply_kernel<__nv_bfloat16, 2><<<grid, BN_THREADS, 0, stream>>>((const __nv_bfloat16 *)x, (const __nv_bfloat16 *)dy, n, c, mean,
                                                                       invstd, gamma, nullptr, sums, 0, (__nv_bfloat16 *)dx);
  }
  return check_launch("tsg_bn_backward");
}

}  // extern "C"
