// Point <-> voxel feature transforms of SPVCNN, B200-first (north_star subsystem 5).
//
// The reference kernels (TS/backend/voxelize/voxelize_cuda.cu:12-42, TS/backend/devoxelize/devoxelize_cuda.cu:11-57) give
// one scalar element to every thread: a 64-bit divide to find (row, channel), the row's index / weight words re-read by
// every one of its C threads, 4-byte accesses, and fp32 atomicAdd for the scatter-mean (order = whatever the scheduler
// does, so results change from run to run).  Here:
//   * a row of C channels is covered by a GROUP of lanes, 16 bytes (4 fp32 / 8 bf16 channels) per lane, a warp holds
//     32 / group rows; index and weight words are loaded once per row (same address across the group = one broadcast
//     transaction), rows move as full 16-byte vectors;
//   * the scatter-mean is a SEGMENTED REDUCTION: point ids are radix-sorted by voxel once per index tensor (a "plan",
//     reused by every voxelize over the same points), and one lane group walks a voxel's points in point order and writes
//     the mean once — no atomics, no zero-fill, no fp32 staging buffer for 16-bit features, bit-identical run to run;
//   * the trilinear gather keeps its 8 (index, weight) pairs in registers and accumulates in fp32;
//   * the backward scatter of devoxelize uses 16-byte vector atomics (red.global.add.v4.f32, sm_90+).
// Rows whose byte size is not a multiple of 16 fall back to the scalar kernels of points.cu.
#include "common.cuh"

namespace tsg {

template <typename T> struct Vec;   // 16 bytes of channels
template <> struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float *p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  static __device__ __forceinline__ void store(float *p, const float (&v)[8]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8]) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
    const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] = __uint_as_float(w[j] << 16);
      v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[8]) {
    unsigned w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      w[j] = *reinterpret_cast<const unsigned *>(&h);
    }
    *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct Vec<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __half *p, float (&v)[8]) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
    const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[j]));
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__half *p, const float (&v)[8]) {
    unsigned w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      w[j] = *reinterpret_cast<const unsigned *>(&h);
    }
    *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// ---------------------------------------------------------------- voxelize plan: points sorted by voxel
__global__ void vx_keys_kernel(const int *__restrict__ idx, int64_t n, int64_t m, unsigned long long *__restrict__ keys) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = __ldg(idx + i);
    keys[i] = (v >= 0 && v < m) ? (unsigned long long)v : (unsigned long long)m;   // unmatched points sort behind every voxel
  }
}
// seg[v] = {first sorted position, one past the last} of voxel v ({0, 0}: no point; the caller pre-fills with zeros)
__global__ void vx_heads_kernel(const unsigned long long *__restrict__ skeys, int64_t n, int64_t m, int2 *__restrict__ seg) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = skeys[j];
    if (k >= (unsigned long long)m) continue;
    if (j == 0 || skeys[j - 1] != k) seg[k].x = (int)j;
    if (j == n - 1 || skeys[j + 1] != k) seg[k].y = (int)j + 1;
  }
}

// out[v] = sum over the voxel's points, in point order, of feats[i] / counts[v]   (voxelize_cuda.cu:12-25 semantics)
template <typename T, int GS>
__global__ void __launch_bounds__(256) voxelize_seg_kernel(const T *__restrict__ feats, const unsigned *__restrict__ order,
                                                           const int2 *__restrict__ seg, const int *__restrict__ counts,
                                                           int64_t n, int c, int64_t m, T *__restrict__ out) {
  constexpr int V = Vec<T>::N;
  const int gl = threadIdx.x % GS;
  const int vpr = c / V;   // 16-byte vectors per row
  const int64_t groups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t v = blockIdx.x * (int64_t)(blockDim.x / GS) + threadIdx.x / GS; v < m; v += groups) {
    const int2 sg = __ldg(seg + v);   // the segment is known up front: the row loads below are independent of each other
    const int cnt = __ldg(counts + v);
    const float fc = (float)cnt;
    for (int col = gl; col < vpr; col += GS) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (cnt != 0) {
        int j = sg.x;
        for (; j + 1 < sg.y; j += 2) {   // two rows in flight; accumulation stays in point order
          const unsigned r0 = __ldg(order + j), r1 = __ldg(order + j + 1);
          float f0[8], f1[8];
          Vec<T>::load(feats + (int64_t)r0 * c + col * V, f0);
          Vec<T>::load(feats + (int64_t)r1 * c + col * V, f1);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] += __fdiv_rn(f0[e], fc);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] += __fdiv_rn(f1[e], fc);
        }
        if (j < sg.y) {
          float f[8];
          Vec<T>::load(feats + (int64_t)__ldg(order + j) * c + col * V, f);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] += __fdiv_rn(f[e], fc);
        }
      }
      Vec<T>::store(out + v * c + col * V, acc);
    }
  }
}

// bottom[i] = top[idx[i]] / counts[idx[i]]   (voxelize_cuda.cu:28-42)
template <typename T, int GS>
__global__ void __launch_bounds__(256) voxelize_bwd_vec_kernel(const T *__restrict__ top, const int *__restrict__ idx,
                                                               const int *__restrict__ counts, int64_t n, int c,
                                                               T *__restrict__ bottom) {
  constexpr int V = Vec<T>::N;
  const int gl = threadIdx.x % GS;
  const int vpr = c / V;
  const int64_t groups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t i = blockIdx.x * (int64_t)(blockDim.x / GS) + threadIdx.x / GS; i < n; i += groups) {
    const int v = __ldg(idx + i);
    const int cnt = v >= 0 ? __ldg(counts + v) : 0;
    for (int col = gl; col < vpr; col += GS) {
      float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (cnt != 0) {
        Vec<T>::load(top + (int64_t)v * c + col * V, g);
#pragma unroll
        for (int e = 0; e < V; ++e) g[e] = __fdiv_rn(g[e], (float)cnt);
      }
      Vec<T>::store(bottom + i * c + col * V, g);
    }
  }
}

// out[i] = sum_k w[i,k] * feats[idx[i,k]]   (devoxelize_cuda.cu:11-33): 8 (index, weight) pairs per point in registers
template <typename T, int GS>
__global__ void __launch_bounds__(256) devoxelize_vec_kernel(const T *__restrict__ feats, const int *__restrict__ idx8,
                                                             const float *__restrict__ w8, int64_t n, int c,
                                                             T *__restrict__ out) {
  constexpr int V = Vec<T>::N;
  const int gl = threadIdx.x % GS;
  const int vpr = c / V;
  const int64_t groups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t i = blockIdx.x * (int64_t)(blockDim.x / GS) + threadIdx.x / GS; i < n; i += groups) {
    const int4 ia = __ldg(reinterpret_cast<const int4 *>(idx8 + i * 8)), ib = __ldg(reinterpret_cast<const int4 *>(idx8 + i * 8) + 1);
    const float4 wa = __ldg(reinterpret_cast<const float4 *>(w8 + i * 8)), wb = __ldg(reinterpret_cast<const float4 *>(w8 + i * 8) + 1);
    const int id[8] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
    const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    for (int col = gl; col < vpr; col += GS) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (id[k] >= 0) {
          float f[8];
          Vec<T>::load(feats + (int64_t)id[k] * c + col * V, f);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] += w[k] * f[e];
        }
      }
      Vec<T>::store(out + i * c + col * V, acc);
    }
  }
}

// acc[idx[i,k]] += w[i,k] * top[i]   (devoxelize_cuda.cu:36-57): 16-byte vector reductions into the fp32 accumulator
template <typename T, int GS>
__global__ void __launch_bounds__(256) devoxelize_bwd_vec_kernel(const T *__restrict__ top, const int *__restrict__ idx8,
                                                                 const float *__restrict__ w8, int64_t n, int c,
                                                                 float *__restrict__ acc) {
  constexpr int V = Vec<T>::N;
  const int gl = threadIdx.x % GS;
  const int vpr = c / V;
  const int64_t groups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t i = blockIdx.x * (int64_t)(blockDim.x / GS) + threadIdx.x / GS; i < n; i += groups) {
    const int4 ia = __ldg(reinterpret_cast<const int4 *>(idx8 + i * 8)), ib = __ldg(reinterpret_cast<const int4 *>(idx8 + i * 8) + 1);
    const float4 wa = __ldg(reinterpret_cast<const float4 *>(w8 + i * 8)), wb = __ldg(reinterpret_cast<const float4 *>(w8 + i * 8) + 1);
    const int id[8] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
    const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    for (int col = gl; col < vpr; col += GS) {
      float g[8];
      Vec<T>::load(top + i * c + col * V, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (id[k] < 0) continue;
        float *dst = acc + (int64_t)id[k] * c + col * V;
#pragma unroll
        for (int h = 0; h < V / 4; ++h)
          atomicAdd(reinterpret_cast<float4 *>(dst) + h,
                    make_float4(w[k] * g[4 * h], w[k] * g[4 * h + 1], w[k] * g[4 * h + 2], w[k] * g[4 * h + 3]));
      }
    }
  }
}

template <typename T>
__global__ void cast_f32_vec_kernel(const float *__restrict__ in, int64_t n, T *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = from_f32<T>(in[i]);
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static int group_size(int vecs_per_row) {
  int g = 1;
  while (g < vecs_per_row && g < 32) g <<= 1;
  return g;
}
static bool vec_ok(const void *a, const void *b, int c, int elt) {
  return (c * elt) % 16 == 0 && ((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0;
}

}  // namespace tsg

using namespace tsg;

#define DISPATCH_T(dtype, ...)                                                 \
  switch (dtype) {                                                             \
    case TSG_F32: { using T = float; __VA_ARGS__; break; }                     \
    case TSG_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }            \
    case TSG_F16: { using T = __half; __VA_ARGS__; break; }                    \
    default: set_error("unknown dtype %d", dtype); return TSG_ERR_INVALID;     \
  }
#define DISPATCH_GS(gs, ...)                                                   \
  switch (gs) {                                                                \
    case 1: { constexpr int GS = 1; __VA_ARGS__; break; }                      \
    case 2: { constexpr int GS = 2; __VA_ARGS__; break; }                      \
    case 4: { constexpr int GS = 4; __VA_ARGS__; break; }                      \
    case 8: { constexpr int GS = 8; __VA_ARGS__; break; }                      \
    case 16: { constexpr int GS = 16; __VA_ARGS__; break; }                    \
    default: { constexpr int GS = 32; __VA_ARGS__; break; }                    \
  }

extern "C" {

int tsg_pv_vector_ok(const void *a, const void *b, int c, int dtype) {
  return vec_ok(a, b, c, dtype == TSG_F32 ? 4 : 2) ? 1 : 0;
}

size_t tsg_voxelize_plan_ws_bytes(int64_t n) {
  const int64_t k = n > 0 ? n : 1;
  return align256((size_t)k * 8) + tsg_sort_ws_bytes(k);
}

int tsg_voxelize_plan(const int32_t *idx, int64_t n, int64_t m, uint32_t *order, uint64_t *skeys, int32_t *seg,
                      void *ws, size_t ws_bytes, tsg_stream_t stream) {
  if (m <= 0) return TSG_OK;
  if (m >= (1ll << 31) || n >= (1ll << 30)) { set_error("tsg_voxelize_plan: too many rows"); return TSG_ERR_UNSUPPORTED; }
  if (ws_bytes < tsg_voxelize_plan_ws_bytes(n)) { set_error("tsg_voxelize_plan: workspace too small"); return TSG_ERR_WORKSPACE; }
  TSG_CUDA(cudaMemsetAsync(seg, 0, (size_t)m * 2 * sizeof(int), stream));
  if (n <= 0) return TSG_OK;
  unsigned long long *keys = (unsigned long long *)ws;
  char *sort_ws = (char *)ws + align256((size_t)n * 8);
  vx_keys_kernel<<<grid_for(n, 256), 256, 0, stream>>>(idx, n, m, keys);
  int bits = 1;
  while ((1ll << bits) <= m) ++bits;   // keys are 0..m
  const int rc = sort_pairs(keys, nullptr, n, 0, bits, (unsigned long long *)skeys, order, sort_ws,
                            ws_bytes - align256((size_t)n * 8), stream);
  if (rc != TSG_OK) return rc;
  vx_heads_kernel<<<grid_for(n, 256), 256, 0, stream>>>((const unsigned long long *)skeys, n, m, (int2 *)seg);
  return check_launch("tsg_voxelize_plan");
}

int tsg_voxelize_fwd_seg(const void *feats, int dtype, const uint32_t *order, const int32_t *seg, const int32_t *counts,
                         int64_t n, int c, int64_t m, void *out, tsg_stream_t stream) {
  if (m <= 0 || c <= 0) return TSG_OK;
  const int elt = dtype == TSG_F32 ? 4 : 2;
  if (!vec_ok(feats, out, c, elt)) {
    set_error("tsg_voxelize_fwd_seg: rows must be a multiple of 16 bytes and 16-byte aligned");
    return TSG_ERR_UNSUPPORTED;
  }
  const int gs = group_size(c * elt / 16);
  DISPATCH_T(dtype, DISPATCH_GS(gs, (voxelize_seg_kernel<T, GS><<<grid_for(m * gs, 256, 16), 256, 0, stream>>>(
      (const T *)feats, order, (const int2 *)seg, counts, n, c, m, (T *)out))));
  return check_launch("tsg_voxelize_fwd_seg");
}

int tsg_voxelize_bwd_vec(const void *top_grad, int dtype, const int32_t *idx, const int32_t *counts, int64_t n, int c,
                         void *bottom_grad, tsg_stream_t stream) {
  if (n <= 0 || c <= 0) return TSG_OK;
  const int elt = dtype == TSG_F32 ? 4 : 2;
  if (!vec_ok(top_grad, bottom_grad, c, elt)) { set_error("tsg_voxelize_bwd_vec: unaligned rows"); return TSG_ERR_UNSUPPORTED; }
  const int gs = group_size(c * elt / 16);
  DISPATCH_T(dtype, DISPATCH_GS(gs, (voxelize_bwd_vec_kernel<T, GS><<<grid_for(n * gs, 256, 16), 256, 0, stream>>>(
      (const T *)top_grad, idx, counts, n, c, (T *)bottom_grad))));
  return check_launch("tsg_voxelize_bwd_vec");
}

int tsg_devoxelize_fwd_vec(const void *feats, int dtype, const int32_t *idx8, const float *w8, int64_t n, int c, void *out,
                           tsg_stream_t stream) {
  if (n <= 0 || c <= 0) return TSG_OK;
  const int elt = dtype == TSG_F32 ? 4 : 2;
  if (!vec_ok(feats, out, c, elt) || ((uintptr_t)idx8 % 16) || ((uintptr_t)w8 % 16)) {
    set_error("tsg_devoxelize_fwd_vec: unaligned rows");
    return TSG_ERR_UNSUPPORTED;
  }
  const int gs = group_size(c * elt / 16);
  DISPATCH_T(dtype, DISPATCH_GS(gs, (devoxelize_vec_kernel<T, GS><<<grid_for(n * gs, 256, 16), 256, 0, stream>>>(
      (const T *)feats, idx8, w8, n, c, (T *)out))));
  return check_launch("tsg_devoxelize_fwd_vec");
}

int tsg_devoxelize_bwd_vec(const void *top_grad, int dtype, const int32_t *idx8, const float *w8, int64_t n, int c,
                           int64_t m, void *bottom_grad, float *acc_ws, tsg_stream_t stream) {
  if (m <= 0 || c <= 0) return TSG_OK;
  const int elt = dtype == TSG_F32 ? 4 : 2;
  float *acc = dtype == TSG_F32 ? (float *)bottom_grad : acc_ws;
  if (!acc) { set_error("tsg_devoxelize_bwd_vec: fp32 accumulation workspace required for 16-bit features"); return TSG_ERR_WORKSPACE; }
  if (!vec_ok(top_grad, acc, c, elt) || (c * 4) % 16 || ((uintptr_t)idx8 % 16) || ((uintptr_t)w8 % 16)) {
    set_error("tsg_devoxelize_bwd_vec: unaligned rows");
    return TSG_ERR_UNSUPPORTED;
  }
  TSG_CUDA(cudaMemsetAsync(acc, 0, (size_t)m * c * sizeof(float), stream));
  if (n > 0) {
    const int gs = group_size(c * elt / 16);
    DISPATCH_T(dtype, DISPATCH_GS(gs, (devoxelize_bwd_vec_kernel<T, GS><<<grid_for(n * gs, 256, 16), 256, 0, stream>>>(
        (const T *)top_grad, idx8, w8, n, c, acc))));
  }
  if (dtype != TSG_F32) {
    DISPATCH_T(dtype, (cast_f32_vec_kernel<T><<<grid_for(m * c, 256), 256, 0, stream>>>(acc, m * (int64_t)c, (T *)bottom_grad)));
  }
  return check_launch("tsg_devoxelize_bwd_vec");
}

}  // extern "C"
