// Multi-frame aggregation + quantisation front end and the point<->voxel transforms (a10-a14).
// HBM-bound: every kernel is a single coalesced pass (float4 / int4 vectors where rows are 16 B), fp32 arithmetic
// written with explicit round-to-nearest intrinsics wherever the reference's bit pattern must be reproduced.
#include <cstring>

#include "common.cuh"

namespace tsg {

// ---------------------------------------------------------------- pose warp
struct Pose2 {
  float p0[16], p[16];
};

// SemantickittiMsDataset.fuse_multi_scan (semantickitti_ms.py:403-417): every product rounded, sums left to right.
__device__ inline float3 warp_point(float x, float y, float z, const float *p0, const float *p) {
  float nw[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float acc = __fmul_rn(x, p[4 * j + 0]);
    acc = __fadd_rn(acc, __fmul_rn(y, p[4 * j + 1]));
    acc = __fadd_rn(acc, __fmul_rn(z, p[4 * j + 2]));
    acc = __fadd_rn(acc, p[4 * j + 3]);  // 1.0f * P[j][3] is exact
    nw[j] = __fsub_rn(acc, p0[4 * j + 3]);
  }
  float o[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float acc = __fmul_rn(nw[0], p0[0 * 4 + j]);
    acc = __fadd_rn(acc, __fmul_rn(nw[1], p0[1 * 4 + j]));
    acc = __fadd_rn(acc, __fmul_rn(nw[2], p0[2 * 4 + j]));
    o[j] = acc;
  }
  return make_float3(o[0], o[1], o[2]);
}

__global__ void fuse_kernel(const float *__restrict__ pts, int64_t n, int c, Pose2 ps, float *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float *r = pts + i * c;
    const float3 w = warp_point(r[0], r[1], r[2], ps.p0, ps.p);
    float *o = out + i * c;
    o[0] = w.x; o[1] = w.y; o[2] = w.z;
    for (int j = 3; j < c; ++j) o[j] = r[j];
  }
}

struct RT {
  double R[9], T[3];
};
// nuscenes_ms.py:371: float32 points promoted to float64, row-vector times R plus T, stored back as float32
__global__ void transform_point_kernel(const float *__restrict__ pts, int64_t n, int c, RT rt, float *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float *r = pts + i * c;
    const double x = r[0], y = r[1], z = r[2];
    float *o = out + i * c;
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o[j] = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, rt.R[j]), __dmul_rn(y, rt.R[3 + j])),
                                        __dmul_rn(z, rt.R[6 + j])), rt.T[j]);
    for (int j = 3; j < c; ++j) o[j] = r[j];
  }
}

// ---------------------------------------------------------------- aggregation front end
__device__ inline void atomic_min_float(float *addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned *>(addr), __float_as_uint(v));
}

struct AggWs {  // per sample
  float cur_min[4];
  int ms_min[4];
  int ms_max[4];
};

__global__ void agg_init_kernel(AggWs *ws, int n_samples) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_samples) {
    for (int j = 0; j < 4; ++j) {
      ws[i].cur_min[j] = __int_as_float(0x7f800000);
      ws[i].ms_min[j] = 0x7fffffff;
      ws[i].ms_max[j] = -0x7fffffff;
    }
  }
}

// pass 1 (grid.y = frame): warp history frames, append the time flag, min corner of the current scan
__global__ void agg_warp_kernel(const float *__restrict__ pts, int c_in, const tsg_frame *__restrict__ frames,
                                float *__restrict__ feats, AggWs *ws) {
  const tsg_frame f = frames[blockIdx.y];
  const int c_out = c_in + 1;
  float mn[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < f.count; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = f.offset + t;
    const float *r = pts + i * c_in;
    float3 w = make_float3(r[0], r[1], r[2]);
    if (!f.is_cur) w = warp_point(w.x, w.y, w.z, f.pose0, f.pose);
    float *o = feats + i * c_out;
    o[0] = w.x; o[1] = w.y; o[2] = w.z;
    if (c_in > 3) o[3] = r[3];
    o[c_in > 3 ? 4 : 3] = f.is_cur ? 1.f : 0.f;  // append_time_flag (semantickitti_ms.py:253-257)
    for (int j = 4; j < c_in; ++j) o[j + 1] = r[j];
    if (f.is_cur) {
      mn[0] = fminf(mn[0], w.x); mn[1] = fminf(mn[1], w.y); mn[2] = fminf(mn[2], w.z);
    }
  }
  if (f.is_cur) {  // one atomic per block and axis: per-warp atomics on the same few words serialise in the L2
    __shared__ float s_mn[8][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float v = mn[j];
      for (int s = 16; s; s >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, s));
      if ((threadIdx.x & 31) == 0) s_mn[threadIdx.x >> 5][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      float v = s_mn[0][threadIdx.x];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = fminf(v, s_mn[w][threadIdx.x]);
      if (v < __int_as_float(0x7f800000)) atomic_min_float(&ws[f.sample].cur_min[threadIdx.x], v);
    }
  }
}

// nuScenes pass 1 (grid.y = sweep; nuscenes_ms.py:284-341, :348-373): per point of sweep k the ego-box test on the RAW
// coordinates (|x| < 1 & |y| < 1.5 -> dropped), the sweep's time lag into column 4, the float64 warp into the key
// frame (p @ R + T: products and sums rounded one by one, like numpy's matmul of a row by a 3x3 matrix), the min corner
// of the key sweep's SURVIVING points.  flags = "outside the ego box"; pass 2 (agg_quant_kernel) ANDs the clamp.
__global__ void agg_warp_nus_kernel(const float *__restrict__ pts, int c, const tsg_sweep *__restrict__ sweeps,
                                    float *__restrict__ feats, uint8_t *__restrict__ flags, AggWs *ws) {
  const tsg_sweep f = sweeps[blockIdx.y];
  float mn[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < f.count; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = f.offset + t;
    const float *r = pts + i * c;
    const float x0 = r[0], y0 = r[1], z0 = r[2];
    const bool keep = !(fabsf(x0) < 1.0f && fabsf(y0) < 1.5f);
    float w[3] = {x0, y0, z0};
    if (!f.is_key) {
      const double x = x0, y = y0, z = z0;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        w[j] = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, f.R[j]), __dmul_rn(y, f.R[3 + j])), __dmul_rn(z, f.R[6 + j])), f.T[j]);
    }
    float *o = feats + i * c;
    o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
    for (int j = 3; j < c; ++j) o[j] = j == 4 ? f.dt : r[j];
    flags[i] = keep ? 1 : 0;
    if (f.is_key && keep) {
      mn[0] = fminf(mn[0], w[0]); mn[1] = fminf(mn[1], w[1]); mn[2] = fminf(mn[2], w[2]);
    }
  }
  if (f.is_key) {
    __shared__ float s_mn[8][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float v = mn[j];
      for (int s = 16; s; s >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, s));
      if ((threadIdx.x & 31) == 0) s_mn[threadIdx.x >> 5][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      float v = s_mn[0][threadIdx.x];
      for (int w2 = 1; w2 < (int)(blockDim.x >> 5); ++w2) v = fminf(v, s_mn[w2][threadIdx.x]);
      if (v < __int_as_float(0x7f800000)) atomic_min_float(&ws[f.sample].cur_min[threadIdx.x], v);
    }
  }
}

// pass 2: keep mask (FSA & clamp to the current scan's min corner), quantise, per-sample min of kept voxels
__global__ void agg_quant_kernel(const float *__restrict__ feats, int c_out, const tsg_frame *__restrict__ frames,
                                 const uint8_t *keep, float voxel, int4 *__restrict__ coords, uint8_t *flags, AggWs *ws) {
  // keep / flags may be the same array (nuScenes: the ego-box flags of pass 1 are refined in place)
  const tsg_frame f = frames[blockIdx.y];
  const float cx = ws[f.sample].cur_min[0], cy = ws[f.sample].cur_min[1], cz = ws[f.sample].cur_min[2];
  int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
  int mx[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < f.count; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = f.offset + t;
    const float *r = feats + i * c_out;
    const float x = r[0], y = r[1], z = r[2];
    bool k = f.is_cur || !keep || keep[i];
    k = k && (x >= cx) && (y >= cy) && (z >= cz);  // semantickitti_voxel_ms.py:121
    // np.round(xyz / voxel).astype(int32): IEEE fp32 divide, round half to even (semantickitti_voxel_ms.py:127-128)
    const int qx = __float2int_rn(__fdiv_rn(x, voxel)), qy = __float2int_rn(__fdiv_rn(y, voxel)),
              qz = __float2int_rn(__fdiv_rn(z, voxel));
    coords[i] = make_int4(qx, qy, qz, f.sample);
    flags[i] = k ? 1 : 0;
    if (k) {
      mn[0] = min(mn[0], qx); mn[1] = min(mn[1], qy); mn[2] = min(mn[2], qz);
      mx[0] = max(mx[0], qx); mx[1] = max(mx[1], qy); mx[2] = max(mx[2], qz);
    }
  }
  __shared__ int s_mn[8][3], s_mx[8][3];  // one atomic pair per block and axis (see agg_warp_kernel)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    int v = mn[j], u = mx[j];
    for (int s = 16; s; s >>= 1) {
      v = min(v, __shfl_xor_sync(0xffffffffu, v, s));
      u = max(u, __shfl_xor_sync(0xffffffffu, u, s));
    }
    if ((threadIdx.x & 31) == 0) {
      s_mn[threadIdx.x >> 5][j] = v;
      s_mx[threadIdx.x >> 5][j] = u;
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    int v = s_mn[0][threadIdx.x], u = s_mx[0][threadIdx.x];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      v = min(v, s_mn[w][threadIdx.x]);
      u = max(u, s_mx[w][threadIdx.x]);
    }
    if (v != 0x7fffffff) {
      atomicMin(&ws[f.sample].ms_min[threadIdx.x], v);
      atomicMax(&ws[f.sample].ms_max[threadIdx.x], u);
    }
  }
}

// pass 3: pc -= pc.min(0) per sample (semantickitti_voxel_ms.py:151)
__global__ void agg_shift_kernel(const tsg_frame *__restrict__ frames, int4 *__restrict__ coords, const AggWs *ws) {
  const tsg_frame f = frames[blockIdx.y];
  const int mx = ws[f.sample].ms_min[0], my = ws[f.sample].ms_min[1], mz = ws[f.sample].ms_min[2];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < f.count; t += (int64_t)gridDim.x * blockDim.x) {
    int4 c = coords[f.offset + t];
    c.x -= mx; c.y -= my; c.z -= mz;
    coords[f.offset + t] = c;
  }
}

// ---------------------------------------------------------------- stable compaction
constexpr int CP_ROWS = 1024;
__global__ void __launch_bounds__(256) cp_count_kernel(const uint8_t *__restrict__ flags, int64_t n, int *blocksum) {
  const int64_t base = (int64_t)blockIdx.x * CP_ROWS + threadIdx.x * 4;
  int c = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (base + r < n) c += flags[base + r] ? 1 : 0;
  int tot;
  block_exclusive_scan<256>(c, &tot);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) cp_scan_kernel(int *data, int64_t n, int *total_out) {
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b = 0; b < n; b += 1024) {
    const int64_t i = b + threadIdx.x;
    const int v = i < n ? data[i] : 0;
    int tot;
    const int ex = block_exclusive_scan<1024>(v, &tot);
    if (i < n) data[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}
__global__ void __launch_bounds__(256) cp_write_kernel(const uint8_t *__restrict__ flags, int64_t n,
                                                       const int *__restrict__ blockoff, const uint32_t *rows_a, int wa,
                                                       uint32_t *out_a, const uint32_t *rows_b, int wb, uint32_t *out_b,
                                                       int *pos_out) {
  __shared__ int s_pos[CP_ROWS];  // destination row of each of the block's rows, or -1
  const int64_t row0 = (int64_t)blockIdx.x * CP_ROWS;
  const int64_t base = row0 + threadIdx.x * 4;
  bool k[4];
  int c = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    k[r] = base + r < n && flags[base + r];
    c += k[r] ? 1 : 0;
  }
  int pos = block_exclusive_scan<256>(c, nullptr) + blockoff[blockIdx.x];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    s_pos[threadIdx.x * 4 + r] = k[r] ? pos : -1;
    if (base + r < n && pos_out) pos_out[base + r] = k[r] ? pos : -1;
    pos += k[r] ? 1 : 0;
  }
  __syncthreads();
  // the block's rows are one contiguous span of words: consecutive threads move consecutive words, and since kept
  // rows stay in order the writes are (piecewise) contiguous as well
  const int rows_here = (int)(n - row0 < CP_ROWS ? n - row0 : CP_ROWS);
  if (rows_a)
    for (int w = threadIdx.x; w < rows_here * wa; w += 256) {
      const int r = w / wa, d = s_pos[r];
      if (d >= 0) out_a[(int64_t)d * wa + (w - r * wa)] = rows_a[row0 * wa + w];
    }
  if (rows_b)
    for (int w = threadIdx.x; w < rows_here * wb; w += 256) {
      const int r = w / wb, d = s_pos[r];
      if (d >= 0) out_b[(int64_t)d * wb + (w - r * wb)] = rows_b[row0 * wb + w];
    }
}

__global__ void gather_rows_kernel(const uint32_t *__restrict__ src, int width, const int *__restrict__ idx, int64_t n,
                                   const int *__restrict__ n_dev, uint32_t *__restrict__ out) {
  n = dev_count(n_dev, n);
  const int64_t total = n * width;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / width;
    const int j = (int)(t - i * width);
    const int s = idx[i];
    out[t] = s >= 0 ? src[(int64_t)s * width + j] : 0u;
  }
}

// ---------------------------------------------------------------- count / voxelize / devoxelize
__global__ void count_kernel(const int *__restrict__ idx, int64_t n, int *__restrict__ out, int64_t m) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = idx[i];
    if (v >= 0 && v < m) atomicAdd(&out[v], 1);
  }
}

// scatter-mean (voxelize_cuda.cu:12-25): acc[idx[i]] += feat[i] / count[idx[i]]
template <typename T>
__global__ void voxelize_fwd_kernel(const T *__restrict__ feats, const int *__restrict__ idx,
                                    const int *__restrict__ counts, int64_t n, int c, float *__restrict__ acc) {
  const int64_t total = n * c;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / c;
    const int j = (int)(t - i * c);
    const int v = idx[i];
    if (v < 0) continue;
    const int cnt = counts[v];
    if (cnt == 0) continue;
    atomicAdd(&acc[(int64_t)v * c + j], to_f32(feats[t]) / (float)cnt);
  }
}
template <typename T>
__global__ void cast_from_f32_kernel(const float *__restrict__ in, int64_t n, T *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = from_f32<T>(in[i]);
}
template <typename T>
__global__ void voxelize_bwd_kernel(const T *__restrict__ top, const int *__restrict__ idx,
                                    const int *__restrict__ counts, int64_t n, int c, T *__restrict__ bottom) {
  const int64_t total = n * c;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / c;
    const int j = (int)(t - i * c);
    const int v = idx[i];
    float g = 0.f;
    if (v >= 0 && counts[v] != 0) g = to_f32(top[(int64_t)v * c + j]) / (float)counts[v];
    bottom[t] = from_f32<T>(g);
  }
}

// trilinear gather (devoxelize_cuda.cu:11-33): out[i] = sum_k w[i,k] * feat[idx[i,k]]
template <typename T>
__global__ void devoxelize_fwd_kernel(const T *__restrict__ feats, const int *__restrict__ idx8,
                                      const float *__restrict__ w8, int64_t n, int c, T *__restrict__ out) {
  const int64_t total = n * c;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / c;
    const int j = (int)(t - i * c);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int v = __ldg(idx8 + i * 8 + k);
      if (v >= 0) acc += __ldg(w8 + i * 8 + k) * to_f32(feats[(int64_t)v * c + j]);
    }
    out[t] = from_f32<T>(acc);
  }
}
template <typename T>
__global__ void devoxelize_bwd_kernel(const T *__restrict__ top, const int *__restrict__ idx8,
                                      const float *__restrict__ w8, int64_t n, int c, float *__restrict__ acc) {
  const int64_t total = n * c;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / c;
    const int j = (int)(t - i * c);
    const float g = to_f32(top[t]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int v = __ldg(idx8 + i * 8 + k);
      if (v >= 0) atomicAdd(&acc[(int64_t)v * c + j], __ldg(w8 + i * 8 + k) * g);
    }
  }
}

// ---------------------------------------------------------------- fused point -> voxel queries
__device__ inline int floor_to_stride(float v, int s) {
  // torch.floor(C / s).int() * s  (minkunet/utils.py:47, 74): fp32 divide, floor, int multiply
  return (int)floorf(__fdiv_rn(v, (float)s)) * s;
}

__global__ void point_query_kernel(const Slot *__restrict__ tab, unsigned long long mask,
                                   const float4 *__restrict__ pc, int64_t n, int s, int *__restrict__ idx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = __ldg(pc + i);
    const int x = floor_to_stride(p.x, s), y = floor_to_stride(p.y, s), z = floor_to_stride(p.z, s), b = (int)p.w;
    idx[i] = coord_in_range(x, y, z, b) ? table_find_coord(tab, mask, pack_coord(x, y, z, b)) : -1;
  }
}

// voxel_to_point's map: 8 corners of the stride-s cell (get_kernel_offsets(2,s): z fastest) + calc_ti_weights
// (TS/nn/functional/devoxelize.py:10-48) evaluated in the same fp32 operation order, all in registers.
__global__ void trilinear_query_kernel(const Slot *__restrict__ tab, unsigned long long mask,
                                       const float4 *__restrict__ pc, int64_t n, int s, int nearest,
                                       int *__restrict__ idx8, float *__restrict__ w8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = __ldg(pc + i);
    const int bx = floor_to_stride(p.x, s), by = floor_to_stride(p.y, s), bz = floor_to_stride(p.z, s), b = (int)p.w;
    const float fs = (float)s;
    float pf[3], pcn[3];
    const float pv[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      pf[j] = (s != 1) ? __fmul_rn(floorf(__fdiv_rn(pv[j], fs)), fs) : floorf(pv[j]);
      pcn[j] = __fadd_rn(pf[j], fs);
    }
    const float xl = __fsub_rn(pcn[0], pv[0]), xh = __fsub_rn(pv[0], pf[0]);
    const float yl = __fsub_rn(pcn[1], pv[1]), yh = __fsub_rn(pv[1], pf[1]);
    const float zl = __fsub_rn(pcn[2], pv[2]), zh = __fsub_rn(pv[2], pf[2]);
    int id[8];
    float w[8];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
      const int x = bx + ix * s, y = by + iy * s, z = bz + iz * s;
      id[k] = coord_in_range(x, y, z, b) ? table_find_coord(tab, mask, pack_coord(x, y, z, b)) : -1;
      float wk = __fmul_rn(__fmul_rn(ix ? xh : xl, iy ? yh : yl), iz ? zh : zl);
      if (s != 1) wk = __fdiv_rn(wk, (float)(s * s * s));
      if (id[k] < 0) wk = 0.f;
      w[k] = wk;
      sum = __fadd_rn(sum, wk);
    }
    const float den = __fadd_rn(sum, 1e-8f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float wk = __fdiv_rn(w[k], den);
      int v = id[k];
      if (nearest && k > 0) { wk = 0.f; v = -1; }
      idx8[i * 8 + k] = v;
      w8[i * 8 + k] = wk;
    }
  }
}

// Fused multi-scale devoxelisation (the tail of MinkUNet / MinkUNetMs once the classifier has been applied per scale at
// voxel level): out[j,:] = sum_s sum_k w_{s,k}(p) * feats_s[idx_{s,k}(p), :], p = pcoords[rows[j]].  The 8-corner query,
// calc_ti_weights and the gather of every scale happen in registers: no (N,8) index / weight tensors, no per-scale
// outputs, no adds, no final row gather.  8 lanes per point: lane j probes corner j, then owns channels 4j..4j+3.
struct DevoxScales {
  const Slot *tab[4];
  unsigned long long mask[4];
  const float *feats[4];
  int stride[4];
  int count;
};
__global__ void __launch_bounds__(256) devoxelize_multi_kernel(DevoxScales sc, const float4 *__restrict__ pc,
                                                               const int *__restrict__ rows, int64_t m, int c,
                                                               float *__restrict__ out, int c_out) {
  const int sub = threadIdx.x & 7;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);  // the 8 lanes of this point
  const int gl0 = threadIdx.x & 24;
  const int64_t step = (int64_t)gridDim.x * (blockDim.x >> 3);
  const int64_t m_pad = (m + 3) & ~(int64_t)3;         // whole warps stay converged for the shuffles
  for (int64_t j = blockIdx.x * (int64_t)(blockDim.x >> 3) + (threadIdx.x >> 3); j < m_pad; j += step) {
    const bool live = j < m;
    const int64_t pi = live ? (rows ? (int64_t)__ldg(rows + j) : j) : 0;
    const float4 p = __ldg(pc + pi);
    const float pv[3] = {p.x, p.y, p.z};
    const int b = (int)p.w;
    const int ix = sub >> 2, iy = (sub >> 1) & 1, iz = sub & 1;
    // (1) every lane probes ITS corner of every scale and the 8 lanes normalise the weights (calc_ti_weights): done once
    // per point, outside the channel loop, so all 8 lanes take part in the shuffles whatever the channel count
    int ids[4];
    float wts[4];
#pragma unroll
    for (int si = 0; si < 4; ++si) {
      ids[si] = -1;
      wts[si] = 0.f;
      if (si >= sc.count) continue;
      const int s = sc.stride[si];
      const float fs = (float)s;
      float pf[3], lo[3], hi[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        pf[a] = (s != 1) ? __fmul_rn(floorf(__fdiv_rn(pv[a], fs)), fs) : floorf(pv[a]);
        lo[a] = __fsub_rn(__fadd_rn(pf[a], fs), pv[a]);
        hi[a] = __fsub_rn(pv[a], pf[a]);
      }
      float wk = __fmul_rn(__fmul_rn(ix ? hi[0] : lo[0], iy ? hi[1] : lo[1]), iz ? hi[2] : lo[2]);
      if (s != 1) wk = __fdiv_rn(wk, (float)(s * s * s));
      int id = -1;
      if (wk != 0.f && live) {   // a zero-weight corner contributes nothing whether it exists or not
        const int x = floor_to_stride(pv[0], s) + ix * s, y = floor_to_stride(pv[1], s) + iy * s,
                  z = floor_to_stride(pv[2], s) + iz * s;
        if (coord_in_range(x, y, z, b)) id = table_find_coord(sc.tab[si], sc.mask[si], pack_coord(x, y, z, b));
      }
      if (id < 0) wk = 0.f;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sum = __fadd_rn(sum, __shfl_sync(gmask, wk, gl0 + k));
      ids[si] = id;
      wts[si] = __fdiv_rn(wk, __fadd_rn(sum, 1e-8f));
    }
    // (2) channel passes: lane `sub` owns channels ch0 + 4 sub .. + 3 of every pass (one pass for c <= 32); lanes whose
    // channels lie beyond c still run the shuffles and skip only the loads / stores
    for (int ch0 = 0; ch0 < c; ch0 += 32) {
      const int ch = ch0 + 4 * sub;
      const bool mine = ch < c;
      float4 total = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int si = 0; si < 4; ++si) {
        if (si >= sc.count) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *f = sc.feats[si];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int v = __shfl_sync(gmask, ids[si], gl0 + k);
          const float w = __shfl_sync(gmask, wts[si], gl0 + k);
          if (v >= 0 && mine) {
            const float4 r = __ldg(reinterpret_cast<const float4 *>(f + (int64_t)v * c + ch));
            acc.x += w * r.x; acc.y += w * r.y; acc.z += w * r.z; acc.w += w * r.w;
          }
        }
        total.x += acc.x; total.y += acc.y; total.z += acc.z; total.w += acc.w;
      }
      if (live && mine) {
        float *o = out + j * c_out + ch;
        if (ch + 0 < c_out) o[0] = total.x;
        if (ch + 1 < c_out) o[1] = total.y;
        if (ch + 2 < c_out) o[2] = total.z;
        if (ch + 3 < c_out) o[3] = total.w;
      }
    }
  }
}

__global__ void rescale_coords_kernel(const float4 *__restrict__ pc, int64_t n, float init_res, float after_res,
                                      float4 *__restrict__ out_f, int4 *__restrict__ out_i) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = __ldg(pc + i);
    float4 f;
    f.x = __fdiv_rn(__fmul_rn(p.x, init_res), after_res);
    f.y = __fdiv_rn(__fmul_rn(p.y, init_res), after_res);
    f.z = __fdiv_rn(__fmul_rn(p.z, init_res), after_res);
    f.w = p.w;
    if (out_f) out_f[i] = f;
    if (out_i) out_i[i] = make_int4((int)floorf(f.x), (int)floorf(f.y), (int)floorf(f.z), (int)floorf(f.w));
  }
}

__global__ void cast_pad_bf16_kernel(const float *__restrict__ in, int64_t n, const int *__restrict__ n_dev, int c,
                                     int c_pad, __nv_bfloat16 *__restrict__ out) {
  n = dev_count(n_dev, n);
  const int64_t total = n * c_pad;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / c_pad;
    const int j = (int)(t - i * c_pad);
    out[t] = __float2bfloat16_rn(j < c ? in[i * c + j] : 0.f);
  }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static bool pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace tsg

using namespace tsg;

#define DISPATCH_DTYPE(dtype, ...)                                             \
  switch (dtype) {                                                             \
    case TSG_F32: { using T = float; __VA_ARGS__; break; }                     \
    case TSG_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }            \
    case TSG_F16: { using T = __half; __VA_ARGS__; break; }                    \
    default: set_error("unknown dtype %d", dtype); return TSG_ERR_INVALID;     \
  }

extern "C" {

int tsg_fuse_multi_scan(const float *pts, int64_t n, int c, const float *pose0, const float *pose, float *out,
                        tsg_stream_t stream) {
  if (c < 3) { set_error("tsg_fuse_multi_scan: need at least 3 columns"); return TSG_ERR_INVALID; }
  if (n <= 0) return TSG_OK;
  Pose2 ps;
  for (int i = 0; i < 16; ++i) { ps.p0[i] = pose0[i]; ps.p[i] = pose[i]; }
  fuse_kernel<<<grid_for(n, 256), 256, 0, stream>>>(pts, n, c, ps, out);
  return check_launch("tsg_fuse_multi_scan");
}

int tsg_transform_point(const float *pts, int64_t n, int c, const double *R, const double *T, float *out,
                        tsg_stream_t stream) {
  if (c < 3) { set_error("tsg_transform_point: need at least 3 columns"); return TSG_ERR_INVALID; }
  if (n <= 0) return TSG_OK;
  RT rt;
  for (int i = 0; i < 9; ++i) rt.R[i] = R[i];
  for (int i = 0; i < 3; ++i) rt.T[i] = T[i];
  transform_point_kernel<<<grid_for(n, 256), 256, 0, stream>>>(pts, n, c, rt, out);
  return check_launch("tsg_transform_point");
}

size_t tsg_aggregate_ws_bytes(int n_samples) {
  return align256(sizeof(AggWs) * (size_t)(n_samples > 0 ? n_samples : 1)) + align256(sizeof(tsg_frame) * 1024);
}

int tsg_aggregate_quantize(const float *pts, int c_in, const tsg_frame *frames_host, int n_frames, int n_samples,
                           const uint8_t *keep, float voxel_size, float *feats, int32_t *coords, uint8_t *flags,
                           void *ws, size_t ws_bytes, tsg_stream_t stream) {
  if (n_frames <= 0) return TSG_OK;
  if (n_frames > 1024 || c_in < 3 || n_samples <= 0 || n_samples > 128) {
    set_error("tsg_aggregate_quantize: need 1<=frames<=1024, c_in>=3, 1<=samples<=128");
    return TSG_ERR_INVALID;
  }
  if (ws_bytes < tsg_aggregate_ws_bytes(n_samples)) { set_error("tsg_aggregate_quantize: workspace too small"); return TSG_ERR_WORKSPACE; }
  AggWs *aw = (AggWs *)ws;
  tsg_frame *fr = (tsg_frame *)((char *)ws + align256(sizeof(AggWs) * (size_t)n_samples));
  int64_t maxcount = 1;
  for (int i = 0; i < n_frames; ++i) {
    if (frames_host[i].sample < 0 || frames_host[i].sample >= n_samples) { set_error("tsg_aggregate_quantize: bad sample id"); return TSG_ERR_INVALID; }
    if (frames_host[i].count > maxcount) maxcount = frames_host[i].count;
  }
  TSG_CUDA(cudaMemcpyAsync(fr, frames_host, sizeof(tsg_frame) * n_frames, cudaMemcpyHostToDevice, stream));
  agg_init_kernel<<<1, 128, 0, stream>>>(aw, n_samples);
  dim3 grid(grid_for(maxcount, 256, 2), n_frames);
  agg_warp_kernel<<<grid, 256, 0, stream>>>(pts, c_in, fr, feats, aw);
  agg_quant_kernel<<<grid, 256, 0, stream>>>(feats, c_in + 1, fr, keep, voxel_size, (int4 *)coords, flags, aw);
  agg_shift_kernel<<<grid, 256, 0, stream>>>(fr, (int4 *)coords, aw);
  return check_launch("tsg_aggregate_quantize");
}

/* Same passes with the frame table already on the device (graph-capturable: no host-to-device copy inside): frames_dev
 * holds n_frames tsg_frame records, max_count >= every frame's count (sizes the grid). */
int tsg_aggregate_quantize_dev(const float *pts, int c_in, const tsg_frame *frames_dev, int n_frames, int64_t max_count,
                               int n_samples, const uint8_t *keep, float voxel_size, float *feats, int32_t *coords,
                               uint8_t *flags, void *ws, size_t ws_bytes, tsg_stream_t stream) {
  if (n_frames <= 0) return TSG_OK;
  if (n_frames > 1024 || c_in < 3 || n_samples <= 0 || n_samples > 128 || max_count <= 0) {
    set_error("tsg_aggregate_quantize_dev: need 1<=frames<=1024, c_in>=3, 1<=samples<=128, max_count>0");
    return TSG_ERR_INVALID;
  }
  if (ws_bytes < align256(sizeof(AggWs) * (size_t)n_samples)) { set_error("tsg_aggregate_quantize_dev: workspace too small"); return TSG_ERR_WORKSPACE; }
  AggWs *aw = (AggWs *)ws;
  agg_init_kernel<<<1, 128, 0, stream>>>(aw, n_samples);
  dim3 grid(grid_for(max_count, 256, 2), n_frames);
  agg_warp_kernel<<<grid, 256, 0, stream>>>(pts, c_in, frames_dev, feats, aw);
  agg_quant_kernel<<<grid, 256, 0, stream>>>(feats, c_in + 1, frames_dev, keep, voxel_size, (int4 *)coords, flags, aw);
  agg_shift_kernel<<<grid, 256, 0, stream>>>(frames_dev, (int4 *)coords, aw);
  return check_launch("tsg_aggregate_quantize_dev");
}

size_t tsg_aggregate_nus_ws_bytes(int n_samples) {
  return align256(sizeof(AggWs) * (size_t)(n_samples > 0 ? n_samples : 1)) + align256(sizeof(tsg_frame) * 1024) +
         align256(sizeof(tsg_sweep) * 1024);
}

int tsg_aggregate_quantize_nus(const float *pts, int c, const tsg_sweep *sweeps_host, int n_sweeps, int n_samples,
                               float voxel_size, float *feats, int32_t *coords, uint8_t *flags, void *ws, size_t ws_bytes,
                               tsg_stream_t stream) {
  if (n_sweeps <= 0) return TSG_OK;
  if (n_sweeps > 1024 || c < 5 || n_samples <= 0 || n_samples > 128) {
    set_error("tsg_aggregate_quantize_nus: need 1<=sweeps<=1024, c>=5 (x,y,z,intensity,dt), 1<=samples<=128");
    return TSG_ERR_INVALID;
  }
  if (ws_bytes < tsg_aggregate_nus_ws_bytes(n_samples)) { set_error("tsg_aggregate_quantize_nus: workspace too small"); return TSG_ERR_WORKSPACE; }
  AggWs *aw = (AggWs *)ws;
  tsg_frame *fr = (tsg_frame *)((char *)ws + align256(sizeof(AggWs) * (size_t)n_samples));
  tsg_sweep *sw = (tsg_sweep *)((char *)fr + align256(sizeof(tsg_frame) * 1024));
  static thread_local tsg_frame frames_host[1024];   // (offset, count, sample) view of the sweeps for the shared passes 2 and 3
  int64_t maxcount = 1;
  for (int i = 0; i < n_sweeps; ++i) {
    if (sweeps_host[i].sample < 0 || sweeps_host[i].sample >= n_samples) { set_error("tsg_aggregate_quantize_nus: bad sample id"); return TSG_ERR_INVALID; }
    if (sweeps_host[i].count > maxcount) maxcount = sweeps_host[i].count;
    memset(&frames_host[i], 0, sizeof(tsg_frame));
    frames_host[i].offset = sweeps_host[i].offset;
    frames_host[i].count = sweeps_host[i].count;
    frames_host[i].sample = sweeps_host[i].sample;
    frames_host[i].is_cur = 0;   // the ego-box mask applies to the key sweep too
  }
  TSG_CUDA(cudaMemcpyAsync(sw, sweeps_host, sizeof(tsg_sweep) * n_sweeps, cudaMemcpyHostToDevice, stream));
  TSG_CUDA(cudaMemcpyAsync(fr, frames_host, sizeof(tsg_frame) * n_sweeps, cudaMemcpyHostToDevice, stream));
  // (pageable sources: cudaMemcpyAsync returns once the data sits in the driver's staging buffer, so frames_host may be reused)
  agg_init_kernel<<<1, 128, 0, stream>>>(aw, n_samples);
  dim3 grid(grid_for(maxcount, 256, 2), n_sweeps);
  agg_warp_nus_kernel<<<grid, 256, 0, stream>>>(pts, c, sw, feats, flags, aw);
  agg_quant_kernel<<<grid, 256, 0, stream>>>(feats, c, fr, flags, voxel_size, (int4 *)coords, flags, aw);
  agg_shift_kernel<<<grid, 256, 0, stream>>>(fr, (int4 *)coords, aw);
  return check_launch("tsg_aggregate_quantize_nus");
}

size_t tsg_compact_ws_bytes(int64_t n) { return align256((size_t)((n > 0 ? n : 1) / CP_ROWS + 2) * 4); }

int tsg_compact_rows(const uint8_t *flags, int64_t n, const void *rows_a, int wa, void *out_a, const void *rows_b,
                     int wb, void *out_b, int32_t *pos, int32_t *m_dev, void *ws, size_t ws_bytes,
                     tsg_stream_t stream) {
  if (n <= 0) {
    if (m_dev) TSG_CUDA(cudaMemsetAsync(m_dev, 0, sizeof(int), stream));
    return TSG_OK;
  }
  if (ws_bytes < tsg_compact_ws_bytes(n)) { set_error("tsg_compact_rows: workspace too small"); return TSG_ERR_WORKSPACE; }
  const int64_t nblk = (n + CP_ROWS - 1) / CP_ROWS;
  int *bs = (int *)ws;
  cp_count_kernel<<<(unsigned)nblk, 256, 0, stream>>>(flags, n, bs);
  cp_scan_kernel<<<1, 1024, 0, stream>>>(bs, nblk, m_dev);
  cp_write_kernel<<<(unsigned)nblk, 256, 0, stream>>>(flags, n, bs, (const uint32_t *)rows_a, wa, (uint32_t *)out_a,
                                                      (const uint32_t *)rows_b, wb, (uint32_t *)out_b, pos);
  return check_launch("tsg_compact_rows");
}

int tsg_gather_rows(const void *src, int width, const int32_t *idx, int64_t n, void *out, tsg_stream_t stream) {
  if (n <= 0 || width <= 0) return TSG_OK;
  gather_rows_kernel<<<grid_for(n * width, 256), 256, 0, stream>>>((const uint32_t *)src, width, idx, n, nullptr, (uint32_t *)out);
  return check_launch("tsg_gather_rows");
}

int tsg_gather_rows_dev(const void *src, int width, const int32_t *idx, int64_t n_cap, const int32_t *n_dev, void *out,
                        tsg_stream_t stream) {
  if (n_cap <= 0 || width <= 0) return TSG_OK;
  gather_rows_kernel<<<grid_for(n_cap * width, 256), 256, 0, stream>>>((const uint32_t *)src, width, idx, n_cap, n_dev,
                                                                      (uint32_t *)out);
  return check_launch("tsg_gather_rows_dev");
}

int tsg_count(const int32_t *idx, int64_t n, int32_t *out, int64_t m, tsg_stream_t stream) {
  if (m > 0) TSG_CUDA(cudaMemsetAsync(out, 0, m * sizeof(int), stream));
  if (n <= 0 || m <= 0) return TSG_OK;
  count_kernel<<<grid_for(n, 256), 256, 0, stream>>>(idx, n, out, m);
  return check_launch("tsg_count");
}

int tsg_voxelize_fwd(const void *feats, int dtype, const int32_t *idx, const int32_t *counts, int64_t n, int c,
                     int64_t m, void *out, float *acc_ws, tsg_stream_t stream) {
  if (m <= 0 || c <= 0) return TSG_OK;
  float *acc = dtype == TSG_F32 ? (float *)out : acc_ws;
  if (!acc) { set_error("tsg_voxelize_fwd: fp32 accumulation workspace (m*c floats) required for 16-bit features"); return TSG_ERR_WORKSPACE; }
  TSG_CUDA(cudaMemsetAsync(acc, 0, (size_t)m * c * sizeof(float), stream));
  if (n > 0) {
    DISPATCH_DTYPE(dtype, (voxelize_fwd_kernel<T><<<grid_for(n * c, 256), 256, 0, stream>>>((const T *)feats, idx, counts, n, c, acc)));
  }
  if (dtype != TSG_F32) {
    DISPATCH_DTYPE(dtype, (cast_from_f32_kernel<T><<<grid_for(m * c, 256), 256, 0, stream>>>(acc, m * c, (T *)out)));
  }
  return check_launch("tsg_voxelize_fwd");
}

int tsg_voxelize_bwd(const void *top_grad, int dtype, const int32_t *idx, const int32_t *counts, int64_t n, int c,
                     void *bottom_grad, tsg_stream_t stream) {
  if (n <= 0 || c <= 0) return TSG_OK;
  DISPATCH_DTYPE(dtype, (voxelize_bwd_kernel<T><<<grid_for(n * c, 256), 256, 0, stream>>>((const T *)top_grad, idx, counts, n, c, (T *)bottom_grad)));
  return check_launch("tsg_voxelize_bwd");
}

int tsg_devoxelize_fwd(const void *feats, int dtype, const int32_t *idx8, const float *w8, int64_t n, int c,
                       void *out, tsg_stream_t stream) {
  if (n <= 0 || c <= 0) return TSG_OK;
  DISPATCH_DTYPE(dtype, (devoxelize_fwd_kernel<T><<<grid_for(n * c, 256), 256, 0, stream>>>((const T *)feats, idx8, w8, n, c, (T *)out)));
  return check_launch("tsg_devoxelize_fwd");
}

int tsg_devoxelize_bwd(const void *top_grad, int dtype, const int32_t *idx8, const float *w8, int64_t n, int c,
                       int64_t m, void *bottom_grad, float *acc_ws, tsg_stream_t stream) {
  if (m <= 0 || c <= 0) return TSG_OK;
  float *acc = dtype == TSG_F32 ? (float *)bottom_grad : acc_ws;
  if (!acc) { set_error("tsg_devoxelize_bwd: fp32 accumulation workspace required for 16-bit features"); return TSG_ERR_WORKSPACE; }
  TSG_CUDA(cudaMemsetAsync(acc, 0, (size_t)m * c * sizeof(float), stream));
  if (n > 0) {
    DISPATCH_DTYPE(dtype, (devoxelize_bwd_kernel<T><<<grid_for(n * c, 256), 256, 0, stream>>>((const T *)top_grad, idx8, w8, n, c, acc)));
  }
  if (dtype != TSG_F32) {
    DISPATCH_DTYPE(dtype, (cast_from_f32_kernel<T><<<grid_for(m * c, 256), 256, 0, stream>>>(acc, m * c, (T *)bottom_grad)));
  }
  return check_launch("tsg_devoxelize_bwd");
}

int tsg_point_query(const void *table, int64_t slots, const float *pcoords, int64_t n, int stride, int32_t *idx,
                    tsg_stream_t stream) {
  if (!pow2(slots) || stride <= 0) { set_error("tsg_point_query: bad table size or stride"); return TSG_ERR_INVALID; }
  if (n <= 0) return TSG_OK;
  point_query_kernel<<<grid_for(n, 256), 256, 0, stream>>>((const Slot *)table, (unsigned long long)(slots - 1),
                                                           (const float4 *)pcoords, n, stride, idx);
  return check_launch("tsg_point_query");
}

int tsg_trilinear_query(const void *table, int64_t slots, const float *pcoords, int64_t n, int stride, int nearest,
                        int32_t *idx8, float *w8, tsg_stream_t stream) {
  if (!pow2(slots) || stride <= 0) { set_error("tsg_trilinear_query: bad table size or stride"); return TSG_ERR_INVALID; }
  if (n <= 0) return TSG_OK;
  trilinear_query_kernel<<<grid_for(n, 128), 128, 0, stream>>>((const Slot *)table, (unsigned long long)(slots - 1),
                                                               (const float4 *)pcoords, n, stride, nearest, idx8, w8);
  return check_launch("tsg_trilinear_query");
}

int tsg_devoxelize_multi(int n_scales, const void *const *tables, const int64_t *slots, const int32_t *strides,
                         const float *const *feats, int c, const float *pcoords, const int32_t *rows, int64_t m,
                         float *out, int c_out, tsg_stream_t stream) {
  if (n_scales < 1 || n_scales > 4 || c % 4 || c_out > c || c_out <= 0) {
    set_error("tsg_devoxelize_multi: need 1..4 scales, c %% 4 == 0, 0 < c_out <= c");
    return TSG_ERR_INVALID;
  }
  if (m <= 0) return TSG_OK;
  DevoxScales sc;
  sc.count = n_scales;
  for (int i = 0; i < n_scales; ++i) {
    if (slots[i] <= 0 || (slots[i] & (slots[i] - 1))) {
      set_error("tsg_devoxelize_multi: slots must be a power of two");
      return TSG_ERR_INVALID;
    }
    sc.tab[i] = (const Slot *)tables[i];
    sc.mask[i] = (unsigned long long)(slots[i] - 1);
    sc.feats[i] = feats[i];
    sc.stride[i] = strides[i];
  }
  devoxelize_multi_kernel<<<grid_for((m + 3) / 4 * 4 * 8, 256), 256, 0, stream>>>(sc, (const float4 *)pcoords, rows, m, c, out,
                                                                                 c_out);
  return check_launch("tsg_devoxelize_multi");
}

int tsg_rescale_coords(const float *pcoords, int64_t n, float init_res, float after_res, float *out_f, int32_t *out_i,
                       tsg_stream_t stream) {
  if (n <= 0) return TSG_OK;
  rescale_coords_kernel<<<grid_for(n, 256), 256, 0, stream>>>((const float4 *)pcoords, n, init_res, after_res,
                                                              (float4 *)out_f, (int4 *)out_i);
  return check_launch("tsg_rescale_coords");
}

int tsg_cast_pad_bf16(const float *in, int64_t n, int c, int c_pad, void *out, tsg_stream_t stream) {
  if (c_pad < c) { set_error("tsg_cast_pad_bf16: c_pad < c"); return TSG_ERR_INVALID; }
  if (n <= 0) return TSG_OK;
  cast_pad_bf16_kernel<<<grid_for(n * c_pad, 256), 256, 0, stream>>>(in, n, nullptr, c, c_pad, (__nv_bfloat16 *)out);
  return check_launch("tsg_cast_pad_bf16");
}

int tsg_cast_pad_bf16_dev(const float *in, int64_t n_cap, const int32_t *n_dev, int c, int c_pad, void *out,
                          tsg_stream_t stream) {
  if (c_pad < c) { set_error("tsg_cast_pad_bf16_dev: c_pad < c"); return TSG_ERR_INVALID; }
  if (n_cap <= 0) return TSG_OK;
  cast_pad_bf16_kernel<<<grid_for(n_cap * c_pad, 256), 256, 0, stream>>>(in, n_cap, n_dev, c, c_pad, (__nv_bfloat16 *)out);
  return check_launch("tsg_cast_pad_bf16_dev");
}

}  // extern "C"
