// Stable LSD radix sort (64-bit keys, 32-bit payload) and the unique-voxel builders on top of it
// (SURVEY §8 a5 sparse_quantize, a12 initial_voxelize ordering, a16 spdownsample).
// Sorting packed (b,x,y,z) keys yields the reference's lexicographic voxel order directly, and stability makes the
// head of every run the FIRST occurrence in input order (np.unique's return_index contract).
#include "common.cuh"

namespace tsg {

// One-sweep LSD radix sort, 9-bit digits: ONE histogram kernel for all passes, then one kernel per pass that ranks its
// tile, obtains its global offsets by decoupled look-back over the tiles before it and scatters.  (The first version
// used three kernels per 8-bit pass — tile histogram, per-digit scan, scatter with three block barriers per element
// round — and cost ~40 us per pass whatever n: 53 passes were 21 % of a benchmark step, profiles/README.md.)
constexpr int RS_BITS = 9;
constexpr int RS_BINS = 1 << RS_BITS;  // 512 = threads per CTA: thread d owns digit d
constexpr int RS_THREADS = RS_BINS;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per tile, 256 consecutive keys per warp
constexpr int RS_MAX_PASSES = 8;                // 64-bit keys: ceil(64 / 9)

struct RsPasses {
  int bit[RS_MAX_PASSES];
  unsigned mask[RS_MAX_PASSES];
  int count;
};

// global digit histograms of every pass in one read of the keys: ghist[p][d]
__global__ void __launch_bounds__(RS_THREADS) rs_global_hist_kernel(const unsigned long long *__restrict__ keys, int64_t n,
                                                                    const int *__restrict__ n_dev, RsPasses ps,
                                                                    int *__restrict__ ghist) {
  __shared__ int hist[RS_MAX_PASSES][RS_BINS];
  n = dev_count(n_dev, n);
  for (int p = 0; p < ps.count; ++p) hist[p][threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = keys[i];
    for (int p = 0; p < ps.count; ++p) atomicAdd(&hist[p][(unsigned)(key >> ps.bit[p]) & ps.mask[p]], 1);
  }
  __syncthreads();
  for (int p = 0; p < ps.count; ++p) {
    const int v = hist[p][threadIdx.x];
    if (v) atomicAdd(&ghist[p * RS_BINS + threadIdx.x], v);
  }
}

constexpr unsigned LOOK_LOCAL = 1u << 30, LOOK_PREFIX = 2u << 30, LOOK_VALUE = (1u << 30) - 1;

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned *p, unsigned v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(RS_THREADS) rs_pass_kernel(const unsigned long long *__restrict__ keys_in,
                                                             const unsigned *__restrict__ vals_in, int64_t n,
                                                             const int *__restrict__ n_dev, int bit,
                                                             unsigned dmask, const int *__restrict__ ghist,
                                                             unsigned *__restrict__ look, int *__restrict__ tile_counter,
                                                             unsigned long long *__restrict__ keys_out,
                                                             unsigned *__restrict__ vals_out) {
  __shared__ int wh[RS_WARPS][RS_BINS];  // per-warp digit counts, then per-warp exclusive offsets inside the tile
  __shared__ int base[RS_BINS];          // first output slot of digit d for this tile
  __shared__ int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1);  // tiles are numbered in start order: look-back never waits on an unstarted tile
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) wh[w][tid] = 0;
  __syncthreads();
  const int tile = s_tile;
  n = dev_count(n_dev, n);
  if ((int64_t)tile * RS_TILE >= n) return;  // capacity-sized grid: tiles past the device count have no keys and no successors
  const int64_t w0 = (int64_t)tile * RS_TILE + warp * (32 * RS_ITEMS);
  const unsigned lane_lt = (1u << lane) - 1;
  unsigned long long key[RS_ITEMS];
  int rk[RS_ITEMS];  // rank among the warp's keys with the same digit (stable: round-major, then lane)
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const int64_t i = w0 + j * 32 + lane;
    const bool valid = i < n;
    key[j] = valid ? keys_in[i] : 0ull;
    const int d = valid ? (int)((unsigned)(key[j] >> bit) & dmask) : RS_BINS + lane;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int r = __popc(peers & lane_lt);
    const int prior = valid ? wh[warp][d] : 0;
    __syncwarp();
    if (valid && r == 0) wh[warp][d] = prior + __popc(peers);
    __syncwarp();
    rk[j] = prior + r;
  }
  __syncthreads();
  int sum = 0;  // thread tid = digit tid: exclusive prefix over the warps, tile total
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) {
    const int t = wh[w][tid];
    wh[w][tid] = sum;
    sum += t;
  }
  const int gex = block_exclusive_scan<RS_THREADS>(__ldg(ghist + tid), nullptr);  // first slot of digit tid overall
  // decoupled look-back: keys with digit tid in the tiles before this one
  unsigned *mine = look + (int64_t)tile * RS_BINS + tid;
  int excl = 0;
  if (tile == 0) {
    st_volatile_u32(mine, LOOK_PREFIX | (unsigned)sum);
  } else {
    st_volatile_u32(mine, LOOK_LOCAL | (unsigned)sum);
    int t = tile - 1;
    bool done = false;
    while (!done) {
      unsigned v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = t - q >= 0 ? ld_volatile_u32(look + (int64_t)(t - q) * RS_BINS + tid) : LOOK_PREFIX;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (done || (v[q] >> 30) == 0) break;  // not published yet: poll again from here
        excl += (int)(v[q] & LOOK_VALUE);
        --t;
        done = (v[q] >> 30) == 2;
      }
    }
    st_volatile_u32(mine, LOOK_PREFIX | (unsigned)(excl + sum));
  }
  base[tid] = gex + excl;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const int64_t i = w0 + j * 32 + lane;
    if (i < n) {
      const int d = (int)((unsigned)(key[j] >> bit) & dmask);
      const int pos = base[d] + wh[warp][d] + rk[j];
      keys_out[pos] = key[j];
      vals_out[pos] = vals_in ? vals_in[i] : (unsigned)i;
    }
  }
}

struct SortWs {
  unsigned long long *keys_tmp;
  unsigned *vals_tmp;
  int *ghist;         // [RS_MAX_PASSES][RS_BINS]
  int *tile_counter;  // [RS_MAX_PASSES]
  unsigned *look;     // [RS_MAX_PASSES][ntiles][RS_BINS]
  size_t zero_bytes;  // ghist .. end of look are cleared by one memset per sort
};

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static size_t sort_ws_layout(int64_t n, char *base, SortWs *ws) {
  const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  size_t off = 0;
  if (ws) ws->keys_tmp = (unsigned long long *)(base + off);
  off += align256((size_t)n * 8);
  if (ws) ws->vals_tmp = (unsigned *)(base + off);
  off += align256((size_t)n * 4);
  const size_t z0 = off;
  if (ws) ws->ghist = (int *)(base + off);
  off += align256((size_t)RS_MAX_PASSES * RS_BINS * 4);
  if (ws) ws->tile_counter = (int *)(base + off);
  off += align256(RS_MAX_PASSES * 4);
  if (ws) ws->look = (unsigned *)(base + off);
  off += align256((size_t)RS_MAX_PASSES * (ntiles > 0 ? ntiles : 1) * RS_BINS * 4);
  if (ws) ws->zero_bytes = off - z0;
  return off;
}

int sort_pairs(const unsigned long long *keys_in, const unsigned *vals_in, int64_t n, int begin_bit,
               int end_bit, unsigned long long *keys_out, unsigned *vals_out, void *ws_mem, size_t ws_bytes,
               cudaStream_t stream, const int *n_dev) {
  if (n <= 0) return TSG_OK;
  if (n >= (1ll << 30)) {
    set_error("tsg_sort_pairs: n must be < 2^30");
    return TSG_ERR_INVALID;
  }
  SortWs ws;
  if (sort_ws_layout(n, (char *)ws_mem, &ws) > ws_bytes) {
    set_error("tsg_sort_pairs: workspace too small");
    return TSG_ERR_WORKSPACE;
  }
  if (end_bit > 64) end_bit = 64;
  const int passes = (end_bit - begin_bit + RS_BITS - 1) / RS_BITS;
  if (passes <= 0 || begin_bit < 0) {
    set_error("tsg_sort_pairs: empty bit range");
    return TSG_ERR_INVALID;
  }
  const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  RsPasses ps;
  ps.count = passes;
  for (int p = 0; p < passes; ++p) {
    ps.bit[p] = begin_bit + RS_BITS * p;
    const int w = end_bit - ps.bit[p] < RS_BITS ? end_bit - ps.bit[p] : RS_BITS;
    ps.mask[p] = (1u << w) - 1u;
  }
  TSG_CUDA(cudaMemsetAsync(ws.ghist, 0, ws.zero_bytes, stream));
  const int64_t hist_ctas = ntiles < 4 * num_sms() ? ntiles : 4 * num_sms();
  rs_global_hist_kernel<<<(unsigned)hist_ctas, RS_THREADS, 0, stream>>>(keys_in, n, n_dev, ps, ws.ghist);
  const unsigned long long *src_k = keys_in;
  const unsigned *src_v = vals_in;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) & 1) == 0;
    unsigned long long *dst_k = to_out ? keys_out : ws.keys_tmp;
    unsigned *dst_v = to_out ? vals_out : ws.vals_tmp;
    rs_pass_kernel<<<(unsigned)ntiles, RS_THREADS, 0, stream>>>(src_k, src_v, n, n_dev, ps.bit[p], ps.mask[p],
                                                               ws.ghist + p * RS_BINS, ws.look + (size_t)p * ntiles * RS_BINS,
                                                               ws.tile_counter + p, dst_k, dst_v);
    src_k = dst_k;
    src_v = dst_v;
  }
  return check_launch("tsg_sort_pairs");
}


// ---------------------------------------------------------------- mask-sorted tile rows for the tensor-core convolution
// key[o] = K-bit mask "offset k has a neighbour" of output row o (coalesced k-major reads of nbr)
// Bit position of offset k inside the sort key.  Rows are ordered so that RARE offsets are the most significant bits:
// rows that own a rare neighbour then share tiles and the offset is skipped everywhere else.  For 3x3x3 kernels the
// static order corners > outer-plane edges > outer face centres > middle-plane corners > middle-plane edges > centre
// matches the measured frequencies on LiDAR scans (7 / 12 / 18 / 36 / 47 / 100 %) and cuts the active (tile, offset)
// pairs from 0.40 (plain integer order of the mask) to 0.33; other kernel sizes keep the natural order.
__global__ void row_mask_keys_kernel(const int *__restrict__ nbr, int K, int64_t n_out, const int *__restrict__ n_dev,
                                     int64_t in_stride, KeyBits kb, unsigned long long *__restrict__ keys) {
  n_out = dev_count(n_dev, n_out);
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_out; o += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long m = 0;
    for (int k = 0; k < K; ++k) m |= (unsigned long long)(__ldg(nbr + (int64_t)k * in_stride + o) >= 0) << kb.pos[k];
    keys[o] = m;
  }
}
// nbr_sorted[k, r] = nbr[k, perm[r]] for r < n_out and -1 in the padding rows [n_out, out_stride); the same CTA (256 tile
// rows = two tiles) ORs the sorted keys of its rows into the two tile masks, key bits mapped back to offsets (one launch
// instead of two: the 13 mask kernels of a step were ~10 us of launch latency each).
__global__ void __launch_bounds__(256) permute_nbr_kernel(const int *__restrict__ nbr, const int *__restrict__ perm, int K, int64_t n_out,
                                                          const int *__restrict__ n_dev, int64_t in_stride, int64_t out_stride,
                                                          int *__restrict__ nbr_sorted, const unsigned long long *__restrict__ keys_sorted,
                                                          KeyBits kb, unsigned *__restrict__ tile_mask, int64_t n_tiles) {
  __shared__ unsigned long long s_or[8];
  n_out = dev_count(n_dev, n_out);
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  unsigned long long m = r < n_out ? keys_sorted[r] : 0ull;
  if (r < out_stride) {
    if (r < n_out) {
      const int o = __ldg(perm + r);
      for (int k = 0; k < K; ++k) nbr_sorted[(int64_t)k * out_stride + r] = __ldg(nbr + (int64_t)k * in_stride + o);
    } else {
      for (int k = 0; k < K; ++k) nbr_sorted[(int64_t)k * out_stride + r] = -1;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
  if ((threadIdx.x & 31) == 0) s_or[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 2) {
    const int64_t tile = (int64_t)blockIdx.x * 2 + threadIdx.x;
    if (tile < n_tiles) {
      const unsigned long long mm = s_or[4 * threadIdx.x] | s_or[4 * threadIdx.x + 1] | s_or[4 * threadIdx.x + 2] | s_or[4 * threadIdx.x + 3];
      unsigned out = 0;
      for (int k = 0; k < K; ++k) out |= (unsigned)((mm >> kb.pos[k]) & 1ull) << k;
      tile_mask[tile] = out;
    }
  }
}
// ---------------------------------------------------------------- unique voxels
// fb = bit widths of (x, y, z, b) when the caller knows every coordinate lies in [0, 2^bits): keys are then packed
// densely so the sort needs ceil(sum/8) passes instead of 8; fb.x == 0 selects the general 64-bit packing.
__device__ inline unsigned long long pack_dense(int4 c, int4 fb) {
  return ((((unsigned long long)(unsigned)c.w << fb.x | (unsigned)c.x) << fb.y | (unsigned)c.y) << fb.z) | (unsigned)c.z;
}
__device__ inline int4 unpack_dense(unsigned long long k, int4 fb) {
  int4 c;
  c.z = (int)(k & ((1ull << fb.z) - 1)); k >>= fb.z;
  c.y = (int)(k & ((1ull << fb.y) - 1)); k >>= fb.y;
  c.x = (int)(k & ((1ull << fb.x) - 1)); k >>= fb.x;
  c.w = (int)k;
  return c;
}

__global__ void make_coord_keys_kernel(const int4 *__restrict__ coords, int64_t n, const int *__restrict__ n_dev,
                                       int trunc_stride, int4 fb, unsigned long long *__restrict__ keys, int *status) {
  n = dev_count(n_dev, n);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    if (trunc_stride > 0) {  // torch.div(...).trunc() * stride: C integer division truncates toward zero as well
      c.x = (c.x / trunc_stride) * trunc_stride;
      c.y = (c.y / trunc_stride) * trunc_stride;
      c.z = (c.z / trunc_stride) * trunc_stride;
    }
    if (fb.x > 0) {
      if ((unsigned)c.x >> fb.x || (unsigned)c.y >> fb.y || (unsigned)c.z >> fb.z || (unsigned)c.w >> fb.w) {
        if (status) atomicOr(status, 1);
        c = make_int4(0, 0, 0, 0);
      }
      keys[i] = pack_dense(c, fb);
      continue;
    }
    if (!coord_in_range(c.x, c.y, c.z, c.w)) {
      if (status) atomicOr(status, 1);
      c = make_int4(0, 0, 0, 0);
    }
    keys[i] = pack_coord(c.x, c.y, c.z, c.w);
  }
}

__global__ void make_hash_keys_kernel(const int4 *__restrict__ coords, int64_t n, unsigned long long *__restrict__ keys) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = __ldg(coords + i);
    keys[i] = (unsigned long long)fnv60(c.x, c.y, c.z, c.w);
  }
}

constexpr int UQ_ROWS = 1024;

__global__ void __launch_bounds__(256) uq_count_kernel(const unsigned long long *__restrict__ skeys, int64_t n,
                                                       const int *__restrict__ n_dev, int *__restrict__ blocksum) {
  n = dev_count(n_dev, n);
  const int64_t base = (int64_t)blockIdx.x * UQ_ROWS + threadIdx.x * 4;
  int c = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t j = base + r;
    if (j < n) c += (j == 0 || skeys[j] != skeys[j - 1]) ? 1 : 0;
  }
  int tot;
  block_exclusive_scan<256>(c, &tot);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = tot;
}

// total_out = number of runs, clamped to out_cap (status bit 1 = the consumer's capacity was exceeded)
__global__ void __launch_bounds__(1024) uq_scan_kernel(int *data, int64_t n, int *total_out, int64_t out_cap, int *status) {
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
  if (threadIdx.x == 0 && total_out) {
    if (out_cap > 0 && carry > out_cap) {
      if (status) atomicOr(status, 2);
      carry = (int)out_cap;
    }
    *total_out = carry;
  }
}

// hash_mode: coordinates are not recoverable from the key -> copy them from the first member of the run
__global__ void __launch_bounds__(256) uq_write_kernel(const unsigned long long *__restrict__ skeys,
                                                       const unsigned *__restrict__ sidx, int64_t n,
                                                       const int *__restrict__ n_dev, int64_t out_cap,
                                                       const int *__restrict__ blockoff, int hash_mode, int4 fb,
                                                       const int4 *__restrict__ in_coords, int4 *__restrict__ out_coords,
                                                       int *__restrict__ first_idx, int *__restrict__ inverse) {
  n = dev_count(n_dev, n);
  const int64_t base = (int64_t)blockIdx.x * UQ_ROWS + threadIdx.x * 4;
  bool head[4];
  int c = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t j = base + r;
    head[r] = j < n && (j == 0 || skeys[j] != skeys[j - 1]);
    c += head[r] ? 1 : 0;
  }
  int vid = block_exclusive_scan<256>(c, nullptr) + blockoff[blockIdx.x] - 1;  // id of the run open before base
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t j = base + r;
    if (j >= n) break;
    const unsigned src = sidx[j];
    if (head[r]) {
      ++vid;
      if (out_cap > 0 && vid >= out_cap) continue;   // capacity overflow (flagged by uq_scan_kernel): drop, stay in bounds
      if (out_coords)
        out_coords[vid] = hash_mode ? __ldg(in_coords + src) : (fb.x > 0 ? unpack_dense(skeys[j], fb) : unpack_coord(skeys[j]));
      if (first_idx) first_idx[vid] = (int)src;
    }
    if (inverse) inverse[src] = vid;
  }
}

struct UniqueWs {
  unsigned long long *keys, *skeys;
  unsigned *sidx;
  int *blocksum;
  char *sort_ws;
  size_t sort_bytes;
};

static size_t unique_ws_layout(int64_t n, char *base, UniqueWs *ws) {
  size_t off = 0;
  if (ws) ws->keys = (unsigned long long *)(base + off);
  off += align256((size_t)n * 8);
  if (ws) ws->skeys = (unsigned long long *)(base + off);
  off += align256((size_t)n * 8);
  if (ws) ws->sidx = (unsigned *)(base + off);
  off += align256((size_t)n * 4);
  if (ws) ws->blocksum = (int *)(base + off);
  off += align256((size_t)((n + UQ_ROWS - 1) / UQ_ROWS + 1) * 4);
  const size_t sb = sort_ws_layout(n, nullptr, nullptr);
  if (ws) {
    ws->sort_ws = base + off;
    ws->sort_bytes = sb;
  }
  off += sb;
  return off;
}

static int unique_impl(const int32_t *in_coords, int64_t n, int trunc_stride, bool hash_mode, int4 fb, int32_t *out_coords,
                       int32_t *first_idx, int32_t *inverse, int32_t *m_dev, int32_t *status, void *ws_mem,
                       size_t ws_bytes, cudaStream_t stream, const int *n_dev = nullptr, int64_t out_cap = 0) {
  if (n <= 0) {
    if (m_dev) TSG_CUDA(cudaMemsetAsync(m_dev, 0, sizeof(int), stream));
    return TSG_OK;
  }
  UniqueWs ws;
  if (unique_ws_layout(n, (char *)ws_mem, &ws) > ws_bytes) {
    set_error("tsg_unique: workspace too small");
    return TSG_ERR_WORKSPACE;
  }
  if (hash_mode)
    make_hash_keys_kernel<<<grid_for(n, 256), 256, 0, stream>>>((const int4 *)in_coords, n, ws.keys);
  else
    make_coord_keys_kernel<<<grid_for(n, 256), 256, 0, stream>>>((const int4 *)in_coords, n, n_dev, trunc_stride, fb, ws.keys,
                                                                status);
  const int key_bits = hash_mode ? 60 : (fb.x > 0 ? fb.x + fb.y + fb.z + fb.w : 64);
  int rc = sort_pairs(ws.keys, nullptr, n, 0, key_bits, ws.skeys, ws.sidx, ws.sort_ws, ws.sort_bytes, stream, n_dev);
  if (rc) return rc;
  const int64_t nblk = (n + UQ_ROWS - 1) / UQ_ROWS;
  uq_count_kernel<<<(unsigned)nblk, 256, 0, stream>>>(ws.skeys, n, n_dev, ws.blocksum);
  uq_scan_kernel<<<1, 1024, 0, stream>>>(ws.blocksum, nblk, m_dev, out_cap, status);
  uq_write_kernel<<<(unsigned)nblk, 256, 0, stream>>>(ws.skeys, ws.sidx, n, n_dev, out_cap, ws.blocksum, hash_mode ? 1 : 0, fb,
                                                      (const int4 *)in_coords, (int4 *)out_coords, first_idx, inverse);
  return check_launch("tsg_unique");
}

}  // namespace tsg

using namespace tsg;

extern "C" {

size_t tsg_sort_ws_bytes(int64_t n) { return sort_ws_layout(n > 0 ? n : 1, nullptr, nullptr); }

int tsg_sort_pairs(const uint64_t *keys_in, const uint32_t *vals_in, int64_t n, int begin_bit, int end_bit,
                   uint64_t *keys_out, uint32_t *vals_out, void *ws, size_t ws_bytes, tsg_stream_t stream) {
  return sort_pairs((const unsigned long long *)keys_in, vals_in, n, begin_bit, end_bit, (unsigned long long *)keys_out,
                    vals_out, ws, ws_bytes, stream);
}

size_t tsg_kmap_sort_ws_bytes(int64_t n_out) {
  const int64_t n = n_out > 0 ? n_out : 1;
  return 2 * align256((size_t)n * 8) + sort_ws_layout(n, nullptr, nullptr);
}

int64_t tsg_kmap_sort_stride(int64_t n_out) { return (n_out + 255) / 256 * 256; }

static int kmap_sort_rows_impl(const int32_t *nbr, int k, int64_t n_out, const int *n_dev, int64_t in_stride, int32_t *perm,
                               int32_t *nbr_sorted, int64_t out_stride, uint32_t *tile_mask, void *ws, size_t ws_bytes,
                               cudaStream_t stream, const unsigned long long *row_keys = nullptr) {
  if (out_stride < n_out) {
    set_error("tsg_kmap_sort_rows: out_stride < n_out");
    return TSG_ERR_INVALID;
  }
  if (k <= 0 || k > 32) {
    set_error("tsg_kmap_sort_rows: need 1 <= K <= 32");
    return TSG_ERR_INVALID;
  }
  if (n_out <= 0) return TSG_OK;
  if (ws_bytes < tsg_kmap_sort_ws_bytes(n_out)) {
    set_error("tsg_kmap_sort_rows: workspace too small");
    return TSG_ERR_WORKSPACE;
  }
  char *base = (char *)ws;
  unsigned long long *keys = (unsigned long long *)base;
  unsigned long long *keys_sorted = (unsigned long long *)(base + align256((size_t)n_out * 8));
  char *sort_ws = base + 2 * align256((size_t)n_out * 8);
  if (row_keys) keys = const_cast<unsigned long long *>(row_keys);   // emitted by the kernel-map build (tsg_kmap_build_dev2)
  else row_mask_keys_kernel<<<grid_for(n_out, 256), 256, 0, stream>>>(nbr, k, n_out, n_dev, in_stride, key_bits_for(k), keys);
  const int rc = sort_pairs(keys, nullptr, n_out, 0, k, keys_sorted, (unsigned *)perm, sort_ws,
                            ws_bytes - 2 * align256((size_t)n_out * 8), stream, n_dev);
  if (rc != TSG_OK) return rc;
  const int64_t n_tiles = (n_out + 127) / 128, rows = out_stride > n_tiles * 128 ? out_stride : n_tiles * 128;
  permute_nbr_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, stream>>>(nbr, perm, k, n_out, n_dev, in_stride, out_stride, nbr_sorted,
                                                                       keys_sorted, key_bits_for(k), tile_mask, n_tiles);
  return check_launch("tsg_kmap_sort_rows");
}

int tsg_kmap_sort_rows(const int32_t *nbr, int k, int64_t n_out, int32_t *perm, int32_t *nbr_sorted, int64_t out_stride,
                       uint32_t *tile_mask, void *ws, size_t ws_bytes, tsg_stream_t stream) {
  return kmap_sort_rows_impl(nbr, k, n_out, nullptr, n_out, perm, nbr_sorted, out_stride, tile_mask, ws, ws_bytes, stream);
}

int tsg_kmap_sort_rows_dev(const int32_t *nbr, int k, int64_t n_cap, const int32_t *n_dev, int64_t in_stride, int32_t *perm,
                           int32_t *nbr_sorted, int64_t out_stride, uint32_t *tile_mask, void *ws, size_t ws_bytes,
                           tsg_stream_t stream) {
  return kmap_sort_rows_impl(nbr, k, n_cap, n_dev, in_stride, perm, nbr_sorted, out_stride, tile_mask, ws, ws_bytes, stream);
}

int tsg_kmap_sort_rows_dev2(const int32_t *nbr, int k, int64_t n_cap, const int32_t *n_dev, int64_t in_stride, int32_t *perm,
                            int32_t *nbr_sorted, int64_t out_stride, uint32_t *tile_mask, const uint64_t *row_keys, void *ws,
                            size_t ws_bytes, tsg_stream_t stream) {
  return kmap_sort_rows_impl(nbr, k, n_cap, n_dev, in_stride, perm, nbr_sorted, out_stride, tile_mask, ws, ws_bytes, stream,
                             (const unsigned long long *)row_keys);
}

size_t tsg_unique_ws_bytes(int64_t n) { return unique_ws_layout(n > 0 ? n : 1, nullptr, nullptr); }

int tsg_unique_coords(const int32_t *in_coords, int64_t n, int trunc_stride, const int32_t *field_bits_host,
                      int32_t *out_coords, int32_t *first_idx, int32_t *inverse, int32_t *m_dev, int32_t *status,
                      void *ws, size_t ws_bytes, tsg_stream_t stream) {
  int4 fb = make_int4(0, 0, 0, 0);
  if (field_bits_host) {
    fb = make_int4(field_bits_host[0], field_bits_host[1], field_bits_host[2], field_bits_host[3]);
    if (fb.x <= 0 || fb.y <= 0 || fb.z <= 0 || fb.w <= 0 || fb.x > 19 || fb.y > 19 || fb.z > 19 || fb.w > 7) {
      set_error("tsg_unique_coords: field bits must be in 1..19 (x,y,z) and 1..7 (b)");
      return TSG_ERR_INVALID;
    }
  }
  return unique_impl(in_coords, n, trunc_stride, false, fb, out_coords, first_idx, inverse, m_dev, status, ws,
                     ws_bytes, stream);
}

int tsg_unique_coords_dev(const int32_t *in_coords, int64_t n_cap, const int32_t *n_dev, int trunc_stride,
                          const int32_t *field_bits_host, int32_t *out_coords, int64_t out_cap, int32_t *first_idx,
                          int32_t *inverse, int32_t *m_dev, int32_t *status, void *ws, size_t ws_bytes, tsg_stream_t stream) {
  int4 fb = make_int4(0, 0, 0, 0);
  if (field_bits_host) {
    fb = make_int4(field_bits_host[0], field_bits_host[1], field_bits_host[2], field_bits_host[3]);
    if (fb.x <= 0 || fb.y <= 0 || fb.z <= 0 || fb.w <= 0 || fb.x > 19 || fb.y > 19 || fb.z > 19 || fb.w > 7) {
      set_error("tsg_unique_coords_dev: field bits must be in 1..19 (x,y,z) and 1..7 (b)");
      return TSG_ERR_INVALID;
    }
  }
  if (!n_dev || !m_dev || out_cap <= 0) {
    set_error("tsg_unique_coords_dev: need the device counters and a positive output capacity");
    return TSG_ERR_INVALID;
  }
  return unique_impl(in_coords, n_cap, trunc_stride, false, fb, out_coords, first_idx, inverse, m_dev, status, ws, ws_bytes,
                     stream, n_dev, out_cap);
}

int tsg_unique_hash(const int32_t *in_coords, int64_t n, int32_t *out_coords, int32_t *first_idx, int32_t *inverse,
                    int32_t *m_dev, void *ws, size_t ws_bytes, tsg_stream_t stream) {
  return unique_impl(in_coords, n, 0, true, make_int4(0, 0, 0, 0), out_coords, first_idx, inverse, m_dev, nullptr, ws,
                     ws_bytes, stream);
}

}  // extern "C"
