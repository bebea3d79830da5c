// Hashing, key->index tables and kernel-map construction.
// All of this is HBM/L2-bound integer work: coalesced int4 coordinate loads, 16-byte table slots read with one
// 128-bit load per probe, outputs written k-major so that the convolution reads 128 consecutive rows per offset.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace tsg {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------- hashing
__global__ void hash_kernel(const int4 *__restrict__ coords, int64_t n, long long *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = __ldg(coords + i);
    out[i] = fnv60(c.x, c.y, c.z, c.w);
  }
}

// out[k*n + i]: consecutive threads -> consecutive i for one k (coalesced 8-byte stores, coords re-read from L2)
__global__ void kernel_hash_kernel(const int4 *__restrict__ coords, int64_t n, const int *__restrict__ offsets, int K,
                                   long long *__restrict__ out) {
  extern __shared__ int s_off[];
  for (int i = threadIdx.x; i < K * 3; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = __ldg(coords + i);
    for (int k = 0; k < K; ++k)
      out[(int64_t)k * n + i] = fnv60(c.x + s_off[3 * k], c.y + s_off[3 * k + 1], c.z + s_off[3 * k + 2], c.w);
  }
}

// ---------------------------------------------------------------- tables
__global__ void table_clear_kernel(Slot *tab, int64_t slots) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < slots; i += (int64_t)gridDim.x * blockDim.x) {
    Slot s;
    s.key = EMPTY_KEY;
    s.val = 0x7fffffff;
    s.pad = 0;
    tab[i] = s;
  }
}

__global__ void table_insert_keys_kernel(const long long *__restrict__ keys, int64_t n, Slot *tab,
                                         unsigned long long mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    table_insert(tab, mask, (unsigned long long)keys[i], (int)i);
}

__global__ void table_insert_coords_kernel(const int4 *__restrict__ coords, int64_t n, const int *__restrict__ n_dev,
                                           Slot *tab, unsigned long long mask, int *status) {
  n = dev_count(n_dev, n);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = __ldg(coords + i);
    if (!coord_in_range(c.x, c.y, c.z, c.w)) {
      if (status) atomicOr(status, 1);
      continue;
    }
    table_insert_coord(tab, mask, pack_coord(c.x, c.y, c.z, c.w), (int)i);
  }
}

__global__ void table_query_kernel(const Slot *__restrict__ tab, unsigned long long mask,
                                   const long long *__restrict__ q, int64_t nq, long long *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (long long)table_find(tab, mask, (unsigned long long)q[i]);
}

// ---------------------------------------------------------------- kernel maps
constexpr int KM_ROWS = 1024;  // output rows per CTA (one blockcnt column)
struct Offsets {
  int v[32 * 3];
};

// One CTA = 1024 consecutive output voxels (256 threads x 4 rows) x `ksplit` kernel offsets (grid.y covers the rest; small
// maps split the offsets over more CTAs so that every SM has probes in flight).  For a fixed offset a warp probes 32
// consecutive voxels, four independent probes per thread; nbr is written k-major with full 128-byte STREAMING stores
// (the 4 K N bytes of the map must not evict the table from L2: with plain stores the stride-1 map of the benchmark
// batch, 76 MB next to a 32 MB table in a 126 MB L2, ran anywhere between 240 and 840 us); hits are counted with a
// ballot, per-offset CTA totals go to blockcnt[k][cta].
__global__ void __launch_bounds__(256) kmap_build_kernel(const Slot *__restrict__ tab, unsigned long long mask,
                                                         const int4 *__restrict__ out_coords, int64_t n_out,
                                                         const int *__restrict__ n_dev, int64_t nbr_stride,
                                                         Offsets offs, int K, int ksplit, int *__restrict__ nbr,
                                                         int *__restrict__ nbsizes, int *__restrict__ blockcnt,
                                                         int64_t nblk, unsigned long long *__restrict__ row_keys, KeyBits kb) {
  __shared__ int s_cnt[32];
  n_out = dev_count(n_dev, n_out);
  if (threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * KM_ROWS;
  const int k0 = blockIdx.y * ksplit, k1 = min(K, k0 + ksplit);
  int4 c[4];
  bool ok[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t o = base + r * 256 + threadIdx.x;
    ok[r] = o < n_out;
    c[r] = ok[r] ? __ldg(out_coords + o) : make_int4(0, 0, 0, 0);
  }
  unsigned long long key[4] = {0ull, 0ull, 0ull, 0ull};   // "offset k has a neighbour" bits of the thread's rows (the sort key of
                                                          // tsg_kmap_sort_rows: emitted here, the rows are in registers anyway)
  for (int k = k0; k < k1; ++k) {
    const int dx = offs.v[3 * k], dy = offs.v[3 * k + 1], dz = offs.v[3 * k + 2];
    int hits = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int found = -1;
      if (ok[r]) {
        const int x = c[r].x + dx, y = c[r].y + dy, z = c[r].z + dz;
        if (coord_in_range(x, y, z, c[r].w)) found = table_find_coord(tab, mask, pack_coord(x, y, z, c[r].w));
        __stcs(nbr + (int64_t)k * nbr_stride + base + r * 256 + threadIdx.x, found);
      }
      key[r] |= (unsigned long long)(found >= 0) << kb.pos[k];
      hits += __popc(__ballot_sync(0xffffffffu, found >= 0));
    }
    if ((threadIdx.x & 31) == 0 && hits) atomicAdd(&s_cnt[k - k0], hits);
  }
  if (row_keys) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (!ok[r]) continue;
      unsigned long long *dst = row_keys + base + r * 256 + threadIdx.x;
      if (gridDim.y == 1) *dst = key[r];
      else if (key[r]) atomicOr(dst, key[r]);       // offsets split over grid.y: the keys were zeroed by the host function
    }
  }
  __syncthreads();
  if (threadIdx.x < k1 - k0) {
    const int v = s_cnt[threadIdx.x];
    blockcnt[(int64_t)(k0 + threadIdx.x) * nblk + blockIdx.x] = v;
    if (v) atomicAdd(&nbsizes[k0 + threadIdx.x], v);
  }
}

// exclusive scan of blockcnt (K*nblk, k-major) in place: single CTA, sequential over 1024-wide chunks
__global__ void __launch_bounds__(1024) scan_inplace_kernel(int *data, int64_t n) {
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
}

// write reference-format pairs: for (k, cta) the rows with a hit, ascending, at offset blockcnt[k][cta]
__global__ void __launch_bounds__(256) kmap_pairs_kernel(const int *__restrict__ nbr, int K, int64_t n_out,
                                                         const int *__restrict__ blockoff, int64_t nblk,
                                                         long long *__restrict__ nbmaps) {
  const int64_t base = (int64_t)blockIdx.x * KM_ROWS;
  for (int k = 0; k < K; ++k) {
    int64_t off = blockoff[(int64_t)k * nblk + blockIdx.x];
    for (int r = 0; r < 4; ++r) {  // rows r*256 .. r*256+255 in order
      const int64_t o = base + r * 256 + threadIdx.x;
      const int in = o < n_out ? nbr[(int64_t)k * n_out + o] : -1;
      int tot;
      const int pos = block_exclusive_scan<256>(in >= 0 ? 1 : 0, &tot);
      if (in >= 0) {
        nbmaps[2 * (off + pos)] = in;
        nbmaps[2 * (off + pos) + 1] = o;
      }
      off += tot;
    }
  }
}

// ---- compact per-offset pair lists for the weight gradient (conv_wgrad_tc.cu): for offset k the rows o with
// nbr[k, o] >= 0, ascending, as int2 {in row, out row}, all offsets back to back (k-major).  Three launches: per
// (k, 1024-row block) counts, their exclusive scan, the ordered write.  start[k] = first pair of offset k, start[K] = total.
__global__ void __launch_bounds__(256) pair_count_kernel(const int *__restrict__ nbr, int K, int64_t n_rows, int64_t stride,
                                                         int *__restrict__ blockcnt, int64_t nblk) {
  __shared__ int s_cnt[32];
  if (threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * KM_ROWS;
  for (int k = 0; k < K; ++k) {
    int hits = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t o = base + r * 256 + threadIdx.x;
      const int in = o < n_rows ? __ldg(nbr + (int64_t)k * stride + o) : -1;
      hits += __popc(__ballot_sync(0xffffffffu, in >= 0));
    }
    if ((threadIdx.x & 31) == 0 && hits) atomicAdd(&s_cnt[k], hits);
  }
  __syncthreads();
  if (threadIdx.x < K) blockcnt[(int64_t)threadIdx.x * nblk + blockIdx.x] = s_cnt[threadIdx.x];
}

__global__ void __launch_bounds__(256) pair_write_kernel(const int *__restrict__ nbr, int K, int64_t n_rows, int64_t stride,
                                                         const int *__restrict__ blockoff, int64_t nblk, int2 *__restrict__ pairs,
                                                         int *__restrict__ start, const int *__restrict__ last_cnt) {
  const int64_t base = (int64_t)blockIdx.x * KM_ROWS;
  for (int k = 0; k < K; ++k) {
    int64_t off = blockoff[(int64_t)k * nblk + blockIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) start[k] = (int)off;
    for (int r = 0; r < 4; ++r) {  // rows r*256 .. r*256+255 in order
      const int64_t o = base + r * 256 + threadIdx.x;
      const int in = o < n_rows ? __ldg(nbr + (int64_t)k * stride + o) : -1;
      int tot;
      const int pos = block_exclusive_scan<256>(in >= 0 ? 1 : 0, &tot);
      if (in >= 0) pairs[off + pos] = make_int2(in, (int)o);
      off += tot;
    }
    // the last block of the last offset closes the list: its exclusive offset + its own count (saved before the scan)
    if (k == K - 1 && blockIdx.x == nblk - 1 && threadIdx.x == 0) start[K] = blockoff[(int64_t)k * nblk + blockIdx.x] + *last_cnt;
  }
}

__global__ void save_last_kernel(const int *__restrict__ blockcnt, int64_t n, int *__restrict__ last_cnt) { *last_cnt = blockcnt[n - 1]; }

__global__ void fill_i32_kernel(int *p, int64_t n, int v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// nbr (K, in_stride) over n_out rows -> nbr_t (K, out_stride): nbr_t[k, nbr[k, o]] = o
__global__ void kmap_transpose_kernel(const int *__restrict__ nbr, int K, int64_t n_out, const int *__restrict__ n_dev,
                                      int64_t in_stride, int64_t out_stride, int *__restrict__ nbr_t) {
  n_out = dev_count(n_dev, n_out);
  const int64_t total = (int64_t)K * n_out;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = t / n_out, o = t - k * n_out;
    const int i = nbr[k * in_stride + o];
    if (i >= 0) nbr_t[k * out_stride + i] = (int)o;
  }
}

__global__ void kmap_from_pairs_kernel(const int *__restrict__ nbmaps, int64_t begin, int64_t count, int k,
                                       int transposed, int64_t n_rows_out, int *__restrict__ nbr) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
    const int a = nbmaps[2 * (begin + t)], b = nbmaps[2 * (begin + t) + 1];
    const int in = transposed ? b : a, out = transposed ? a : b;
    if (in >= 0 && out >= 0) nbr[(int64_t)k * n_rows_out + out] = in;
  }
}

// OR over the 128 rows of a tile of "offset k has a hit": lets the tensor-core kernel skip empty (tile, offset) pairs
__global__ void __launch_bounds__(128) kmap_tile_mask_kernel(const int *__restrict__ nbr, int K, int64_t n_out,
                                                             unsigned *__restrict__ tile_mask) {
  __shared__ unsigned s_mask;
  if (threadIdx.x == 0) s_mask = 0;
  __syncthreads();
  const int64_t o = (int64_t)blockIdx.x * 128 + threadIdx.x;
  unsigned m = 0;
  if (o < n_out)
    for (int k = 0; k < K; ++k) m |= (nbr[(int64_t)k * n_out + o] >= 0 ? 1u : 0u) << k;
#pragma unroll
  for (int s = 16; s; s >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, s);
  if ((threadIdx.x & 31) == 0) atomicOr(&s_mask, m);
  __syncthreads();
  if (threadIdx.x == 0) tile_mask[blockIdx.x] = s_mask;
}

static bool pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace tsg

using namespace tsg;

extern "C" {

int tsg_version(void) { return 100; }
const char *tsg_last_error(void) { return g_err; }

int tsg_hash(const int32_t *coords, int64_t n, int64_t *out, tsg_stream_t stream) {
  if (n <= 0) return TSG_OK;
  hash_kernel<<<grid_for(n, 256), 256, 0, stream>>>((const int4 *)coords, n, (long long *)out);
  return check_launch("tsg_hash");
}

int tsg_kernel_hash(const int32_t *coords, int64_t n, const int32_t *offsets, int k, int64_t *out,
                    tsg_stream_t stream) {
  if (n <= 0 || k <= 0) return TSG_OK;
  kernel_hash_kernel<<<grid_for(n, 256), 256, k * 3 * sizeof(int), stream>>>((const int4 *)coords, n, offsets, k,
                                                                             (long long *)out);
  return check_launch("tsg_kernel_hash");
}

int64_t tsg_table_slots(int64_t n) {
  int64_t s = 1024;
  while (s < 2 * n) s <<= 1;
  return s;
}

int tsg_table_build(const int64_t *keys, int64_t n, void *table, int64_t slots, tsg_stream_t stream) {
  if (!pow2(slots) || slots < 2 * n) {
    set_error("tsg_table_build: slots must be a power of two >= 2n");
    return TSG_ERR_INVALID;
  }
  table_clear_kernel<<<grid_for(slots, 256), 256, 0, stream>>>((Slot *)table, slots);
  if (n > 0)
    table_insert_keys_kernel<<<grid_for(n, 256), 256, 0, stream>>>((const long long *)keys, n, (Slot *)table,
                                                                   (unsigned long long)(slots - 1));
  return check_launch("tsg_table_build");
}

int tsg_table_query(const void *table, int64_t slots, const int64_t *queries, int64_t nq, int64_t *out,
                    tsg_stream_t stream) {
  if (!pow2(slots)) {
    set_error("tsg_table_query: slots must be a power of two");
    return TSG_ERR_INVALID;
  }
  if (nq <= 0) return TSG_OK;
  table_query_kernel<<<grid_for(nq, 256), 256, 0, stream>>>((const Slot *)table, (unsigned long long)(slots - 1),
                                                            (const long long *)queries, nq, (long long *)out);
  return check_launch("tsg_table_query");
}

int tsg_coord_table_build(const int32_t *coords, int64_t n, void *table, int64_t slots, int32_t *status,
                          tsg_stream_t stream) {
  if (!pow2(slots) || slots < 2 * n) {
    set_error("tsg_coord_table_build: slots must be a power of two >= 2n");
    return TSG_ERR_INVALID;
  }
  table_clear_kernel<<<grid_for(slots, 256), 256, 0, stream>>>((Slot *)table, slots);
  if (n > 0)
    table_insert_coords_kernel<<<grid_for(n, 256), 256, 0, stream>>>((const int4 *)coords, n, nullptr, (Slot *)table,
                                                                     (unsigned long long)(slots - 1), status);
  return check_launch("tsg_coord_table_build");
}

int tsg_coord_table_build_dev(const int32_t *coords, int64_t n_cap, const int32_t *n_dev, void *table, int64_t slots,
                              int32_t *status, tsg_stream_t stream) {
  if (!pow2(slots) || slots < 2 * n_cap) {
    set_error("tsg_coord_table_build_dev: slots must be a power of two >= 2 n_cap");
    return TSG_ERR_INVALID;
  }
  table_clear_kernel<<<grid_for(slots, 256), 256, 0, stream>>>((Slot *)table, slots);
  if (n_cap > 0)
    table_insert_coords_kernel<<<grid_for(n_cap, 256), 256, 0, stream>>>((const int4 *)coords, n_cap, n_dev, (Slot *)table,
                                                                         (unsigned long long)(slots - 1), status);
  return check_launch("tsg_coord_table_build_dev");
}

int64_t tsg_kmap_blocks(int64_t n_out) { return (n_out + KM_ROWS - 1) / KM_ROWS; }

static int kmap_build_impl(const void *table, int64_t slots, const int32_t *out_coords, int64_t n_out, const int *n_dev,
                           int64_t nbr_stride, const int32_t *offsets_host, int k, int32_t *nbr, int32_t *nbsizes,
                           int32_t *blockcnt, cudaStream_t stream, unsigned long long *row_keys = nullptr) {
  if (k <= 0 || k > 32 || !pow2(slots)) {
    set_error("tsg_kmap_build: need 1 <= K <= 32 and power-of-two slots");
    return TSG_ERR_INVALID;
  }
  TSG_CUDA(cudaMemsetAsync(nbsizes, 0, k * sizeof(int), stream));
  if (n_out <= 0) return TSG_OK;
  Offsets offs;
  for (int i = 0; i < 3 * k; ++i) offs.v[i] = offsets_host[i];
  const int64_t nblk = tsg_kmap_blocks(n_out);
  // enough CTAs to cover the chip a few times: split the offsets when the map has few row blocks
  int ksplit = k;
  while (ksplit > 1 && nblk * ((k + ksplit - 1) / ksplit) < 4LL * num_sms()) ksplit = (ksplit + 2) / 3;
  if (row_keys && ksplit < k) TSG_CUDA(cudaMemsetAsync(row_keys, 0, (size_t)n_out * sizeof(unsigned long long), stream));
  kmap_build_kernel<<<dim3((unsigned)nblk, (unsigned)((k + ksplit - 1) / ksplit)), 256, 0, stream>>>(
      (const Slot *)table, (unsigned long long)(slots - 1), (const int4 *)out_coords, n_out, n_dev, nbr_stride, offs, k, ksplit,
      nbr, nbsizes, blockcnt, nblk, row_keys, key_bits_for(k));
  return check_launch("tsg_kmap_build");
}

int tsg_kmap_build(const void *table, int64_t slots, const int32_t *out_coords, int64_t n_out,
                   const int32_t *offsets_host, int k, int32_t *nbr, int32_t *nbsizes, int32_t *blockcnt,
                   tsg_stream_t stream) {
  return kmap_build_impl(table, slots, out_coords, n_out, nullptr, n_out, offsets_host, k, nbr, nbsizes, blockcnt, stream);
}

int tsg_kmap_build_dev(const void *table, int64_t slots, const int32_t *out_coords, int64_t n_cap, const int32_t *n_dev,
                       const int32_t *offsets_host, int k, int32_t *nbr, int32_t *nbsizes, int32_t *blockcnt,
                       tsg_stream_t stream) {
  if (!n_dev) {
    set_error("tsg_kmap_build_dev: need the device row counter");
    return TSG_ERR_INVALID;
  }
  return kmap_build_impl(table, slots, out_coords, n_cap, n_dev, n_cap, offsets_host, k, nbr, nbsizes, blockcnt, stream);
}

int tsg_kmap_build_dev2(const void *table, int64_t slots, const int32_t *out_coords, int64_t n_cap, const int32_t *n_dev,
                       const int32_t *offsets_host, int k, int32_t *nbr, int32_t *nbsizes, int32_t *blockcnt,
                       uint64_t *row_keys, tsg_stream_t stream) {
  if (!n_dev) {
    set_error("tsg_kmap_build_dev2: need the device row counter");
    return TSG_ERR_INVALID;
  }
  return kmap_build_impl(table, slots, out_coords, n_cap, n_dev, n_cap, offsets_host, k, nbr, nbsizes, blockcnt, stream, (unsigned long long *)row_keys);
}

size_t tsg_kmap_pair_list_ws_bytes(int k, int64_t n_rows) { return ((size_t)k * tsg_kmap_blocks(n_rows) + 1) * sizeof(int); }

int tsg_kmap_pair_list(const int32_t *nbr, int k, int64_t n_rows, int64_t nbr_stride, int32_t *pairs, int32_t *start, void *ws,
                       size_t ws_bytes, cudaStream_t stream) {
  if (k <= 0 || k > 32 || n_rows <= 0 || nbr_stride < n_rows || (int64_t)k * n_rows >= (1ll << 31)) {
    set_error("tsg_kmap_pair_list: need 0 < K <= 32, n_rows > 0, nbr_stride >= n_rows and K * n_rows < 2^31");
    return TSG_ERR_INVALID;
  }
  if (ws_bytes < tsg_kmap_pair_list_ws_bytes(k, n_rows)) {
    set_error("tsg_kmap_pair_list: workspace too small");
    return TSG_ERR_INVALID;
  }
  const int64_t nblk = tsg_kmap_blocks(n_rows);
  int *blockcnt = (int *)ws, *last = blockcnt + (int64_t)k * nblk;
  pair_count_kernel<<<(unsigned)nblk, 256, 0, stream>>>(nbr, k, n_rows, nbr_stride, blockcnt, nblk);
  save_last_kernel<<<1, 1, 0, stream>>>(blockcnt, (int64_t)k * nblk, last);
  scan_inplace_kernel<<<1, 1024, 0, stream>>>(blockcnt, (int64_t)k * nblk);
  pair_write_kernel<<<(unsigned)nblk, 256, 0, stream>>>(nbr, k, n_rows, nbr_stride, blockcnt, nblk, (int2 *)pairs, start, last);
  return check_launch("tsg_kmap_pair_list");
}

int tsg_kmap_pairs(const int32_t *nbr, int k, int64_t n_out, int32_t *blockcnt, int64_t *nbmaps,
                   tsg_stream_t stream) {
  if (n_out <= 0 || k <= 0) return TSG_OK;
  const int64_t nblk = tsg_kmap_blocks(n_out);
  scan_inplace_kernel<<<1, 1024, 0, stream>>>(blockcnt, (int64_t)k * nblk);
  kmap_pairs_kernel<<<(unsigned)nblk, 256, 0, stream>>>(nbr, k, n_out, blockcnt, nblk, (long long *)nbmaps);
  return check_launch("tsg_kmap_pairs");
}

int tsg_kmap_transpose(const int32_t *nbr, int k, int64_t n_out, int64_t n_in, int32_t *nbr_t,
                       tsg_stream_t stream) {
  if (k <= 0) return TSG_OK;
  if (n_in > 0) fill_i32_kernel<<<grid_for(k * n_in, 256), 256, 0, stream>>>(nbr_t, (int64_t)k * n_in, -1);
  if (n_out > 0 && n_in > 0)
    kmap_transpose_kernel<<<grid_for(k * n_out, 256), 256, 0, stream>>>(nbr, k, n_out, nullptr, n_out, n_in, nbr_t);
  return check_launch("tsg_kmap_transpose");
}

int tsg_kmap_transpose_dev(const int32_t *nbr, int k, int64_t n_out_cap, const int32_t *n_out_dev, int64_t n_in_cap,
                           int32_t *nbr_t, tsg_stream_t stream) {
  if (k <= 0) return TSG_OK;
  if (n_in_cap > 0) fill_i32_kernel<<<grid_for(k * n_in_cap, 256), 256, 0, stream>>>(nbr_t, (int64_t)k * n_in_cap, -1);
  if (n_out_cap > 0 && n_in_cap > 0)
    kmap_transpose_kernel<<<grid_for(k * n_out_cap, 256), 256, 0, stream>>>(nbr, k, n_out_cap, n_out_dev, n_out_cap, n_in_cap,
                                                                           nbr_t);
  return check_launch("tsg_kmap_transpose_dev");
}

int tsg_kmap_from_pairs(const int32_t *nbmaps, const int32_t *nbsizes_host, int k, int transposed,
                        int64_t n_rows_out, int32_t *nbr, tsg_stream_t stream) {
  if (k <= 0 || n_rows_out <= 0) return TSG_OK;
  fill_i32_kernel<<<grid_for(k * n_rows_out, 256), 256, 0, stream>>>(nbr, (int64_t)k * n_rows_out, -1);
  int64_t begin = 0;
  for (int i = 0; i < k; ++i) {
    const int64_t cnt = nbsizes_host[i];
    if (cnt > 0)
      kmap_from_pairs_kernel<<<grid_for(cnt, 256), 256, 0, stream>>>(nbmaps, begin, cnt, i, transposed, n_rows_out,
                                                                     nbr);
    begin += cnt;
  }
  return check_launch("tsg_kmap_from_pairs");
}

int tsg_kmap_tile_mask(const int32_t *nbr, int k, int64_t n_out, uint32_t *tile_mask, tsg_stream_t stream) {
  if (n_out <= 0) return TSG_OK;
  kmap_tile_mask_kernel<<<(unsigned)((n_out + 127) / 128), 128, 0, stream>>>(nbr, k, n_out, tile_mask);
  return check_launch("tsg_kmap_tile_mask");
}

}  // extern "C"
