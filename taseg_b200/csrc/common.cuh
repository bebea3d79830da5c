// Shared device/host helpers for the taseg_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/taseg_b200.h"

namespace tsg {

void set_error(const char *fmt, ...);

inline int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return TSG_ERR_CUDA;
  }
  return TSG_OK;
}

#define TSG_CUDA(call)                                               \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) {                                        \
      tsg::set_error("%s: %s", #call, cudaGetErrorString(e__));      \
      return TSG_ERR_CUDA;                                           \
    }                                                                \
  } while (0)

inline int num_sms() {   // of the CURRENT device (cached per device: a process may drive several)
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

// Bit position of kernel offset k inside the row key the tile rows are sorted by (sort.cu; also emitted by the kernel-map
// build): RARE offsets are the most significant bits, so rows that own a rare neighbour share tiles.  For 3x3x3 kernels the
// static order corners > outer-plane edges > outer face centres > middle-plane corners > middle-plane edges > centre matches
// the measured frequencies on LiDAR scans (7 / 12 / 18 / 36 / 47 / 100 %); other kernel sizes keep the natural order.
struct KeyBits {
  unsigned char pos[32];
};
inline KeyBits key_bits_for(int K) {
  KeyBits kb;
  for (int k = 0; k < 32; ++k) kb.pos[k] = (unsigned char)k;
  if (K == 27) {
    static const int order[27] = {0, 2, 6, 8, 18, 20, 24, 26, 1, 3, 5, 7, 19, 21, 23, 25, 4, 22, 9, 11, 15, 17, 10, 12, 14, 16, 13};
    for (int r = 0; r < 27; ++r) kb.pos[order[r]] = (unsigned char)(26 - r);  // order[0] is the most significant bit
  }
  return kb;
}

// grid for a grid-stride elementwise kernel: enough CTAs to fill the chip a few times, never more than needed
inline int grid_for(int64_t work_items, int threads, int ctas_per_sm = 8) {
  int64_t need = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Row count of a launch: the host value, or — on the sync-free pipeline, where data-dependent sizes never travel to the
// host — a device counter clamped to the capacity the buffers and the grid were sized for.
__device__ __forceinline__ int64_t dev_count(const int *n_dev, int64_t n_cap) {
  if (!n_dev) return n_cap;
  const int64_t v = *reinterpret_cast<const volatile int *>(n_dev);
  return v < n_cap ? (v > 0 ? v : 0) : n_cap;
}

// ------------------------------------------------------------------ coordinate keys
constexpr int COORD_BITS = 19;
constexpr int COORD_BIAS = 1 << 18;
constexpr unsigned long long EMPTY_KEY = ~0ull;

__host__ __device__ inline bool coord_in_range(int x, int y, int z, int b) {
  return (unsigned)(x + COORD_BIAS) < (1u << COORD_BITS) && (unsigned)(y + COORD_BIAS) < (1u << COORD_BITS) &&
         (unsigned)(z + COORD_BIAS) < (1u << COORD_BITS) && (unsigned)b < 128u;
}
// (b,x,y,z) -> 64-bit key whose unsigned order is the lexicographic order of (b,x,y,z)
__host__ __device__ inline unsigned long long pack_coord(int x, int y, int z, int b) {
  return ((unsigned long long)(unsigned)b << 57) | ((unsigned long long)(unsigned)(x + COORD_BIAS) << 38) |
         ((unsigned long long)(unsigned)(y + COORD_BIAS) << 19) | (unsigned long long)(unsigned)(z + COORD_BIAS);
}
__host__ __device__ inline int4 unpack_coord(unsigned long long k) {
  int4 c;
  c.w = (int)(k >> 57);
  c.x = (int)((k >> 38) & ((1u << COORD_BITS) - 1)) - COORD_BIAS;
  c.y = (int)((k >> 19) & ((1u << COORD_BITS) - 1)) - COORD_BIAS;
  c.z = (int)(k & ((1u << COORD_BITS) - 1)) - COORD_BIAS;
  return c;
}

// reference hash: TS/backend/hash/hash_cuda.cu:10-23
__host__ __device__ inline long long fnv60(int x, int y, int z, int b) {
  unsigned long long h = 14695981039346656037ULL;
  h ^= (unsigned int)x; h *= 1099511628211ULL;
  h ^= (unsigned int)y; h *= 1099511628211ULL;
  h ^= (unsigned int)z; h *= 1099511628211ULL;
  h ^= (unsigned int)b; h *= 1099511628211ULL;
  h = (h >> 60) ^ (h & 0xFFFFFFFFFFFFFFFULL);
  return (long long)h;
}

// ------------------------------------------------------------------ open-addressing table
struct __align__(16) Slot {
  unsigned long long key;
  int val;
  int pad;
};

__device__ inline unsigned long long mix64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

__device__ inline void table_insert(Slot *tab, unsigned long long mask, unsigned long long key, int val) {
  unsigned long long s = mix64(key) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&tab[s].key, EMPTY_KEY, key);
    if (prev == EMPTY_KEY || prev == key) {
      atomicMin(&tab[s].val, val);
      return;
    }
    s = (s + 1) & mask;
  }
}

// Home slot of a packed COORDINATE key: the line (8 slots = 128 B) is chosen by hashing everything but the three low z
// bits, the slot inside the line by those bits.  Voxels that are neighbours in z then sit in the same 128-byte line, so
// the three dz probes of a kernel map (and the probes of a warp's 32 consecutive, mostly z-adjacent rows) hit one or
// two L2 sectors instead of three random ones.  Any collision falls back to the usual linear probing.
__device__ __forceinline__ unsigned long long coord_home(unsigned long long key, unsigned long long mask) {
#ifdef TSG_COORD_LOCALITY
  return ((mix64(key >> 3) << 3) | (key & 7ull)) & mask;
#else
  return mix64(key) & mask;
#endif
}

__device__ inline void table_insert_coord(Slot *tab, unsigned long long mask, unsigned long long key, int val) {
  unsigned long long s = coord_home(key, mask);
  while (true) {
    unsigned long long prev = atomicCAS(&tab[s].key, EMPTY_KEY, key);
    if (prev == EMPTY_KEY || prev == key) {
      atomicMin(&tab[s].val, val);
      return;
    }
    s = (s + 1) & mask;
  }
}

__device__ inline int table_find_coord(const Slot *__restrict__ tab, unsigned long long mask, unsigned long long key) {
  unsigned long long s = coord_home(key, mask);
  while (true) {
    const ulonglong2 raw = __ldg(reinterpret_cast<const ulonglong2 *>(tab + s));
    if (raw.x == key) return (int)(unsigned)(raw.y & 0xffffffffull);
    if (raw.x == EMPTY_KEY) return -1;
    s = (s + 1) & mask;
  }
}

__device__ inline int table_find(const Slot *__restrict__ tab, unsigned long long mask, unsigned long long key) {
  unsigned long long s = mix64(key) & mask;
  while (true) {
    const ulonglong2 raw = __ldg(reinterpret_cast<const ulonglong2 *>(tab + s));
    if (raw.x == key) return (int)(unsigned)(raw.y & 0xffffffffull);
    if (raw.x == EMPTY_KEY) return -1;
    s = (s + 1) & mask;
  }
}

// stable LSD radix sort of (64-bit key, 32-bit payload) pairs over key bits [begin_bit, end_bit) (sort.cu); vals_in NULL =
// the payload is the input position; workspace tsg_sort_ws_bytes(n); n_dev: optional device row counter (n = capacity)
int sort_pairs(const unsigned long long *keys_in, const unsigned *vals_in, int64_t n, int begin_bit, int end_bit,
               unsigned long long *keys_out, unsigned *vals_out, void *ws_mem, size_t ws_bytes, cudaStream_t stream,
               const int *n_dev = nullptr);

// ------------------------------------------------------------------ block scan (power-of-two block sizes, <=1024)
template <int THREADS>
__device__ inline int block_exclusive_scan(int v, int *total) {
  __shared__ int warp_sums[32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < THREADS / 32 ? warp_sums[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < THREADS / 32) warp_sums[lane] = winc - w;
    if (lane == 31) block_total = winc;
  }
  __syncthreads();
  int res = inc - v + warp_sums[warp];
  if (total) *total = block_total;
  __syncthreads();
  return res;
}

// ------------------------------------------------------------------ dtype helpers
template <typename T> __device__ inline float to_f32(T v);
template <> __device__ inline float to_f32<float>(float v) { return v; }
template <> __device__ inline float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ inline float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ inline T from_f32(float v);
template <> __device__ inline float from_f32<float>(float v) { return v; }
template <> __device__ inline __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ inline __half from_f32<__half>(float v) { return __float2half_rn(v); }

}  // namespace tsg
