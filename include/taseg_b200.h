/*
 * taseg_b200 — C ABI of the B200-native (sm_100a) sparse-convolution hot path of LittlePey/TASeg.
 *
 * Drop-in boundary.  The reference reaches its native code through a pybind11 module
 * `torchsparse.backend` (TS/torchsparse/backend/pybind_cuda.cpp:18-39, TS = torchsparse/ inside
 * /root/reference/package/torchsparse.zip) whose functions take at::Tensor by value.  This library
 * replaces that module with plain C entry points: raw DEVICE pointers, explicit sizes, a dtype enum
 * and the CUDA stream; every function returns a tsg_status.  No torch type appears here.
 *
 * Ownership.  The library never allocates or frees device memory: outputs and workspaces are owned by
 * the caller (the Python host uses torch's caching allocator).  `*_ws_bytes` functions size workspaces.
 * All work is enqueued on `stream`; no function synchronises the device or the host.
 * Data-dependent sizes (number of unique voxels, map pairs) are written to caller-provided device
 * counters so the host decides when to read them.
 *
 * Coordinates are int32 rows [x, y, z, b] (TS/torchsparse/utils/collate.py:28-33).
 * Exact-coordinate tables require -2^18 <= x,y,z < 2^18 and 0 <= b < 128 (packed 64-bit keys);
 * a violation raises bit 0 of the caller's `status` word on the device (TSG_ERR_RANGE when read back).
 */
#ifndef TASEG_B200_H
#define TASEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *tsg_stream_t;

typedef enum {
  TSG_OK = 0,
  TSG_ERR_INVALID = 1,     /* shape/argument mismatch: the reference throws std::invalid_argument
                              (convolution_cuda.cu:57-59) -> Python ValueError */
  TSG_ERR_CUDA = 2,        /* a CUDA runtime call failed; tsg_last_error() has the string */
  TSG_ERR_WORKSPACE = 3,   /* workspace too small */
  TSG_ERR_RANGE = 4,       /* coordinate outside the packable range */
  TSG_ERR_UNSUPPORTED = 5  /* shape not supported by this kernel (e.g. channels not padded) */
} tsg_status;

typedef enum { TSG_F32 = 0, TSG_BF16 = 1, TSG_F16 = 2 } tsg_dtype;

/* epilogue flags for the convolution entry points */
enum { TSG_EPI_BIAS = 1, TSG_EPI_RELU = 2, TSG_EPI_RESIDUAL = 4, TSG_EPI_ACCUMULATE = 8 };

int tsg_version(void);
const char *tsg_last_error(void); /* thread-local, valid until the next call on this thread */

/* ---------------------------------------------------------------- hashing
 * tsg_hash        replaces hash_cuda(idx)                    TS/backend/hash/hash_cuda.h:5,  hash_cuda.cu:10-23,67-73
 * tsg_kernel_hash replaces kernel_hash_cuda(idx, offsets)    hash_cuda.h:6-7, hash_cuda.cu:27-55,75-84
 * 64-bit FNV-1a over the four 32-bit words, folded to 60 bits; out is int64 (K x N for the kernel form,
 * every row keeps its own batch index).  `offsets` is a DEVICE (K,3) int32 array. */
int tsg_hash(const int32_t *coords, int64_t n, int64_t *out, tsg_stream_t stream);
int tsg_kernel_hash(const int32_t *coords, int64_t n, const int32_t *offsets, int k, int64_t *out,
                    tsg_stream_t stream);

/* ---------------------------------------------------------------- key -> index tables (a9)
 * Open-addressing table of 16-byte slots {uint64 key, int32 value}; `slots` must be a power of two
 * >= 2n (tsg_table_slots).  Replaces CuckooHashTableCuda_Multi (TS/backend/hashmap/hashmap_cuda.cu:139-212)
 * and hash_query_cuda(hash_query, hash_target, idx_target) (TS/backend/others/query_cuda.h:5-7,
 * query_cuda.cu:9-56): value = position of the key in `keys`; on duplicate keys the SMALLEST position wins
 * (the CPU reference's insert-first behaviour, query_cpu.cpp:22-26).  The table is built once and reused
 * for every query against the same key set (the reference rebuilds it on every call).
 * tsg_table_query writes position or -1 (the "-1 = miss" the Python wrapper produces, TS/nn/functional/query.py:32). */
int64_t tsg_table_slots(int64_t n);
int tsg_table_build(const int64_t *keys, int64_t n, void *table, int64_t slots, tsg_stream_t stream);
int tsg_table_query(const void *table, int64_t slots, const int64_t *queries, int64_t nq, int64_t *out,
                    tsg_stream_t stream);

/* Exact-coordinate table: keys are the packed (b,x,y,z) themselves, not their hash (no collisions). */
int tsg_coord_table_build(const int32_t *coords, int64_t n, void *table, int64_t slots, int32_t *status,
                          tsg_stream_t stream);

/* ---------------------------------------------------------------- kernel maps (a15, a17; K2+K4+nonzero+sum fused)
 * For every output voxel o and kernel offset k, nbr[k*n_out + o] = row of in-voxel at out_coords[o]+offsets[k],
 * or -1.  `offsets` is a HOST (K,3) int32 array in the reference's weight order
 * (get_kernel_offsets, TS/nn/utils/kernel.py:11-32), already scaled by tensor stride and dilation; K <= 32.
 * nbsizes[k] (device int32, K) = number of hits for offset k (TS/nn/functional/conv.py:168).
 * blockcnt (device int32, K * tsg_kmap_blocks(n_out)) are per-1024-row hit counts, consumed by tsg_kmap_pairs. */
int64_t tsg_kmap_blocks(int64_t n_out);
int tsg_kmap_build(const void *table, int64_t slots, const int32_t *out_coords, int64_t n_out,
                   const int32_t *offsets_host, int k, int32_t *nbr, int32_t *nbsizes, int32_t *blockcnt,
                   tsg_stream_t stream);
/* Reference-format map: nbmaps (P,2) int64 rows [in_idx, out_idx], grouped by k, ascending out_idx
 * (TS/nn/functional/conv.py:169-172).  P = sum(nbsizes) must be known by the caller (it sized nbmaps). */
int tsg_kmap_pairs(const int32_t *nbr, int k, int64_t n_out, int32_t *blockcnt, int64_t *nbmaps,
                   tsg_stream_t stream);
/* Transposed table for dgrad / transposed convolution: nbr_t[k*n_in + i] = o where nbr[k*n_out+o] == i, else -1. */
int tsg_kmap_transpose(const int32_t *nbr, int k, int64_t n_out, int64_t n_in, int32_t *nbr_t,
                       tsg_stream_t stream);
/* Reference-format pairs -> output-stationary table (for callers that hold torchsparse kmaps):
 * nbmaps (P,2) int32 [in,out], nbsizes HOST (K) int32 (the reference passes nbsizes.cpu(), conv.py:56). */
int tsg_kmap_from_pairs(const int32_t *nbmaps, const int32_t *nbsizes_host, int k, int transposed,
                        int64_t n_rows_out, int32_t *nbr, tsg_stream_t stream);

/* ---------------------------------------------------------------- sort / unique (a5, a12, a16)
 * Stable LSD radix sort (9-bit digits, one-sweep with decoupled look-back) of 64-bit keys with a 32-bit payload
 * (payload = 0..n-1 when vals_in is NULL); n < 2^30.
 * Results land in keys_out/vals_out.  ws: tsg_sort_ws_bytes(n). */
size_t tsg_sort_ws_bytes(int64_t n);
int tsg_sort_pairs(const uint64_t *keys_in, const uint32_t *vals_in, int64_t n, int begin_bit, int end_bit,
                   uint64_t *keys_out, uint32_t *vals_out, void *ws, size_t ws_bytes, tsg_stream_t stream);

/* Unique voxels of a coordinate list, ordered lexicographically by (b, x, y, z):
 *   sparse_quantize + sparse_collate   TS/utils/quantize.py:24-46, TS/utils/collate.py:11-37
 *   spdownsample (after truncation)    TS/nn/functional/downsample.py:47-52
 * in_coords (N,4) [x,y,z,b]; if trunc_stride > 0 every x,y,z is first truncated toward zero to a multiple of it
 * (downsample.py:27-29).  Outputs (any may be NULL): out_coords (<=N,4), first_idx (<=N) = first occurrence in
 * input order, inverse (N) = voxel row of every input point; *m_dev = number of unique voxels.
 * field_bits_host: NULL, or HOST int32[4] = bit widths of x,y,z,b when the caller knows 0 <= coordinate < 2^bits
 * (true after the loader's min-shift): sort keys are then packed densely and the radix sort runs
 * ceil(sum/9) passes instead of 8; a coordinate outside the promise raises bit 0 of *status. */
size_t tsg_unique_ws_bytes(int64_t n);
int tsg_unique_coords(const int32_t *in_coords, int64_t n, int trunc_stride, const int32_t *field_bits_host,
                      int32_t *out_coords, int32_t *first_idx, int32_t *inverse, int32_t *m_dev, int32_t *status,
                      void *ws, size_t ws_bytes, tsg_stream_t stream);
/* Unique by ascending 60-bit FNV hash — the voxel order of initial_voxelize
 * (R/pcseg/model/segmentor/voxel/minkunet/utils.py:17-18: torch.unique(pc_hash)). */
int tsg_unique_hash(const int32_t *in_coords, int64_t n, int32_t *out_coords, int32_t *first_idx,
                    int32_t *inverse, int32_t *m_dev, void *ws, size_t ws_bytes, tsg_stream_t stream);

/* ---------------------------------------------------------------- multi-frame aggregation + quantisation (a1-a4)
 * tsg_fuse_multi_scan replaces SemantickittiMsDataset.fuse_multi_scan
 * (R/pcseg/data/dataset/semantickitti/semantickitti_ms.py:403-417): p' = ((P.[p;1])_xyz - t0) . R0 in fp32,
 * products rounded before left-to-right sums (no FMA) — bit exact with numpy.  pose0/pose: HOST float[16]
 * row-major 4x4.  pts/out (N,c) fp32, c >= 3, extra columns copied. */
int tsg_fuse_multi_scan(const float *pts, int64_t n, int c, const float *pose0, const float *pose, float *out,
                        tsg_stream_t stream);
/* nuScenes warp (R/pcseg/data/dataset/nuscenes/nuscenes_ms.py:371): xyz = fp32(fp64(xyz) @ R + T); R (3,3), T (3) HOST double. */
int tsg_transform_point(const float *pts, int64_t n, int c, const double *R, const double *T, float *out,
                        tsg_stream_t stream);

/* Whole-batch aggregation front end (semantickitti_ms.py:140-149,253-257,263-320 + semantickitti_voxel_ms.py:121-151).
 * `frames` describes F frames laid out back to back in `pts` (sum n, c_in) fp32:
 *   sample  batch index b of the frame
 *   is_cur  1 for the current scan of its sample (listed first within the sample), 0 for history
 *   pose0/pose  float[16] each (ignored for is_cur)
 * For every point: warp (history only), append the time flag column (1 current / 0 history) after column 3,
 * apply the optional keep mask (FSA, `keep` (sum n) uint8 or NULL), clamp history to the current scan's min corner,
 * quantise round-half-even(xyz / voxel) in fp32, shift by the per-sample min over kept points.
 * Outputs: feats (sum n, c_in+1), coords (sum n,4) [x,y,z,b], flags (sum n) uint8 = kept.
 * ws: tsg_aggregate_ws_bytes(n_samples); on return its first n_samples records of 12 x 4 bytes hold, per sample,
 * float cur_min[4], int32 ms_min[4], int32 ms_max[4] (quantised extent before the shift; the caller may read it to
 * derive field_bits for tsg_unique_coords). */
typedef struct {
  int64_t offset, count;
  int32_t sample, is_cur;
  float pose0[16], pose[16];
} tsg_frame;
size_t tsg_aggregate_ws_bytes(int n_samples);
int tsg_aggregate_quantize(const float *pts, int c_in, const tsg_frame *frames_host, int n_frames, int n_samples,
                           const uint8_t *keep, float voxel_size, float *feats, int32_t *coords, uint8_t *flags,
                           void *ws, size_t ws_bytes, tsg_stream_t stream);
/* nuScenes multi-sweep aggregation (BASELINE configs[3]) in the same three passes:
 *   R/pcseg/data/dataset/nuscenes/nuscenes_ms.py:284-341  per sweep: drop the ego box |x| < 1 & |y| < 1.5 (tested on the
 *                                                          RAW points), time lag into column 4, warp into the key frame
 *   R/pcseg/data/dataset/nuscenes/nuscenes_ms.py:348-373  transform_point: float64 p @ R + T, stored back as float32
 *   R/pcseg/data/dataset/nuscenes/nuscenes_voxel_ms.py:122-160  clamp to the key sweep's min corner, round, min-shift
 * pts (sum count, c >= 5) rows [x, y, z, intensity, .]; sweeps_host: n_sweeps records (key sweep first within a sample).
 * Outputs as tsg_aggregate_quantize: feats (sum count, c) with column 4 = dt, coords [x,y,z,sample], flags = kept. */
typedef struct tsg_sweep {
  int64_t offset, count;
  int32_t sample, is_key;
  double R[9], T[3];
  float dt, pad_;
} tsg_sweep;
size_t tsg_aggregate_nus_ws_bytes(int n_samples);
int tsg_aggregate_quantize_nus(const float *pts, int c, const tsg_sweep *sweeps_host, int n_sweeps, int n_samples,
                               float voxel_size, float *feats, int32_t *coords, uint8_t *flags, void *ws, size_t ws_bytes,
                               tsg_stream_t stream);
/* Stable stream compaction of rows by a uint8 flag: used after tsg_aggregate_quantize.
 * rows_a (n, wa) and rows_b (n, wb) are 4-byte-element rows (either may be NULL); *m_dev = rows kept;
 * pos (n) int32 = destination row or -1.  ws: tsg_compact_ws_bytes(n). */
size_t tsg_compact_ws_bytes(int64_t n);
int tsg_compact_rows(const uint8_t *flags, int64_t n, const void *rows_a, int wa, void *out_a, const void *rows_b,
                     int wb, void *out_b, int32_t *pos, int32_t *m_dev, void *ws, size_t ws_bytes,
                     tsg_stream_t stream);
/* out[i, :] = src[idx[i], :] for 4-byte-element rows (feature pick of voxel representatives, logits to points). */
int tsg_gather_rows(const void *src, int width, const int32_t *idx, int64_t n, void *out, tsg_stream_t stream);

/* ---------------------------------------------------------------- point <-> voxel (a10-a14)
 * tsg_count replaces count_cuda(idx, s)                                   TS/backend/others/count_cuda.h:5
 * tsg_voxelize_fwd/bwd replace voxelize_forward_cuda / voxelize_backward_cuda   TS/backend/voxelize/voxelize_cuda.h:5-10
 * tsg_devoxelize_fwd/bwd replace devoxelize_forward_cuda / _backward_cuda        TS/backend/devoxelize/devoxelize_cuda.h:5-11
 * feats are (rows, c) of `dtype`; accumulation is fp32. */
int tsg_count(const int32_t *idx, int64_t n, int32_t *out, int64_t m, tsg_stream_t stream);
/* B200-first versions of the same four transforms (taseg_b200/csrc/pointvoxel.cu): 16-byte vector access with a lane
 * group per row, index / weight words read once per row, and — for the scatter-mean of
 * TS/backend/voxelize/voxelize_cuda.cu:12-25 — a deterministic SEGMENTED REDUCTION instead of fp32 atomics:
 *   tsg_voxelize_plan     sorts the point ids by voxel (stable radix sort): order (n) uint32 = point ids grouped by voxel
 *                         in point order, skeys (n) uint64 = their voxel ids (scratch), seg (m, 2) int32 = [first, one past
 *                         last] sorted position of every voxel ({0, 0}: empty).  One plan serves every voxelize over `idx`.
 *   tsg_voxelize_fwd_seg  out[v] = sum_{i in voxel v, point order} feats[i] / counts[v], written once, any dtype, no
 *                         accumulation workspace; run-to-run bit-identical.
 *   tsg_voxelize_bwd_vec / tsg_devoxelize_fwd_vec / tsg_devoxelize_bwd_vec: vectorised gather / trilinear gather / vector
 *                         atomic scatter (TS/backend/voxelize/voxelize_cuda.cu:28-42, devoxelize_cuda.cu:11-57).
 * All need rows of a multiple of 16 bytes, 16-byte aligned (tsg_pv_vector_ok); the scalar entry points above remain. */
int tsg_pv_vector_ok(const void *a, const void *b, int c, int dtype);
size_t tsg_voxelize_plan_ws_bytes(int64_t n);
int tsg_voxelize_plan(const int32_t *idx, int64_t n, int64_t m, uint32_t *order, uint64_t *skeys, int32_t *seg,
                      void *ws, size_t ws_bytes, tsg_stream_t stream);
int tsg_voxelize_fwd_seg(const void *feats, int dtype, const uint32_t *order, const int32_t *seg, const int32_t *counts,
                         int64_t n, int c, int64_t m, void *out, tsg_stream_t stream);
int tsg_voxelize_bwd_vec(const void *top_grad, int dtype, const int32_t *idx, const int32_t *counts, int64_t n, int c,
                         void *bottom_grad, tsg_stream_t stream);
int tsg_devoxelize_fwd_vec(const void *feats, int dtype, const int32_t *idx8, const float *w8, int64_t n, int c, void *out,
                           tsg_stream_t stream);
int tsg_devoxelize_bwd_vec(const void *top_grad, int dtype, const int32_t *idx8, const float *w8, int64_t n, int c,
                           int64_t m, void *bottom_grad, float *acc_ws, tsg_stream_t stream);
int tsg_voxelize_fwd(const void *feats, int dtype, const int32_t *idx, const int32_t *counts, int64_t n, int c,
                     int64_t m, void *out, float *acc_ws, tsg_stream_t stream);
int tsg_voxelize_bwd(const void *top_grad, int dtype, const int32_t *idx, const int32_t *counts, int64_t n, int c,
                     void *bottom_grad, tsg_stream_t stream);
int tsg_devoxelize_fwd(const void *feats, int dtype, const int32_t *idx8, const float *w8, int64_t n, int c,
                       void *out, tsg_stream_t stream);
int tsg_devoxelize_bwd(const void *top_grad, int dtype, const int32_t *idx8, const float *w8, int64_t n, int c,
                       int64_t m, void *bottom_grad, float *acc_ws, tsg_stream_t stream);
/* Fused point->voxel lookups on an exact-coordinate table built over the voxels of tensor stride s:
 *   tsg_point_query      idx[i] = row of voxel floor(p/s)*s                 (point_to_voxel, minkunet/utils.py:44-51)
 *   tsg_trilinear_query  idx8/w8 = 8 corner rows + calc_ti_weights          (voxel_to_point, minkunet/utils.py:72-85;
 *                                                                            TS/nn/functional/devoxelize.py:10-48)
 * pcoords (N,4) fp32 [x,y,z,b].  Corner order = get_kernel_offsets(2, s): z fastest. */
int tsg_point_query(const void *table, int64_t slots, const float *pcoords, int64_t n, int stride, int32_t *idx,
                    tsg_stream_t stream);
int tsg_trilinear_query(const void *table, int64_t slots, const float *pcoords, int64_t n, int stride, int nearest,
                        int32_t *idx8, float *w8, tsg_stream_t stream);
/* Fused multi-scale devoxelisation: out[j, :c_out] = sum_s sum_k w_{s,k}(p) * feats_s[idx_{s,k}(p), :c_out] with
 * p = pcoords[rows[j]] (rows == NULL: p = pcoords[j]) — tsg_trilinear_query + tsg_devoxelize_fwd of up to four scales,
 * their sum and the final row gather in one pass (voxel_to_point x3 + classifier tail of minkunet_ms.py:392-420 once the
 * Linear has been applied per scale at voxel level).  tables/slots/strides/feats: HOST arrays of n_scales entries
 * (device tables from tsg_coord_table_build, fp32 device features (n_s, c), c % 4 == 0); out (m, c_out) fp32. */
int tsg_devoxelize_multi(int n_scales, const void *const *tables, const int64_t *slots, const int32_t *strides,
                         const float *const *feats, int c, const float *pcoords, const int32_t *rows, int64_t m,
                         float *out, int c_out, tsg_stream_t stream);
/* initial_voxelize front end (minkunet/utils.py:11-16): fc = (C*init_res)/after_res in fp32 (mul then div),
 * out_f (N,4) fp32 = [fc, b], out_i (N,4) int32 = floor. */
int tsg_rescale_coords(const float *pcoords, int64_t n, float init_res, float after_res, float *out_f,
                       int32_t *out_i, tsg_stream_t stream);

/* ---------------------------------------------------------------- sparse convolution (a17-a19)
 * tsg_conv_fwd replaces convolution_forward_cuda(in_feat, out_feat, kernel, neighbor_map, neighbor_offset, transpose)
 * (TS/backend/convolution/convolution_cuda.h:5-7, convolution_cuda.cu:53-165):
 *     out[o,:] = sum_k in[nbr[k,o],:] @ W[k]            (nbr < 0 contributes nothing)
 * as ONE output-stationary implicit-GEMM launch: no gather/scatter buffers, no per-offset GEMMs, no host sync.
 * in (n_in, c_in), weight (K, c_in, c_out), out (n_out, c_out): all `dtype`, fp32 accumulate.
 * Epilogue (eval-mode fusion): out = relu?( acc*scale[c] + bias[c] + residual[o,c] );
 * scale/bias fp32 (c_out) or NULL, residual same dtype as out or NULL.
 * Returns TSG_ERR_INVALID when c_in does not match the kernel ("Input feature size and kernel size mismatch"). */
int tsg_conv_fwd(const void *in, int dtype, int64_t n_in, int c_in, const void *weight, int k, int c_in_w,
                 int c_out, const int32_t *nbr, int64_t n_out, void *out, const float *scale, const float *bias,
                 const void *residual, int relu, tsg_stream_t stream);
/* Backward (convolution_backward_cuda, convolution_cuda.h:9-13, convolution_cuda.cu:167-278):
 *   dgrad: grad_in[i,:]  = sum_k grad_out[nbr_t[k,i],:] @ W[k]^T      (same kernel, transposed table + weights)
 *   wgrad: grad_w[k]     = sum_o in[nbr[k,o],:]^T @ grad_out[o,:]
 * fp32 only (training master precision). */
int tsg_conv_dgrad(const float *grad_out, int64_t n_out, int c_out, const float *weight, int k, int c_in,
                   const int32_t *nbr_t, int64_t n_in, float *grad_in, tsg_stream_t stream);
int tsg_conv_wgrad(const float *in, int64_t n_in, int c_in, const float *grad_out, int64_t n_out, int c_out,
                   const int32_t *nbr, int k, float *grad_w, tsg_stream_t stream);
/* The same weight gradient for the autocast (bf16) training path: bf16 `in` / `grad_out` rows (c_in, c_out multiples of
 * 8), fp32 accumulation and fp32 grad_w.  Rows that have a neighbour at offset k are compacted into a pair list inside the
 * kernel, so only the P = sum_k nbsizes[k] real pairs are multiplied (warp-level tensor-core MMAs). */
int tsg_conv_wgrad_bf16(const void *in, int64_t n_in, int c_in, const void *grad_out, int64_t n_out, int c_out,
                        const int32_t *nbr, int k, float *grad_w, tsg_stream_t stream);
/* Compact pair lists of a neighbour table for the weight gradient: pairs (capacity K * n_rows, int32 x 2) receives, offset
 * by offset, the {in row = nbr[k, o], out row = o} of every hit in ascending o — the order of the reference's nbmaps
 * (TS/nn/functional/conv.py:169-172) in 32-bit — and start (K + 1 int32, device) the first pair of every offset
 * (start[K] = P).  Built once per kernel map, shared by every layer that uses the map.  Nothing returns to the host. */
size_t tsg_kmap_pair_list_ws_bytes(int k, int64_t n_rows);
int tsg_kmap_pair_list(const int32_t *nbr, int k, int64_t n_rows, int64_t nbr_stride, int32_t *pairs, int32_t *start, void *ws,
                       size_t ws_bytes, tsg_stream_t stream);
/* ... and on the 5th-generation tensor cores (tcgen05.mma, fp32 accumulators in TMEM): the gathered rows — 64 channels
 * = 128 bytes per pair, eight pairs per swizzle atom — are MN-major operand tiles as they land in shared memory, so
 * grad_w[k] = X_k^T gY_k needs no transposition; the pair list is the GEMM's K dimension (taseg_b200/csrc/conv_wgrad_tc.cu).
 * Replaces the per-offset gather + torch::mm_out of TS/backend/convolution/convolution_cuda.cu:167-278 under autocast.
 * c_in, c_out multiples of 8, c_out <= 256.  pairs / start: tsg_kmap_pair_list of the table whose rows are the forward's
 * outputs; pair_cap = capacity of `pairs` (sizes the grid: the true count stays on the device). */
int tsg_conv_wgrad_tc(const void *in, int64_t n_in, int c_in, const void *grad_out, int64_t n_out, int c_out,
                      const int32_t *pairs, const int32_t *start, int64_t pair_cap, int k, float *grad_w, tsg_stream_t stream);

/* ---------------------------------------------------------------- BatchNorm over feature rows (a21)
 * spnn.BatchNorm is nn.BatchNorm1d applied to SparseTensor.F (TS/torchsparse/nn/modules/norm.py:10-13; 63 of them in the
 * benchmark network, one after every convolution).  x, y, dy, dx: (n, c) rows, fp32 or bf16, c a multiple of 8 <= 1024;
 * statistics, affine parameters and their gradients fp32.  ws: tsg_bn_ws_bytes(n, c) bytes.  Deterministic (per-CTA
 * partials combined in double precision in a fixed order; no floating-point atomics on global memory).
 *   tsg_bn_stats:    batch mean / invstd = 1 / sqrt(biased var + eps); running statistics updated in place when given
 *                    (momentum, unbiased variance — nn.BatchNorm1d's rule)
 *   tsg_bn_apply:    y = (x - mean) invstd gamma + beta   (evaluation: mean / invstd derived from the running statistics);
 *                    relu != 0 fuses the ReLU that follows every BatchNorm of a convolution block (blocks: conv, BN, ReLU —
 *                    R/pcseg/model/segmentor/voxel/minkunet/minkunet.py:23-60) into the same pass, forward and backward
 *   tsg_bn_backward: sums = {dbeta, dgamma}; dx = gamma invstd (dy - dbeta / n - xhat dgamma / n)  (training) or
 *                    gamma invstd dy (evaluation) */
size_t tsg_bn_ws_bytes(int64_t n, int c);
int tsg_bn_stats(const void *x, int dtype, int64_t n, int c, float eps, float momentum, float *running_mean, float *running_var,
                 float *mean, float *invstd, void *ws, size_t ws_bytes, tsg_stream_t stream);
/* tsg_bn_stats + nn.BatchNorm1d's `num_batches_tracked += 1` (int64 device scalar, may be NULL) in the same launch */
int tsg_bn_stats2(const void *x, int dtype, int64_t n, int c, float eps, float momentum, float *running_mean, float *running_var,
                  int64_t *num_batches_tracked, float *mean, float *invstd, void *ws, size_t ws_bytes, tsg_stream_t stream);
int tsg_bn_apply(const void *x, int dtype, int64_t n, int c, const float *mean, const float *invstd, const float *gamma,
                 const float *beta, int relu, void *y, tsg_stream_t stream);
int tsg_bn_backward(const void *x, const void *dy, int dtype, int64_t n, int c, const float *mean, const float *invstd,
                    const float *gamma, const float *beta, int relu, int training, float *sums, void *dx, void *ws, size_t ws_bytes,
                    tsg_stream_t stream);

/* tcgen05/TMEM path (bf16 operands, fp32 accumulate in tensor memory).
 * Weights are packed once per layer into the shared-memory image the MMA consumes (128B-swizzled K-major
 * [c_out][64] blocks per kernel offset and 64-channel slice; single-source layers with c0 of 16 or 32 put 64/c0
 * kernel offsets side by side in one block): tsg_conv_pack_bytes / tsg_conv_pack_weights.
 * The A operand may come from two feature tensors (channel concat of a decoder feature and its encoder skip,
 * TS/operators.py:10-17, without materialising the concat): in0 (n_in, c0) then in1 (n_in, c1); c1 may be 0.
 * c0, c1, c_out multiples of 16; c_out <= 256.  tile_mask (ceil(n_out/128)) uint32 from tsg_kmap_tile_mask.
 * nbr == NULL means the identity map (row o reads row o for every k: 1x1x1 convolutions and point MLPs) and
 * tile_mask == NULL means every offset is active.
 * perm == NULL: tile row r is output row r.  Otherwise nbr / tile_mask describe tile rows in the order produced by
 * tsg_kmap_sort_rows and row r is written to out[perm[r]] (residual read from residual[perm[r]]): rows with the same
 * neighbour pattern share a tile, so most (tile, offset) pairs are empty and skipped.
 * nbr_stride: elements between consecutive offsets of nbr (>= n_out).  The producers fetch eight neighbour indices per
 * thread with two 16-byte loads and never test row bounds, so nbr_stride must be a multiple of 256 with -1 in the padding
 * rows [n_out, nbr_stride) — what tsg_kmap_sort_rows writes for out_stride = tsg_kmap_sort_stride(n_out).
 * sched: NULL (tiles are dealt to the CTAs round-robin) or two int32 that are ZERO on entry: the kernel hands tiles out
 * dynamically through them, heaviest first, and leaves them zero again; launches that may overlap need their own pair.
 * out dtype TSG_BF16 or TSG_F32. */
size_t tsg_conv_pack_bytes(int k, int c0, int c1, int c_out);
int tsg_conv_pack_weights(const float *weight, int k, int c_in, int c_out, int c0, int c1,
                          const float *out_scale, void *packed, tsg_stream_t stream);
/* ... from a strided (K, c_in, c_out) view (element strides): W[k]^T for the data gradient and column blocks of wide layers are
 * packed straight from the parameter, without a transposed / sliced copy (TS/backend/convolution/convolution_cuda.cu:229-246
 * multiplies by the transposed kernel through cuBLAS flags) */
int tsg_conv_pack_weights2(const float *weight, int k, int c_in, int c_out, int64_t stride_k, int64_t stride_cin,
                           int64_t stride_cout, int c0, int c1, const float *out_scale, void *packed, tsg_stream_t stream);
int tsg_kmap_tile_mask(const int32_t *nbr, int k, int64_t n_out, uint32_t *tile_mask, tsg_stream_t stream);
int tsg_conv_fwd_tc(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                    int c_out, const int32_t *nbr, int64_t nbr_stride, const uint32_t *tile_mask, const int32_t *perm, int64_t n_out,
                    void *out, int out_dtype, const float *bias, const void *residual, int relu, int num_sms,
                    int32_t *sched, tsg_stream_t stream);
/* Same convolution with an optional SECOND K PHASE accumulated into the same output tile before the epilogue: the
 * 1x1x1 shortcut convolution of a residual block (R/pcseg/model/segmentor/voxel/minkunet/minkunet.py:83-129:
 * relu(net(x) + downsample(x)); downsample = spnn.Conv3d(inc, outc, 1) + BN), i.e.
 *   out[o] = epilogue( sum_k in[nbr[k,o]] @ W[k]  +  sc_in[o] @ W_sc ),
 * which removes the shortcut launch, its (n_out, c_out) output and the residual read of the block's last convolution.
 * sc_in0/sc_in1 (n_out, sc_c0 / sc_c1) bf16 are the block's input (two tensors when it is a decoder concat), sc_packed_w =
 * tsg_conv_pack_weights(k = 1, sc_c0, sc_c1, c_out) with the shortcut's folded BN scale; bias must hold the SUM of both
 * folded BN shifts.  sc_idx: index line of the identity in TILE-ROW order, nbr_stride entries with -1 in the padding —
 * for a 3x3x3 stride-1 map this is the centre offset's line of the (sorted) table, nbr + (K/2) * nbr_stride; NULL = tile
 * row r reads row r (only valid without perm).  sc_in0 == NULL: no second phase (== tsg_conv_fwd_tc).
 * Launches with fewer tiles than SMs and c_out >= 128 are split into (tile, column half) work items (bit-identical
 * results: every output element still sees the same MMAs in the same order). */
int tsg_conv_fwd_tc2(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                     int c_out, const int32_t *nbr, int64_t nbr_stride, const uint32_t *tile_mask, const int32_t *perm,
                     int64_t n_out, const void *sc_in0, int sc_c0, const void *sc_in1, int sc_c1,
                     const void *sc_packed_w, const int32_t *sc_idx, void *out, int out_dtype, const float *bias,
                     const void *residual, int relu, int num_sms, int32_t *sched, tsg_stream_t stream);
/* ---- sync-free variants (graph-capturable inference pipeline, taseg_b200/pipeline.py).
 * The reference reads every data-dependent size back to the host (`torch.unique`, `nonzero`, `.item()` in
 * TS/nn/functional/{conv,downsample}.py, TS/utils/quantize.py).  These entry points take the row count from a DEVICE
 * counter instead: buffers and grids are sized by a host-side capacity `n_cap`, the kernels process
 * min(*n_dev, n_cap) rows, producers of counters clamp them to the consumer's capacity and set status bit 1 (value 2)
 * on overflow.  Nothing here allocates, copies from pageable host memory or synchronises, so a whole forward can be
 * captured into one CUDA graph and replayed. */
int tsg_conv_fwd_tc3(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                     int c_out, const int32_t *nbr, int64_t nbr_stride, const uint32_t *tile_mask, const int32_t *perm,
                     int64_t n_out_cap, const int32_t *n_out_dev, const void *sc_in0, int sc_c0, const void *sc_in1,
                     int sc_c1, const void *sc_packed_w, const int32_t *sc_idx, void *out, int out_dtype,
                     const float *bias, const void *residual, int relu, int num_sms, int32_t *sched, tsg_stream_t stream);
/* K-split work items for launches with about as many tiles as SMs (the stride-16 level of the benchmark: 120 tiles).
 * The reference has no counterpart: its 27 x (gather, cuBLAS GEMM, scatter) of TS/backend/convolution/convolution_cuda.cu:53-165
 * is balanced by cuBLAS; an output-stationary kernel must balance (tile, offset) work itself.  tsg_conv_split_items turns the
 * tile masks of a (mask-sorted) kernel map into a list of work items {tile, offset mask, part | parts << 8, slot} (4 int32
 * each, heaviest first; capacity ceil(n_out_cap / 128) + max_slots (max_parts - 1) items): a tile with more than `cap` active
 * offsets is summed by ceil(active / cap) <= max_parts items over disjoint runs of its offsets.  tsg_conv_fwd_tc4 with `items`
 * runs them on different SMs; every item but the last to arrive parks its fp32 accumulators in its slab of split_scratch
 * (max_slots x max_parts x 128 x c_out floats), the last one sums the parts in part order before bias / residual / ReLU, so
 * the result does not depend on the arrival order.  split_state: max_slots x 16 int32, zero before the first launch (the
 * kernel re-arms it).  items == NULL: tsg_conv_fwd_tc3. */
int tsg_conv_split_items(const uint32_t *tile_mask, int64_t n_out_cap, const int32_t *n_out_dev, int k, int cap, int max_parts,
                         int max_slots, int32_t *items, int32_t *n_items, tsg_stream_t stream);
int tsg_conv_fwd_tc4(const void *in0, int c0, const void *in1, int c1, int64_t n_in, const void *packed_w, int k,
                     int c_out, const int32_t *nbr, int64_t nbr_stride, const uint32_t *tile_mask, const int32_t *perm,
                     int64_t n_out_cap, const int32_t *n_out_dev, const void *sc_in0, int sc_c0, const void *sc_in1,
                     int sc_c1, const void *sc_packed_w, const int32_t *sc_idx, void *out, int out_dtype,
                     const float *bias, const void *residual, int relu, int num_sms, int32_t *sched, const int32_t *items,
                     const int32_t *n_items, int max_slots, int max_parts, float *split_scratch, int32_t *split_state,
                     tsg_stream_t stream);
int tsg_unique_coords_dev(const int32_t *in_coords, int64_t n_cap, const int32_t *n_dev, int trunc_stride,
                          const int32_t *field_bits_host, int32_t *out_coords, int64_t out_cap, int32_t *first_idx,
                          int32_t *inverse, int32_t *m_dev, int32_t *status, void *ws, size_t ws_bytes, tsg_stream_t stream);
int tsg_coord_table_build_dev(const int32_t *coords, int64_t n_cap, const int32_t *n_dev, void *table, int64_t slots,
                              int32_t *status, tsg_stream_t stream);
/* nbr is (K, n_cap): row stride = capacity; rows >= *n_dev are left untouched */
int tsg_kmap_build_dev(const void *table, int64_t slots, const int32_t *out_coords, int64_t n_cap, const int32_t *n_dev,
                       const int32_t *offsets_host, int k, int32_t *nbr, int32_t *nbsizes, int32_t *blockcnt,
                       tsg_stream_t stream);
/* ... the same, also emitting the K-bit "offset k has a neighbour" key of every row (row_keys, n_cap uint64; may be NULL) in the
 * bit order tsg_kmap_sort_rows sorts by, so that tsg_kmap_sort_rows_dev2 does not re-read the K x n table to build it */
int tsg_kmap_build_dev2(const void *table, int64_t slots, const int32_t *out_coords, int64_t n_cap, const int32_t *n_dev,
                        const int32_t *offsets_host, int k, int32_t *nbr, int32_t *nbsizes, int32_t *blockcnt,
                        uint64_t *row_keys, tsg_stream_t stream);
int tsg_kmap_sort_rows_dev2(const int32_t *nbr, int k, int64_t n_cap, const int32_t *n_dev, int64_t in_stride, int32_t *perm,
                            int32_t *nbr_sorted, int64_t out_stride, uint32_t *tile_mask, const uint64_t *row_keys, void *ws,
                            size_t ws_bytes, tsg_stream_t stream);
/* nbr (K, n_out_cap) -> nbr_t (K, n_in_cap), -1 where an input row has no output at that offset */
int tsg_kmap_transpose_dev(const int32_t *nbr, int k, int64_t n_out_cap, const int32_t *n_out_dev, int64_t n_in_cap,
                           int32_t *nbr_t, tsg_stream_t stream);
/* nbr (K, in_stride) -> perm (n_cap), nbr_sorted (K, out_stride) with -1 in rows >= *n_dev, tile_mask (ceil(n_cap/128)) */
int tsg_kmap_sort_rows_dev(const int32_t *nbr, int k, int64_t n_cap, const int32_t *n_dev, int64_t in_stride, int32_t *perm,
                           int32_t *nbr_sorted, int64_t out_stride, uint32_t *tile_mask, void *ws, size_t ws_bytes,
                           tsg_stream_t stream);
int tsg_gather_rows_dev(const void *src, int width, const int32_t *idx, int64_t n_cap, const int32_t *n_dev, void *out,
                        tsg_stream_t stream);
int tsg_cast_pad_bf16_dev(const float *in, int64_t n_cap, const int32_t *n_dev, int c, int c_pad, void *out,
                          tsg_stream_t stream);
/* tsg_aggregate_quantize with the frame table already on the device (no host copy inside); max_count >= every frame's count */
int tsg_aggregate_quantize_dev(const float *pts, int c_in, const tsg_frame *frames_dev, int n_frames, int64_t max_count,
                               int n_samples, const uint8_t *keep, float voxel_size, float *feats, int32_t *coords,
                               uint8_t *flags, void *ws, size_t ws_bytes, tsg_stream_t stream);

/* Tile-row order for tsg_conv_fwd_tc: stable sort of the n_out output rows by a K-bit key built from their neighbour
 * mask (offset k present iff nbr[k, o] >= 0; for K = 27 the rarest offsets — cube corners, then edges — are the most
 * significant key bits, otherwise bit k = offset k).  Outputs: perm (n_out) int32 = output row of tile row r, nbr_sorted (K, out_stride) with
 * nbr_sorted[k, r] = nbr[k, perm[r]] (-1 in the padding rows r >= n_out), tile_mask (ceil(n_out/128)) of the sorted table.  ws: tsg_kmap_sort_ws_bytes. */
size_t tsg_kmap_sort_ws_bytes(int64_t n_out);
int64_t tsg_kmap_sort_stride(int64_t n_out); /* row stride of nbr_sorted the convolution wants: n_out rounded up to 256 */
int tsg_kmap_sort_rows(const int32_t *nbr, int k, int64_t n_out, int32_t *perm, int32_t *nbr_sorted, int64_t out_stride,
                       uint32_t *tile_mask, void *ws, size_t ws_bytes, tsg_stream_t stream);

/* fp32 -> bf16 with zero padding of the channel dimension to c_pad (first-layer input, 4/5 -> 16 channels). */
int tsg_cast_pad_bf16(const float *in, int64_t n, int c, int c_pad, void *out, tsg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TASEG_B200_H */
