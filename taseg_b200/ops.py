"""Tensor-level host wrappers over the C ABI (allocation with torch's caching allocator + one call each).

Everything here launches on torch's current CUDA stream and never synchronises, except where a
data-dependent size must become a tensor shape (`.item()` on a device counter), which is marked `# sync`.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from ._lib import call, ptr, stream


def _i32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.int32 else t.to(torch.int32)


def _empty(shape, dtype, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device=like.device)


def _status(dev) -> torch.Tensor:
    return torch.zeros(1, dtype=torch.int32, device=dev)


def _check_status(status: torch.Tensor, what: str) -> None:
    if int(status.item()) & 1:
        raise RuntimeError(f"{what}: coordinate outside the packable range (|x|,|y|,|z| < 2^18, 0 <= b < 128)")


# ------------------------------------------------------------------------------- hashing / tables
def sphash(coords: torch.Tensor, offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    L.require_cuda(coords, offsets)
    coords = coords.contiguous()
    n = coords.shape[0]
    if offsets is None:
        out = _empty((n,), torch.int64, coords)
        call("tsg_hash", ptr(coords), n, ptr(out), stream())
        return out
    offsets = offsets.contiguous()
    k = offsets.shape[0]
    out = _empty((k, n), torch.int64, coords)
    call("tsg_kernel_hash", ptr(coords), n, ptr(offsets), k, ptr(out), stream())
    return out


class Table:
    """Open-addressing key -> row table living in one torch buffer (16 B per slot)."""

    def __init__(self, n: int, device):
        self.n = int(n)
        self.slots = int(L.lib().tsg_table_slots(self.n))
        self.buf = torch.empty(self.slots * 2, dtype=torch.int64, device=device)

    @classmethod
    def from_keys(cls, keys: torch.Tensor) -> "Table":
        L.require_cuda(keys)
        keys = keys.contiguous()
        t = cls(keys.numel(), keys.device)
        call("tsg_table_build", ptr(keys), t.n, ptr(t.buf), t.slots, stream())
        return t

    @classmethod
    def from_coords(cls, coords: torch.Tensor, status: Optional[torch.Tensor] = None) -> "Table":
        L.require_cuda(coords)
        coords = coords.contiguous()
        t = cls(coords.shape[0], coords.device)
        t.status = status if status is not None else _status(coords.device)
        call("tsg_coord_table_build", ptr(coords), t.n, ptr(t.buf), t.slots, ptr(t.status), stream())
        return t

    def query(self, queries: torch.Tensor) -> torch.Tensor:
        q = queries.contiguous()
        out = torch.empty_like(q)
        call("tsg_table_query", ptr(self.buf), self.slots, ptr(q), q.numel(), ptr(out), stream())
        return out


def sphashquery(queries: torch.Tensor, references: torch.Tensor) -> torch.Tensor:
    L.require_cuda(queries, references)
    return Table.from_keys(references).query(queries)


# ------------------------------------------------------------------------------- kernel maps
class KernelMap:
    """Kernel map of one (tensor stride, kernel, stride, dilation): the output-stationary neighbour table the
    convolution kernels consume, plus lazily materialised reference-format views.

    Indexable like the reference's `[nbmaps, nbsizes, (n_in, n_out)]` list (TS/nn/functional/conv.py:174-176)."""

    def __init__(self, nbr, nbsizes, blockcnt, n_in, n_out, k):
        self.nbr, self.nbsizes32, self.blockcnt = nbr, nbsizes, blockcnt
        self.n_in, self.n_out, self.k = int(n_in), int(n_out), int(k)
        self._nbmaps = None
        self._nbr_t = None
        self._tile_mask = None
        self._tile_mask_t = None
        self._sorted = None
        self._sorted_t = None

    @property
    def sizes(self) -> Tuple[int, int]:
        return (self.n_in, self.n_out)

    @property
    def nbsizes(self) -> torch.Tensor:
        return self.nbsizes32.to(torch.int64)

    @property
    def nbmaps(self) -> torch.Tensor:
        if self._nbmaps is None:
            total = int(self.nbsizes32.sum().item())  # sync: P is data dependent
            out = torch.empty((total, 2), dtype=torch.int64, device=self.nbr.device)
            if total:
                call("tsg_kmap_pairs", ptr(self.nbr), self.k, self.n_out, ptr(self.blockcnt.clone()), ptr(out), stream())
            self._nbmaps = out
        return self._nbmaps

    @property
    def nbr_t(self) -> torch.Tensor:
        if self._nbr_t is None:
            out = torch.empty((self.k, self.n_in), dtype=torch.int32, device=self.nbr.device)
            call("tsg_kmap_transpose", ptr(self.nbr), self.k, self.n_out, self.n_in, ptr(out), stream())
            self._nbr_t = out
        return self._nbr_t

    def tile_mask(self, transposed: bool = False) -> torch.Tensor:
        attr = "_tile_mask_t" if transposed else "_tile_mask"
        if getattr(self, attr) is None:
            nbr = self.nbr_t if transposed else self.nbr
            rows = self.n_in if transposed else self.n_out
            out = torch.empty(((rows + 127) // 128,), dtype=torch.int32, device=nbr.device)
            call("tsg_kmap_tile_mask", ptr(nbr), self.k, rows, ptr(out), stream())
            setattr(self, attr, out)
        return getattr(self, attr)

    def sorted(self, transposed: bool = False):
        """(nbr_sorted, tile_mask, perm): tile rows ordered by neighbour bit mask (tsg_kmap_sort_rows), so that the
        tensor-core convolution skips the (tile, offset) pairs that are empty; perm[r] = output row of tile row r."""
        attr = "_sorted_t" if transposed else "_sorted"
        if getattr(self, attr) is None:
            nbr = self.nbr_t if transposed else self.nbr
            rows = self.n_in if transposed else self.n_out
            dev = nbr.device
            perm = torch.empty((rows,), dtype=torch.int32, device=dev)
            stride = int(L.lib().tsg_kmap_sort_stride(rows))       # padded row stride (-1 beyond `rows`)
            nbr_s = torch.empty((self.k, stride), dtype=torch.int32, device=dev)
            mask = torch.empty(((rows + 127) // 128,), dtype=torch.int32, device=dev)
            ws_bytes = int(L.lib().tsg_kmap_sort_ws_bytes(rows))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            call("tsg_kmap_sort_rows", ptr(nbr), self.k, rows, ptr(perm), ptr(nbr_s), stride, ptr(mask), ptr(ws), ws_bytes,
                 stream())
            setattr(self, attr, (nbr_s, mask, perm))
        return getattr(self, attr)

    def split_items(self, transposed: bool = False):
        """SplitItems of the mask-sorted map when the launch has few enough tiles to want a K split, else None."""
        rows = self.n_in if transposed else self.n_out
        if not SplitItems.wanted(rows, self.k):
            return None
        attr = "_split_t" if transposed else "_split"
        if getattr(self, attr, None) is None:
            setattr(self, attr, SplitItems(self.sorted(transposed)[1], rows, self.k))
        return getattr(self, attr)

    def __getitem__(self, i):
        return (self.nbmaps, self.nbsizes, self.sizes)[i]

    def __iter__(self):
        return iter((self.nbmaps, self.nbsizes, self.sizes))

    def __len__(self):
        return 3


def build_kmap(table: Table, n_in: int, out_coords: torch.Tensor, offsets: np.ndarray) -> KernelMap:
    L.require_cuda(out_coords)
    out_coords = out_coords.contiguous()
    n_out = out_coords.shape[0]
    offs = np.ascontiguousarray(offsets, dtype=np.int32)
    k = offs.shape[0]
    dev = out_coords.device
    nbr = torch.empty((k, n_out), dtype=torch.int32, device=dev)
    nbsizes = torch.empty((k,), dtype=torch.int32, device=dev)
    blockcnt = torch.empty((k * max(int(L.lib().tsg_kmap_blocks(n_out)), 1),), dtype=torch.int32, device=dev)
    call("tsg_kmap_build", ptr(table.buf), table.slots, ptr(out_coords), n_out,
         offs.ctypes.data_as(ctypes.c_void_p), k, ptr(nbr), ptr(nbsizes), ptr(blockcnt), stream())
    return KernelMap(nbr, nbsizes, blockcnt, n_in, n_out, k)


def kmap_from_pairs(nbmaps: torch.Tensor, nbsizes: torch.Tensor, k: int, transposed: bool, n_rows_out: int) -> torch.Tensor:
    """Reference-format pairs -> neighbour table (for callers holding torchsparse-style kmaps)."""
    nbmaps = _i32(nbmaps).contiguous()
    ns = np.ascontiguousarray(nbsizes.detach().cpu().numpy(), dtype=np.int32)
    nbr = torch.empty((k, n_rows_out), dtype=torch.int32, device=nbmaps.device)
    call("tsg_kmap_from_pairs", ptr(nbmaps), ns.ctypes.data_as(ctypes.c_void_p), k, int(transposed), n_rows_out,
         ptr(nbr), stream())
    return nbr


# ------------------------------------------------------------------------------- sync-free variants (pipeline.py)
def table_from_coords_dev(coords: torch.Tensor, n_dev: torch.Tensor, status: torch.Tensor) -> Table:
    """Table over the first *n_dev rows of a capacity-sized coordinate buffer."""
    t = Table(coords.shape[0], coords.device)
    t.status = status
    call("tsg_coord_table_build_dev", ptr(coords), t.n, ptr(n_dev), ptr(t.buf), t.slots, ptr(status), stream())
    return t


def build_kmap_dev(table: Table, out_coords: torch.Tensor, n_dev: torch.Tensor, offsets: np.ndarray, want_keys: bool = False):
    """nbr (K, cap) int32 of the first *n_dev output rows (rows beyond are left untouched); with want_keys also the rows'
    neighbour-mask sort keys (cap,) int64 for kmap_sort_rows_dev(row_keys=...)."""
    cap = out_coords.shape[0]
    offs = np.ascontiguousarray(offsets, dtype=np.int32)
    k = offs.shape[0]
    dev = out_coords.device
    nbr = torch.empty((k, cap), dtype=torch.int32, device=dev)
    nbsizes = torch.empty((k,), dtype=torch.int32, device=dev)
    blockcnt = torch.empty((k * max(int(L.lib().tsg_kmap_blocks(cap)), 1),), dtype=torch.int32, device=dev)
    keys = torch.empty((cap,), dtype=torch.int64, device=dev) if want_keys else None
    call("tsg_kmap_build_dev2", ptr(table.buf), table.slots, ptr(out_coords), cap, ptr(n_dev),
         offs.ctypes.data_as(ctypes.c_void_p), k, ptr(nbr), ptr(nbsizes), ptr(blockcnt), ptr(keys), stream())
    return (nbr, keys) if want_keys else nbr


def kmap_transpose_dev(nbr: torch.Tensor, n_out_dev: torch.Tensor, n_in_cap: int) -> torch.Tensor:
    k, n_out_cap = nbr.shape
    out = torch.empty((k, n_in_cap), dtype=torch.int32, device=nbr.device)
    call("tsg_kmap_transpose_dev", ptr(nbr), k, n_out_cap, ptr(n_out_dev), n_in_cap, ptr(out), stream())
    return out


def kmap_sort_rows_dev(nbr: torch.Tensor, n_dev: torch.Tensor, row_keys: Optional[torch.Tensor] = None):
    """(nbr_sorted (K, stride), tile_mask, perm) of a capacity-sized table, as KernelMap.sorted().  row_keys: the keys
    build_kmap_dev(want_keys=True) emitted for this table."""
    k, cap = nbr.shape
    dev = nbr.device
    perm = torch.empty((cap,), dtype=torch.int32, device=dev)
    stride = int(L.lib().tsg_kmap_sort_stride(cap))
    nbr_s = torch.empty((k, stride), dtype=torch.int32, device=dev)
    mask = torch.empty(((cap + 127) // 128,), dtype=torch.int32, device=dev)
    ws_bytes = int(L.lib().tsg_kmap_sort_ws_bytes(cap))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    call("tsg_kmap_sort_rows_dev2", ptr(nbr), k, cap, ptr(n_dev), cap, ptr(perm), ptr(nbr_s), stride, ptr(mask), ptr(row_keys),
         ptr(ws), ws_bytes, stream())
    return nbr_s, mask, perm


def unique_coords_dev(coords: torch.Tensor, n_dev: torch.Tensor, out_cap: int, m_dev: torch.Tensor, status: torch.Tensor,
                      trunc_stride: int = 0, field_bits: Optional[Sequence[int]] = None, want_index: bool = False,
                      want_inverse: bool = False):
    """unique_coords over the first *n_dev rows of a capacity-sized buffer; the number of voxels goes to the device
    counter m_dev (clamped to out_cap, status bit 2 on overflow).  Returns (coords (out_cap,4)[, first][, inverse])."""
    coords = _i32(coords).contiguous()
    cap = coords.shape[0]
    dev = coords.device
    out_c = torch.empty((out_cap, 4), dtype=torch.int32, device=dev)
    first = torch.empty((out_cap,), dtype=torch.int32, device=dev) if want_index else None
    inv = torch.empty((cap,), dtype=torch.int32, device=dev) if want_inverse else None
    ws_bytes = int(L.lib().tsg_unique_ws_bytes(cap))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    fb = (ctypes.c_int32 * 4)(*[int(b) for b in field_bits]) if field_bits is not None else None
    call("tsg_unique_coords_dev", ptr(coords), cap, ptr(n_dev), int(trunc_stride), fb, ptr(out_c), int(out_cap), ptr(first),
         ptr(inv), ptr(m_dev), ptr(status), ptr(ws), ws_bytes, stream())
    res = [out_c]
    if want_index:
        res.append(first)
    if want_inverse:
        res.append(inv)
    return res[0] if len(res) == 1 else tuple(res)


def gather_rows_dev(src: torch.Tensor, idx: torch.Tensor, n_dev: torch.Tensor) -> torch.Tensor:
    """out[i] = src[idx[i]] for i < *n_dev (rows of 4-byte elements); idx is capacity-sized."""
    assert src.element_size() == 4
    src = src.contiguous()
    idx = _i32(idx).contiguous()
    width = src.shape[1] if src.dim() > 1 else 1
    out = torch.empty((idx.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    call("tsg_gather_rows_dev", ptr(src), width, ptr(idx), idx.shape[0], ptr(n_dev), ptr(out), stream())
    return out


# ------------------------------------------------------------------------------- unique voxels
def unique_coords(coords: torch.Tensor, trunc_stride: int = 0, want_index: bool = False, want_inverse: bool = False,
                  by_hash: bool = False, field_bits: Optional[Sequence[int]] = None):
    """Unique voxel rows ordered by (b,x,y,z) (or by ascending FNV hash); returns (coords[, first_idx][, inverse]).
    field_bits = (bx,by,bz,bb): promise that 0 <= coordinate < 2^bits, which shortens the radix sort."""
    L.require_cuda(coords)
    coords = _i32(coords).contiguous()
    n = coords.shape[0]
    dev = coords.device
    out_c = torch.empty((n, 4), dtype=torch.int32, device=dev)
    first = torch.empty((n,), dtype=torch.int32, device=dev) if want_index else None
    inv = torch.empty((n,), dtype=torch.int32, device=dev) if want_inverse else None
    m_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = int(L.lib().tsg_unique_ws_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if by_hash:
        call("tsg_unique_hash", ptr(coords), n, ptr(out_c), ptr(first), ptr(inv), ptr(m_dev), ptr(ws), ws_bytes, stream())
        m = int(m_dev.item())  # sync
    else:
        status = _status(dev)
        fb = (ctypes.c_int32 * 4)(*[int(b) for b in field_bits]) if field_bits is not None else None
        call("tsg_unique_coords", ptr(coords), n, int(trunc_stride), fb, ptr(out_c), ptr(first), ptr(inv), ptr(m_dev),
             ptr(status), ptr(ws), ws_bytes, stream())
        m, st = torch.cat([m_dev, status]).tolist()  # sync
        if st & 1:
            raise RuntimeError("unique_coords: coordinate outside the packable range (|x|,|y|,|z| < 2^18, 0 <= b < 128) "
                               "or outside the promised field_bits")
    res = [out_c[:m]]
    if want_index:
        res.append(first[:m])
    if want_inverse:
        res.append(inv)
    return res[0] if len(res) == 1 else tuple(res)


def sort_pairs(keys: torch.Tensor, vals: Optional[torch.Tensor] = None, begin_bit: int = 0, end_bit: int = 64):
    keys = keys.contiguous()
    n = keys.numel()
    ko = torch.empty_like(keys)
    vo = torch.empty((n,), dtype=torch.int32, device=keys.device)
    ws_bytes = int(L.lib().tsg_sort_ws_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
    call("tsg_sort_pairs", ptr(keys), ptr(vals), n, begin_bit, end_bit, ptr(ko), ptr(vo), ptr(ws), ws_bytes, stream())
    return ko, vo


# ------------------------------------------------------------------------------- point <-> voxel
def spcount(idx: torch.Tensor, num: int) -> torch.Tensor:
    L.require_cuda(idx)
    idx = _i32(idx).contiguous()
    out = torch.empty((int(num),), dtype=torch.int32, device=idx.device)
    call("tsg_count", ptr(idx), idx.numel(), ptr(out), int(num), stream())
    return out


class VoxelizePlan:
    """Points grouped by voxel (stable radix sort of the point ids by `idx`): what the segmented scatter-mean walks.
    Built once per index tensor and cached on it, so SPVCNN's repeated point_to_voxel over the same points sorts once."""

    def __init__(self, idx: torch.Tensor, m: int):
        idx = _i32(idx).contiguous()
        n, dev = idx.numel(), idx.device
        self.n, self.m = n, int(m)
        self.order = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
        skeys = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)      # scratch: the sorted voxel ids
        self.seg = torch.empty((max(self.m, 1), 2), dtype=torch.int32, device=dev)
        ws_bytes = int(L.lib().tsg_voxelize_plan_ws_bytes(n))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        call("tsg_voxelize_plan", ptr(idx), n, self.m, ptr(self.order), ptr(skeys), ptr(self.seg), ptr(ws), ws_bytes, stream())

    @staticmethod
    def of(idx: torch.Tensor, m: int) -> "VoxelizePlan":
        key = (idx.data_ptr(), idx._version, int(m), idx.numel())
        cached = getattr(idx, "_tsg_plan", None)
        if cached is None or cached[0] != key:
            cached = (key, VoxelizePlan(idx, m))
            try:
                idx._tsg_plan = cached
            except AttributeError:
                pass
        return cached[1]


def _vec_ok(a: torch.Tensor, b: torch.Tensor, c: int) -> bool:
    return bool(L.lib().tsg_pv_vector_ok(ptr(a), ptr(b), int(c), L.DTYPES[a.dtype]))


def voxelize_forward(feats: torch.Tensor, idx: torch.Tensor, counts: torch.Tensor, plan: Optional[VoxelizePlan] = None) -> torch.Tensor:
    """Scatter-mean of point rows into voxel rows: deterministic segmented reduction over a VoxelizePlan (vector rows),
    or the reference-style atomic kernel for rows that are not a multiple of 16 bytes."""
    L.require_cuda(feats, idx, counts)
    feats = feats.contiguous()
    idx, counts = _i32(idx).contiguous(), _i32(counts).contiguous()
    n, c = feats.shape
    m = counts.shape[0]
    out = torch.empty((m, c), dtype=feats.dtype, device=feats.device)
    if m and c and _vec_ok(feats, out, c):
        plan = plan or VoxelizePlan.of(idx, m)
        call("tsg_voxelize_fwd_seg", ptr(feats), L.DTYPES[feats.dtype], ptr(plan.order), ptr(plan.seg), ptr(counts), n, c, m,
             ptr(out), stream())
        return out
    acc = None if feats.dtype == torch.float32 else torch.empty((m, c), dtype=torch.float32, device=feats.device)
    call("tsg_voxelize_fwd", ptr(feats), L.DTYPES[feats.dtype], ptr(idx), ptr(counts), n, c, m, ptr(out), ptr(acc), stream())
    return out


def voxelize_backward(top_grad: torch.Tensor, idx: torch.Tensor, counts: torch.Tensor, n: int) -> torch.Tensor:
    top_grad = top_grad.contiguous()
    idx, counts = _i32(idx).contiguous(), _i32(counts).contiguous()
    c = top_grad.shape[1]
    out = torch.empty((n, c), dtype=top_grad.dtype, device=top_grad.device)
    name = "tsg_voxelize_bwd_vec" if (n and c and _vec_ok(top_grad, out, c)) else "tsg_voxelize_bwd"
    call(name, ptr(top_grad), L.DTYPES[top_grad.dtype], ptr(idx), ptr(counts), n, c, ptr(out), stream())
    return out


def devoxelize_forward(feats: torch.Tensor, idx8: torch.Tensor, w8: torch.Tensor) -> torch.Tensor:
    L.require_cuda(feats, idx8, w8)
    feats = feats.contiguous()
    idx8, w8 = _i32(idx8).contiguous(), w8.float().contiguous()
    n, c = idx8.shape[0], feats.shape[1]
    out = torch.empty((n, c), dtype=feats.dtype, device=feats.device)
    name = "tsg_devoxelize_fwd_vec" if (n and c and _vec_ok(feats, out, c)) else "tsg_devoxelize_fwd"
    call(name, ptr(feats), L.DTYPES[feats.dtype], ptr(idx8), ptr(w8), n, c, ptr(out), stream())
    return out


def devoxelize_backward(top_grad: torch.Tensor, idx8: torch.Tensor, w8: torch.Tensor, m: int) -> torch.Tensor:
    top_grad = top_grad.contiguous()
    idx8, w8 = _i32(idx8).contiguous(), w8.float().contiguous()
    n, c = top_grad.shape
    out = torch.empty((m, c), dtype=top_grad.dtype, device=top_grad.device)
    acc = None if top_grad.dtype == torch.float32 else torch.empty((m, c), dtype=torch.float32, device=top_grad.device)
    name = "tsg_devoxelize_bwd_vec" if (m and c and c % 4 == 0 and _vec_ok(top_grad, out, c)) else "tsg_devoxelize_bwd"
    call(name, ptr(top_grad), L.DTYPES[top_grad.dtype], ptr(idx8), ptr(w8), n, c, m, ptr(out), ptr(acc), stream())
    return out


def devoxelize_multi(tables: Sequence[Table], strides: Sequence[int], feats: Sequence[torch.Tensor], pcoords: torch.Tensor,
                     c_out: int, rows: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[j] = sum over scales of trilinear_devoxelize(feats_s) at point pcoords[rows[j]] (rows None: every point)."""
    ns = len(tables)
    pc = pcoords.float().contiguous()
    feats = [f.contiguous() for f in feats]
    c = feats[0].shape[1]
    assert all(f.dtype == torch.float32 and f.shape[1] == c for f in feats)
    rows = _i32(rows).contiguous() if rows is not None else None
    m = rows.shape[0] if rows is not None else pc.shape[0]
    out = torch.empty((m, c_out), dtype=torch.float32, device=pc.device)
    tabs = (ctypes.c_void_p * ns)(*[t.buf.data_ptr() for t in tables])
    slots = (ctypes.c_int64 * ns)(*[t.slots for t in tables])
    strd = (ctypes.c_int32 * ns)(*[int(v) for v in strides])
    fts = (ctypes.c_void_p * ns)(*[f.data_ptr() for f in feats])
    call("tsg_devoxelize_multi", ns, tabs, slots, strd, fts, c, ptr(pc), ptr(rows), m, ptr(out), int(c_out), stream())
    return out


def point_query(table: Table, pcoords: torch.Tensor, stride: int) -> torch.Tensor:
    pc = pcoords.float().contiguous()
    out = torch.empty((pc.shape[0],), dtype=torch.int32, device=pc.device)
    call("tsg_point_query", ptr(table.buf), table.slots, ptr(pc), pc.shape[0], int(stride), ptr(out), stream())
    return out


def trilinear_query(table: Table, pcoords: torch.Tensor, stride: int, nearest: bool = False):
    pc = pcoords.float().contiguous()
    n = pc.shape[0]
    idx8 = torch.empty((n, 8), dtype=torch.int32, device=pc.device)
    w8 = torch.empty((n, 8), dtype=torch.float32, device=pc.device)
    call("tsg_trilinear_query", ptr(table.buf), table.slots, ptr(pc), n, int(stride), int(nearest), ptr(idx8), ptr(w8), stream())
    return idx8, w8


def rescale_coords(pcoords: torch.Tensor, init_res: float, after_res: float):
    pc = pcoords.float().contiguous()
    n = pc.shape[0]
    out_f = torch.empty((n, 4), dtype=torch.float32, device=pc.device)
    out_i = torch.empty((n, 4), dtype=torch.int32, device=pc.device)
    call("tsg_rescale_coords", ptr(pc), n, float(init_res), float(after_res), ptr(out_f), ptr(out_i), stream())
    return out_f, out_i


def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[i] = src[idx[i]] for rows of 4-byte elements (fp32 / int32)."""
    assert src.element_size() == 4
    src = src.contiguous()
    idx = _i32(idx).contiguous()
    width = src.shape[1] if src.dim() > 1 else 1
    out = torch.empty((idx.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    call("tsg_gather_rows", ptr(src), width, ptr(idx), idx.shape[0], ptr(out), stream())
    return out


def compact_rows(flags: torch.Tensor, rows_a: Optional[torch.Tensor], rows_b: Optional[torch.Tensor] = None,
                 want_pos: bool = False, sync: bool = True, m_dev: Optional[torch.Tensor] = None):
    """Stable compaction by a uint8 flag; returns (out_a, out_b, pos, m).  With sync=False the outputs keep their
    upper-bound length and m is the device counter (the caller slices after its own readback)."""
    n = flags.numel()
    dev = flags.device
    wa = rows_a.shape[1] if rows_a is not None else 0
    wb = rows_b.shape[1] if rows_b is not None else 0
    out_a = torch.empty_like(rows_a) if rows_a is not None else None
    out_b = torch.empty_like(rows_b) if rows_b is not None else None
    pos = torch.empty((n,), dtype=torch.int32, device=dev) if want_pos else None
    if m_dev is None:
        m_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = int(L.lib().tsg_compact_ws_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    call("tsg_compact_rows", ptr(flags), n, ptr(rows_a), wa, ptr(out_a), ptr(rows_b), wb, ptr(out_b), ptr(pos), ptr(m_dev),
         ptr(ws), ws_bytes, stream())
    if not sync:
        return out_a, out_b, pos, m_dev
    m = int(m_dev.item())  # sync
    return (out_a[:m] if out_a is not None else None, out_b[:m] if out_b is not None else None, pos, m)


# ------------------------------------------------------------------------------- convolution
def conv_forward(feats: torch.Tensor, weight: torch.Tensor, nbr: torch.Tensor, n_out: int,
                 scale: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                 residual: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """CUDA-core implicit GEMM (fp32/bf16/fp16 storage, fp32 accumulate)."""
    L.require_cuda(feats, weight, nbr)
    feats = feats.contiguous()
    weight = weight.contiguous().to(feats.dtype)
    if weight.dim() == 2:
        weight = weight.unsqueeze(0)
    k, c_in_w, c_out = weight.shape
    out = torch.empty((n_out, c_out), dtype=feats.dtype, device=feats.device)
    call("tsg_conv_fwd", ptr(feats), L.DTYPES[feats.dtype], feats.shape[0], feats.shape[1], ptr(weight), k, c_in_w, c_out,
         ptr(nbr), int(n_out), ptr(out), ptr(scale), ptr(bias), ptr(residual), int(relu), stream())
    return out


def conv_dgrad(grad_out: torch.Tensor, weight: torch.Tensor, nbr_t: torch.Tensor, n_in: int) -> torch.Tensor:
    grad_out = grad_out.float().contiguous()
    weight = weight.float().contiguous()
    k, c_in, c_out = weight.shape
    out = torch.empty((n_in, c_in), dtype=torch.float32, device=grad_out.device)
    call("tsg_conv_dgrad", ptr(grad_out), grad_out.shape[0], c_out, ptr(weight), k, c_in, ptr(nbr_t), int(n_in), ptr(out), stream())
    return out


def conv_wgrad(feats: torch.Tensor, grad_out: torch.Tensor, nbr: torch.Tensor, k: int) -> torch.Tensor:
    feats = feats.float().contiguous()
    grad_out = grad_out.float().contiguous()
    c_in, c_out = feats.shape[1], grad_out.shape[1]
    gw = torch.empty((k, c_in, c_out), dtype=torch.float32, device=feats.device)
    call("tsg_conv_wgrad", ptr(feats), feats.shape[0], c_in, ptr(grad_out), grad_out.shape[0], c_out, ptr(nbr), k, ptr(gw), stream())
    return gw


def conv_wgrad_bf16(feats: torch.Tensor, grad_out: torch.Tensor, nbr: torch.Tensor, k: int) -> torch.Tensor:
    """Weight gradient of the autocast path: bf16 rows in, fp32 (K, c_in, c_out) out (tensor-core MMAs over the real pairs)."""
    feats = feats.to(torch.bfloat16).contiguous()
    grad_out = grad_out.to(torch.bfloat16).contiguous()
    c_in, c_out = feats.shape[1], grad_out.shape[1]
    gw = torch.empty((k, c_in, c_out), dtype=torch.float32, device=feats.device)
    call("tsg_conv_wgrad_bf16", ptr(feats), feats.shape[0], c_in, ptr(grad_out), grad_out.shape[0], c_out, ptr(nbr), k, ptr(gw),
         stream())
    return gw


# ------------------------------------------------------------------------------- BatchNorm over feature rows
FUSED_BN = os.environ.get("TSG_BN", "1") != "0"      # A/B switch: 0 = ATen's batch_norm


def bn_supported(x: torch.Tensor) -> bool:
    return (FUSED_BN and x.is_cuda and x.dim() == 2 and x.dtype in (torch.float32, torch.bfloat16) and x.shape[0] > 0
            and x.shape[1] % 8 == 0 and x.shape[1] <= 1024)


def _bn_ws(x: torch.Tensor) -> torch.Tensor:
    return torch.empty(int(L.lib().tsg_bn_ws_bytes(x.shape[0], x.shape[1])), dtype=torch.uint8, device=x.device)


def bn_stats(x: torch.Tensor, eps: float, momentum: float, running_mean: Optional[torch.Tensor],
             running_var: Optional[torch.Tensor], batches_tracked: Optional[torch.Tensor] = None):
    """(mean, invstd) fp32 of the rows of x; running statistics (fp32, contiguous) updated in place, `batches_tracked` (int64
    scalar on the device) incremented in the same launch."""
    assert batches_tracked is None or (batches_tracked.dtype == torch.int64 and batches_tracked.is_cuda)
    x = x.contiguous()
    n, c = x.shape
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    invstd = torch.empty(c, dtype=torch.float32, device=x.device)
    ws = _bn_ws(x)
    call("tsg_bn_stats2", ptr(x), L.DTYPES[x.dtype], n, c, float(eps), float(momentum), ptr(running_mean), ptr(running_var),
         ptr(batches_tracked), ptr(mean), ptr(invstd), ptr(ws), ws.numel(), stream())
    return mean, invstd


def bn_apply(x: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor, gamma: Optional[torch.Tensor],
             beta: Optional[torch.Tensor], relu: bool = False) -> torch.Tensor:
    x = x.contiguous()
    y = torch.empty_like(x)
    call("tsg_bn_apply", ptr(x), L.DTYPES[x.dtype], x.shape[0], x.shape[1], ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), int(relu),
         ptr(y), stream())
    return y


def bn_backward(x: torch.Tensor, dy: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor, gamma: Optional[torch.Tensor],
                training: bool, want_dx: bool = True, beta: Optional[torch.Tensor] = None, relu: bool = False):
    """(dx | None, dgamma, dbeta); relu: dy is the gradient of relu(bn(x)) (training mode only)."""
    x, dy = x.contiguous(), dy.contiguous().to(x.dtype)
    n, c = x.shape
    sums = torch.empty((2, c), dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x) if want_dx else None
    ws = _bn_ws(x)
    call("tsg_bn_backward", ptr(x), ptr(dy), L.DTYPES[x.dtype], n, c, ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), int(relu),
         int(training), ptr(sums), ptr(dx), ptr(ws), ws.numel(), stream())
    return dx, sums[1], sums[0]


WGRAD_TC = os.environ.get("TSG_WGRAD_TC", "1") != "0"   # A/B switch: tcgen05 weight gradient (0: the warp-level MMA kernel)


class PairList:
    """Compact per-offset {in row, out row} lists of a neighbour table (tsg_kmap_pair_list): the K dimension of the
    weight-gradient GEMMs.  Built once per table and cached on it, so every layer that shares a kernel map shares it."""

    def __init__(self, nbr: torch.Tensor):
        assert nbr.dtype == torch.int32 and nbr.dim() == 2 and nbr.is_contiguous()
        k, n = nbr.shape
        dev = nbr.device
        self.k, self.cap = k, k * n
        self.pairs = torch.empty((max(self.cap, 1), 2), dtype=torch.int32, device=dev)
        self.start = torch.zeros(k + 1, dtype=torch.int32, device=dev)
        if n:
            ws_bytes = int(L.lib().tsg_kmap_pair_list_ws_bytes(k, n))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            call("tsg_kmap_pair_list", ptr(nbr), k, n, nbr.stride(0), ptr(self.pairs), ptr(self.start), ptr(ws), ws_bytes, stream())

    @staticmethod
    def of(nbr: torch.Tensor) -> "PairList":
        key = (nbr.data_ptr(), nbr._version, tuple(nbr.shape))
        cached = getattr(nbr, "_tsg_pairs", None)
        if cached is None or cached[0] != key:
            cached = (key, PairList(nbr))
            try:
                nbr._tsg_pairs = cached
            except AttributeError:
                pass
        return cached[1]


def conv_wgrad_tc(feats: torch.Tensor, grad_out: torch.Tensor, nbr: torch.Tensor, k: int,
                  pairs: Optional[PairList] = None) -> torch.Tensor:
    """Weight gradient on tcgen05 (tsg_conv_wgrad_tc): bf16 rows in, fp32 (K, c_in, c_out) out; c_out <= 256.
    nbr (K, n_out): rows = outputs of the forward, values = input rows; its pair list is built once and cached."""
    L.require_cuda(feats, grad_out, nbr)
    feats = feats.to(torch.bfloat16).contiguous()
    grad_out = grad_out.to(torch.bfloat16).contiguous()
    c_in, c_out = feats.shape[1], grad_out.shape[1]
    pl = pairs if pairs is not None else PairList.of(nbr)
    gw = torch.empty((k, c_in, c_out), dtype=torch.float32, device=feats.device)
    call("tsg_conv_wgrad_tc", ptr(feats), feats.shape[0], c_in, ptr(grad_out), grad_out.shape[0], c_out, ptr(pl.pairs),
         ptr(pl.start), pl.cap, k, ptr(gw), stream())
    return gw


def pad16(c: int) -> int:
    return (c + 15) // 16 * 16


def pack_weights(weight: torch.Tensor, c0: int, c1: int = 0, out_scale: Optional[torch.Tensor] = None,
                 c_out_pad: Optional[int] = None) -> torch.Tensor:
    """(K, c_in, c_out) fp32 -> tcgen05 shared-memory image (bf16, swizzled), optionally BN-scale folded and with the
    output channels zero-padded to c_out_pad."""
    w = weight.detach().float()
    if w.dim() == 2:
        w = w.unsqueeze(0)
    k, c_in, c_out = w.shape
    if out_scale is not None:
        out_scale = out_scale.detach().float().contiguous()
    if c_out_pad is not None and c_out_pad != c_out:
        w = torch.nn.functional.pad(w, (0, c_out_pad - c_out))
        if out_scale is not None:
            out_scale = torch.nn.functional.pad(out_scale, (0, c_out_pad - c_out))
        c_out = c_out_pad
    nbytes = int(L.lib().tsg_conv_pack_bytes(k, c0, c1, c_out))
    packed = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    sk, sci, sco = w.stride()          # transposed / column-sliced views are packed in place (element strides)
    call("tsg_conv_pack_weights2", ptr(w), k, c_in, c_out, sk, sci, sco, c0, c1, ptr(out_scale), ptr(packed), stream())
    return packed


def cast_pad_bf16(feats: torch.Tensor, c_pad: int, n_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    feats = feats.float().contiguous()
    n, c = feats.shape
    out = torch.empty((n, c_pad), dtype=torch.bfloat16, device=feats.device)
    if n_dev is not None:
        call("tsg_cast_pad_bf16_dev", ptr(feats), n, ptr(n_dev), c, c_pad, ptr(out), stream())
    else:
        call("tsg_cast_pad_bf16", ptr(feats), n, c, c_pad, ptr(out), stream())
    return out


_SCHED = {}
SCHED_OVERRIDE = None   # a captured pipeline owns its scheduler counters (graphs replay on any stream, possibly concurrently)


def _sched_ws(device) -> torch.Tensor:
    """Two zeroed int32 per (device, stream) for the convolution's dynamic tile scheduler (the kernel re-zeroes them).
    Launches that may overlap need their own pair: eager launches get one per stream; a Pipeline passes its own."""
    if SCHED_OVERRIDE is not None:
        return SCHED_OVERRIDE
    key = (device, stream().value)
    t = _SCHED.get(key)
    if t is None:
        t = _SCHED[key] = torch.zeros(2, dtype=torch.int32, device=device)
    return t


DYNAMIC_TILES = os.environ.get("TSG_DYNAMIC_TILES", "1") != "0"   # A/B switch: in-kernel dynamic tile scheduler
SPLIT_K = os.environ.get("TSG_SPLIT_K", "0") != "0"               # K-split work items for launches with few tiles
SPLIT_MAX_TILES = int(os.environ.get("TSG_SPLIT_MAX_TILES", "222"))   # ... up to 1.5 tiles per SM of a B200
SPLIT_CAP = int(os.environ.get("TSG_SPLIT_CAP", "14"))             # a tile with more active offsets is summed by ceil(active / cap) items
SPLIT_PARTS = int(os.environ.get("TSG_SPLIT_PARTS", "2"))          # ... at most this many (measured: profiles/r02/ksplit_nway.txt)


class SplitItems:
    """Work-item list of a K-split tensor-core launch (tsg_conv_split_items) for one (mask-sorted) kernel map."""

    def __init__(self, tile_mask: torch.Tensor, n_out: int, k: int, n_dev: Optional[torch.Tensor] = None,
                 cap: int = 0, max_slots: int = 0, max_parts: int = 0):
        tiles = (int(n_out) + 127) // 128
        self.max_slots = int(max_slots) if max_slots else min(tiles, 128)
        self.max_parts = max(2, int(max_parts) if max_parts else SPLIT_PARTS)
        dev = tile_mask.device
        self.items = torch.empty((tiles + self.max_slots * (self.max_parts - 1), 4), dtype=torch.int32, device=dev)
        self.n_items = torch.zeros(1, dtype=torch.int32, device=dev)
        self.state = torch.zeros(self.max_slots * 16, dtype=torch.int32, device=dev)
        call("tsg_conv_split_items", ptr(tile_mask), int(n_out), ptr(n_dev), int(k), int(cap or SPLIT_CAP), self.max_parts,
             self.max_slots, ptr(self.items), ptr(self.n_items), stream())

    @staticmethod
    def wanted(n_out: int, k: int) -> bool:
        return SPLIT_K and DYNAMIC_TILES and k > SPLIT_CAP and (int(n_out) + 127) // 128 <= SPLIT_MAX_TILES

PROFILE = None   # set to a list to record (tag, start_event, end_event, pairs, c_in, c_out) per tensor-core launch


def conv_forward_tc(in0: torch.Tensor, in1: Optional[torch.Tensor], packed_w: torch.Tensor, k: int, c_out: int,
                    nbr: torch.Tensor, tile_mask: torch.Tensor, n_out: int, bias: Optional[torch.Tensor] = None,
                    residual: Optional[torch.Tensor] = None, relu: bool = False, out_dtype=torch.bfloat16,
                    num_sms: int = 0, perm: Optional[torch.Tensor] = None, shortcut=None,
                    n_dev: Optional[torch.Tensor] = None, split: Optional["SplitItems"] = None) -> torch.Tensor:
    """tcgen05/TMEM implicit GEMM; in0/in1 bf16 (n_in, c) with c % 16 == 0.  With `perm`, nbr/tile_mask are in the
    mask-sorted tile-row order of KernelMap.sorted() and tile row r is written to out[perm[r]].
    shortcut = (sc_in0, sc_in1 | None, sc_packed_w, sc_idx | None): a 1x1x1 convolution of (sc_in0 | sc_in1) (n_out rows)
    accumulated into the same tile (tsg_conv_fwd_tc2); sc_idx = the centre offset's line of `nbr` for sorted maps.
    n_dev: int32 device counter — n_out is then the capacity of the buffers and min(*n_dev, n_out) rows are computed.
    split: SplitItems of this map (K-split work items; launches with about as many tiles as SMs)."""
    L.require_cuda(in0, in1, packed_w)
    assert in0.dtype == torch.bfloat16 and in0.is_contiguous()
    c0 = in0.shape[1]
    c1 = 0
    if in1 is not None:
        assert in1.dtype == torch.bfloat16 and in1.is_contiguous() and in1.shape[0] == in0.shape[0]
        c1 = in1.shape[1]
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.is_contiguous() and residual.shape == (n_out, c_out)
    out = torch.empty((n_out, c_out), dtype=out_dtype, device=in0.device)
    nbr_stride = 0
    if nbr is not None:
        want = (n_out + 255) // 256 * 256
        if nbr.shape[1] < want:     # plain (K, n_out) table: the kernel wants a 256-padded stride with -1 in the padding
            nbr = torch.nn.functional.pad(nbr, (0, want - nbr.shape[1]), value=-1)
        assert nbr.is_contiguous()
        nbr_stride = nbr.shape[1]
    s0 = s1 = sw = sidx = None
    sc0 = sc1 = 0
    if shortcut is not None:
        s0, s1, sw, sidx = shortcut
        assert s0.dtype == torch.bfloat16 and s0.is_contiguous() and s0.shape[0] == n_out
        sc0 = s0.shape[1]
        if s1 is not None:
            assert s1.dtype == torch.bfloat16 and s1.is_contiguous() and s1.shape[0] == n_out
            sc1 = s1.shape[1]
        if sidx is not None:
            assert sidx.dtype == torch.int32 and sidx.is_contiguous() and sidx.numel() >= nbr_stride
    if PROFILE is not None:
        rows = n_dev[0].double() if n_dev is not None else float(n_out)
        if nbr is None:
            pairs = rows * k + torch.zeros((), dtype=torch.float64, device=in0.device)
        elif n_dev is None:
            pairs = (nbr >= 0).sum().double()
        else:      # capacity-sized table: rows beyond the device count of a sorted table hold -1, of a raw table garbage
            live = torch.arange(nbr.shape[1], device=nbr.device) < n_dev[0]
            pairs = ((nbr >= 0) & live).sum().double()
        if shortcut is not None:     # the folded 1x1 convolution's MACs, expressed in pairs of the main phase's width
            pairs = pairs + rows * (sc0 + sc1) / (c0 + c1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if split is not None and DYNAMIC_TILES:
        scratch = torch.empty((split.max_slots * split.max_parts * 128 * c_out,), dtype=torch.float32, device=in0.device)
        call("tsg_conv_fwd_tc4", ptr(in0), c0, ptr(in1), c1, in0.shape[0], ptr(packed_w), k, c_out, ptr(nbr), nbr_stride,
             ptr(tile_mask), ptr(perm), int(n_out), ptr(n_dev), ptr(s0), sc0, ptr(s1), sc1, ptr(sw), ptr(sidx), ptr(out),
             L.DTYPES[out_dtype], ptr(bias), ptr(residual), int(relu), int(num_sms), ptr(_sched_ws(in0.device)),
             ptr(split.items), ptr(split.n_items), split.max_slots, split.max_parts, ptr(scratch), ptr(split.state), stream())
    else:
        call("tsg_conv_fwd_tc3", ptr(in0), c0, ptr(in1), c1, in0.shape[0], ptr(packed_w), k, c_out, ptr(nbr), nbr_stride,
             ptr(tile_mask), ptr(perm), int(n_out), ptr(n_dev), ptr(s0), sc0, ptr(s1), sc1, ptr(sw), ptr(sidx), ptr(out), L.DTYPES[out_dtype],
             ptr(bias), ptr(residual), int(relu), int(num_sms), ptr(_sched_ws(in0.device)) if DYNAMIC_TILES else None, stream())
    if PROFILE is not None:
        e1.record()
        PROFILE.append((k, e0, e1, pairs, c0 + c1, c_out, rows if n_dev is not None else n_out))
    return out


# ------------------------------------------------------------------------------- multi-frame front end
def fuse_multi_scan(points: torch.Tensor, pose0, pose) -> torch.Tensor:
    L.require_cuda(points)
    pts = points.float().contiguous()
    p0 = (ctypes.c_float * 16)(*np.asarray(pose0, np.float32).reshape(-1).tolist())
    p1 = (ctypes.c_float * 16)(*np.asarray(pose, np.float32).reshape(-1).tolist())
    out = torch.empty_like(pts)
    call("tsg_fuse_multi_scan", ptr(pts), pts.shape[0], pts.shape[1], p0, p1, ptr(out), stream())
    return out


def transform_point(points: torch.Tensor, R, T) -> torch.Tensor:
    L.require_cuda(points)
    pts = points.float().contiguous()
    r = (ctypes.c_double * 9)(*np.asarray(R, np.float64).reshape(-1).tolist())
    t = (ctypes.c_double * 3)(*np.asarray(T, np.float64).reshape(-1).tolist())
    out = torch.empty_like(pts)
    call("tsg_transform_point", ptr(pts), pts.shape[0], pts.shape[1], r, t, ptr(out), stream())
    return out


def aggregate_quantize(points: torch.Tensor, frames: Sequence[dict], n_samples: int, voxel_size: float,
                       keep: Optional[torch.Tensor] = None):
    """points (sum n, c_in) fp32, frames: dicts(offset,count,sample,is_cur,pose0,pose).
    Returns feats (sum n, c_in+1), coords (sum n, 4) int32, flags (sum n) uint8, extent (n_samples, 12) int32 view of
    the per-sample records [cur_min(4 f32 bits), ms_min(4), ms_max(4)]."""
    L.require_cuda(points, keep)
    pts = points.float().contiguous()
    n, c_in = pts.shape
    arr = (L.Frame * len(frames))()
    ident = np.eye(4, dtype=np.float32).reshape(-1)
    for i, f in enumerate(frames):
        arr[i].offset, arr[i].count, arr[i].sample, arr[i].is_cur = int(f["offset"]), int(f["count"]), int(f["sample"]), int(f["is_cur"])
        p0 = np.asarray(f.get("pose0", ident), np.float32).reshape(-1)
        p1 = np.asarray(f.get("pose", ident), np.float32).reshape(-1)
        for j in range(16):
            arr[i].pose0[j] = float(p0[j])
            arr[i].pose[j] = float(p1[j])
    feats = torch.empty((n, c_in + 1), dtype=torch.float32, device=pts.device)
    coords = torch.empty((n, 4), dtype=torch.int32, device=pts.device)
    flags = torch.empty((n,), dtype=torch.uint8, device=pts.device)
    ws_bytes = int(L.lib().tsg_aggregate_ws_bytes(n_samples))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pts.device)
    call("tsg_aggregate_quantize", ptr(pts), c_in, arr, len(frames), n_samples, ptr(keep), float(voxel_size), ptr(feats),
         ptr(coords), ptr(flags), ptr(ws), ws_bytes, stream())
    extent = ws[:n_samples * 48].view(torch.int32).view(n_samples, 12)
    return feats, coords, flags, extent


def frames_to_array(frames: Sequence[dict]):
    """ctypes array of tsg_frame records from the dicts of frontend.MultiFrameBatch.frames."""
    arr = (L.Frame * len(frames))()
    ident = np.eye(4, dtype=np.float32).reshape(-1)
    for i, f in enumerate(frames):
        arr[i].offset, arr[i].count, arr[i].sample, arr[i].is_cur = int(f["offset"]), int(f["count"]), int(f["sample"]), int(f["is_cur"])
        p0 = np.asarray(f.get("pose0", ident), np.float32).reshape(-1)
        p1 = np.asarray(f.get("pose", ident), np.float32).reshape(-1)
        for j in range(16):
            arr[i].pose0[j] = float(p0[j])
            arr[i].pose[j] = float(p1[j])
    return arr


def aggregate_quantize_dev(points: torch.Tensor, frames_dev: torch.Tensor, n_frames: int, max_count: int, n_samples: int,
                           voxel_size: float, keep: Optional[torch.Tensor] = None):
    """aggregate_quantize with the frame table already on the device (uint8 tensor holding n_frames tsg_frame records):
    no host copy inside, so the call can be captured into a CUDA graph."""
    pts = points.float().contiguous()
    n, c_in = pts.shape
    feats = torch.empty((n, c_in + 1), dtype=torch.float32, device=pts.device)
    coords = torch.empty((n, 4), dtype=torch.int32, device=pts.device)
    flags = torch.empty((n,), dtype=torch.uint8, device=pts.device)
    ws_bytes = int(L.lib().tsg_aggregate_ws_bytes(n_samples))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pts.device)
    call("tsg_aggregate_quantize_dev", ptr(pts), c_in, ptr(frames_dev), int(n_frames), int(max_count), int(n_samples), ptr(keep),
         float(voxel_size), ptr(feats), ptr(coords), ptr(flags), ptr(ws), ws_bytes, stream())
    return feats, coords, flags


def aggregate_quantize_nus(points: torch.Tensor, sweeps: Sequence[dict], n_samples: int, voxel_size: float):
    """nuScenes multi-sweep aggregation in three fused passes (tsg_aggregate_quantize_nus).  points (sum n, c >= 5) fp32,
    sweeps: dicts(offset, count, sample, is_key, R (3,3) float64, T (3,) float64, dt) — p_key = p @ R + T.
    Returns feats (sum n, c) [x', y', z', intensity, dt, ...], coords (sum n, 4) int32, flags (sum n) uint8 (kept),
    extent (n_samples, 12) int32 view of [cur_min(4 f32 bits), ms_min(4), ms_max(4)]."""
    L.require_cuda(points)
    pts = points.float().contiguous()
    n, c = pts.shape
    arr = (L.Sweep * len(sweeps))()
    for i, f in enumerate(sweeps):
        arr[i].offset, arr[i].count, arr[i].sample, arr[i].is_key = int(f["offset"]), int(f["count"]), int(f["sample"]), int(f["is_key"])
        R = np.asarray(f["R"], np.float64).reshape(-1)
        T = np.asarray(f["T"], np.float64).reshape(-1)
        for j in range(9):
            arr[i].R[j] = float(R[j])
        for j in range(3):
            arr[i].T[j] = float(T[j])
        arr[i].dt = float(f["dt"])
    feats = torch.empty((n, c), dtype=torch.float32, device=pts.device)
    coords = torch.empty((n, 4), dtype=torch.int32, device=pts.device)
    flags = torch.empty((n,), dtype=torch.uint8, device=pts.device)
    ws_bytes = int(L.lib().tsg_aggregate_nus_ws_bytes(n_samples))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pts.device)
    call("tsg_aggregate_quantize_nus", ptr(pts), c, arr, len(sweeps), n_samples, float(voxel_size), ptr(feats), ptr(coords),
         ptr(flags), ptr(ws), ws_bytes, stream())
    extent = ws[:n_samples * 48].view(torch.int32).view(n_samples, 12)
    return feats, coords, flags, extent
