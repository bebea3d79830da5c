"""Device-side data front end of the multi-frame path: raw per-frame points + poses -> collated, de-duplicated
voxels ready for MinkUNetMs, replacing the reference's CPU DataLoader work
(R/pcseg/data/dataset/semantickitti/semantickitti_ms.py:140-149,253-320 multiscan_fuse/append_time_flag,
 R/pcseg/data/dataset/semantickitti/semantickitti_voxel_ms.py:121-187 clamp/round/shift/sparse_quantize,
 :189-212 collate_batch).  Three device passes + one radix sort per BATCH instead of numpy per sample.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .tensor import SparseTensor


class MultiFrameBatch:
    """Host-side description of a batch: sample b = [current scan, history scans oldest first]."""

    def __init__(self, samples: Sequence[Sequence[np.ndarray]], poses: Sequence[Sequence[np.ndarray]]):
        self.frames: List[dict] = []
        chunks, cur_idx, self.n_cur = [], [], []
        off = 0
        for b, (frs, pss) in enumerate(zip(samples, poses)):
            order = [0] + list(range(len(frs) - 1, 0, -1))     # delta = -MULTISCAN..-1 (semantickitti_ms.py:271)
            for j in order:
                n = len(frs[j])
                self.frames.append(dict(offset=off, count=n, sample=b, is_cur=int(j == 0), pose0=pss[0], pose=pss[j]))
                chunks.append(np.asarray(frs[j], np.float32))
                if j == 0:
                    cur_idx.append(np.arange(off, off + n, dtype=np.int64))
                    self.n_cur.append(n)
                off += n
        self.n_samples = len(samples)
        self.points = np.ascontiguousarray(np.concatenate(chunks, 0))
        self.cur_idx = np.concatenate(cur_idx)
        self.total = off


def aggregate_voxelize(points: torch.Tensor, batch: MultiFrameBatch, voxel_size: float, cur_idx: torch.Tensor,
                       keep: Optional[torch.Tensor] = None):
    """points (sum n, 4) fp32 on the device (layout of `batch`).  Returns dict:
    coords (M,4) int32 [x,y,z,b] / feats (M,5) [x,y,z,i,t] of the de-duplicated voxels (lidar_ms),
    inverse (N') voxel row of every kept point, cur_rows (sum n_cur) voxel row of every current-scan point,
    point_ms (N',5), pc_ms (N',4), inds (M)."""
    feats, coords, flags, extent = ops.aggregate_quantize(points, batch.frames, batch.n_samples, voxel_size, keep)
    point_ms, pc_ms, pos, m_dev = ops.compact_rows(flags, feats, coords, want_pos=True, sync=False)
    span = (extent[:, 8:11] - extent[:, 4:7]).amax(dim=0)                  # quantised extent per axis over the batch
    m, sx, sy, sz = torch.cat([m_dev, span]).tolist()                      # the one host sync of the front end
    point_ms, pc_ms = point_ms[:m], pc_ms[:m]
    bits = [max(1, int(v).bit_length()) for v in (sx, sy, sz, batch.n_samples - 1)]
    vox, first, inverse = ops.unique_coords(pc_ms, want_index=True, want_inverse=True, field_bits=bits)
    vfeat = ops.gather_rows(point_ms, first)
    cur_rows = inverse[pos[cur_idx].long()]
    return dict(coords=vox, feats=vfeat, inverse=inverse, cur_rows=cur_rows, point_ms=point_ms, pc_ms=pc_ms, inds=first,
                pos=pos, field_bits=bits)


class NusBatch:
    """Host-side description of a nuScenes multi-sweep batch: sample b = [key sweep, sweep 1, ..., sweep S-1], each an
    (N_k, >=5) array [x, y, z, intensity, .] in its own sensor frame, with p_key = p_k @ R_k + T_k (float64) and time
    lag dt_k (nuscenes_ms.py:284-341).  One contiguous point buffer + one record per sweep, like MultiFrameBatch."""

    def __init__(self, samples, Rs, Ts, dts):
        self.sweeps: List[dict] = []
        chunks, key_idx, self.n_key = [], [], []
        off = 0
        for b, sweeps in enumerate(samples):
            for k, s in enumerate(sweeps):
                n = int(s.shape[0])
                self.sweeps.append(dict(offset=off, count=n, sample=b, is_key=int(k == 0), R=Rs[b][k], T=Ts[b][k], dt=dts[b][k]))
                chunks.append(s)
                if k == 0:
                    key_idx.append(np.arange(off, off + n, dtype=np.int64))
                    self.n_key.append(n)
                off += n
        self.n_samples = len(samples)
        self.chunks = chunks
        self.key_idx = np.concatenate(key_idx)
        self.key_sample = np.concatenate([np.full(n, b, np.int64) for b, n in enumerate(self.n_key)])
        self.total = off

    def points(self, device="cuda") -> torch.Tensor:
        """The batch's points as one (total, c) fp32 device tensor (chunks may be numpy arrays or device tensors)."""
        parts = [c if isinstance(c, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(c, np.float32)).to(device) for c in self.chunks]
        return torch.cat([p.float() for p in parts], 0).contiguous()


def aggregate_voxelize_nus(samples: Sequence[Sequence[torch.Tensor]], Rs: Sequence[Sequence[np.ndarray]],
                           Ts: Sequence[Sequence[np.ndarray]], dts: Sequence[Sequence[float]], voxel_size: float,
                           batch: Optional[NusBatch] = None, points: Optional[torch.Tensor] = None):
    """nuScenes multi-sweep front end (BASELINE configs[3]) on the device, for a batch of samples.
    samples[b][k] (N_k, 5) fp32 device tensor [x, y, z, intensity, .] in sweep k's own frame (k = 0: key frame);
    Rs/Ts[b][k] float64 with p_key = p_k @ R + T; dts[b][k] seconds.  Per sweep, as
    R/pcseg/data/dataset/nuscenes/nuscenes_ms.py:284-341: the ego box |x| < 1 & |y| < 1.5 is tested on the RAW points,
    dt goes to column 4, the sweep is warped in fp64 (`transform_point`, :348-373), then filtered.  Then the loader's
    clamp / round / shift / sparse_quantize (nuscenes_voxel_ms.py:122-160 = the SemanticKITTI one) per sample, collated.
    All of it is three fused passes over the batch (tsg_aggregate_quantize_nus), one compaction and one radix sort — no
    per-sweep Python loop, one host read-back (kept points, key points per sample, voxel extent) + the voxel count.
    Returns the same dict as aggregate_voxelize (coords (M,4) [x,y,z,b], feats (M,5), inverse, cur_rows, point_ms, pc_ms, inds)
    plus n_cur (key-frame points per sample that survive the ego-box filter)."""
    if batch is None:
        batch = NusBatch(samples, Rs, Ts, dts)
    if points is None:
        points = batch.points()
    dev = points.device
    feats, coords, flags, extent = ops.aggregate_quantize_nus(points, batch.sweeps, batch.n_samples, voxel_size)
    point_ms, pc_ms, pos, m_dev = ops.compact_rows(flags, feats, coords, want_pos=True, sync=False)
    key_pos = pos[torch.from_numpy(batch.key_idx).to(dev)]                     # -1 for key points inside the ego box
    per_sample = torch.zeros(batch.n_samples, dtype=torch.int32, device=dev)
    per_sample.index_add_(0, torch.from_numpy(batch.key_sample).to(dev), (key_pos >= 0).to(torch.int32))
    span = (extent[:, 8:11] - extent[:, 4:7]).amax(dim=0)
    host = torch.cat([m_dev, span, per_sample]).tolist()                       # the one host sync of the front end
    m, (sx, sy, sz), n_cur = host[0], host[1:4], host[4:]
    point_ms, pc_ms = point_ms[:m], pc_ms[:m]
    bits = [max(1, int(v).bit_length()) for v in (sx, sy, sz, batch.n_samples - 1)]
    vox, first, inverse = ops.unique_coords(pc_ms, want_index=True, want_inverse=True, field_bits=bits)
    vfeat = ops.gather_rows(point_ms, first)
    cur_rows = inverse[key_pos[key_pos >= 0].long()]
    return dict(coords=vox, feats=vfeat, inverse=inverse, cur_rows=cur_rows, point_ms=point_ms, pc_ms=pc_ms, inds=first,
                n_cur=[int(v) for v in n_cur], field_bits=bits)


def as_lidar_ms(out: dict) -> SparseTensor:
    return SparseTensor(out["feats"], out["coords"], 1)
