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


def aggregate_voxelize_nus(samples: Sequence[Sequence[torch.Tensor]], Rs: Sequence[Sequence[np.ndarray]],
                           Ts: Sequence[Sequence[np.ndarray]], dts: Sequence[Sequence[float]], voxel_size: float):
    """nuScenes multi-sweep front end (BASELINE configs[3]) on the device, for a batch of samples.
    samples[b][k] (N_k, 5) fp32 device tensor [x, y, z, intensity, .] in sweep k's own frame (k = 0: key frame);
    Rs/Ts[b][k] float64 with p_key = p_k @ R + T; dts[b][k] seconds.  Per sweep, as
    R/pcseg/data/dataset/nuscenes/nuscenes_ms.py:284-341: the ego box |x| < 1 & |y| < 1.5 is tested on the RAW points,
    dt goes to column 4, the sweep is warped in fp64 (`transform_point`, :348-373), then filtered.  Then the loader's
    clamp / round / shift / sparse_quantize (nuscenes_voxel_ms.py:122-160 = the SemanticKITTI one) per sample, collated.
    Returns the same dict as aggregate_voxelize (coords (M,4) [x,y,z,b], feats (M,5), inverse, cur_rows, point_ms, pc_ms, inds)."""
    pts_all, pc_all, n_cur, cur_off = [], [], [], []
    off = 0
    for b, sweeps in enumerate(samples):
        parts = []
        for k, s in enumerate(sweeps):
            s = s.float()
            no_ego = ~((s[:, 0].abs() < 1.0) & (s[:, 1].abs() < 1.5))
            s = s.clone()
            s[:, 4] = float(dts[b][k])
            if k > 0:
                s = ops.transform_point(s, Rs[b][k], Ts[b][k])
            parts.append(s[no_ego])
        cur = parts[0]
        ms = torch.cat(parts, 0)
        mn = cur[:, :3].amin(dim=0)
        ms = ms[(ms[:, 0] >= mn[0]) & (ms[:, 1] >= mn[1]) & (ms[:, 2] >= mn[2])]      # clamp to the key frame's min corner
        pc = torch.round(ms[:, :3] / voxel_size).to(torch.int32)                       # fp32 divide, round half to even
        pc = pc - pc.amin(dim=0, keepdim=True)
        pts_all.append(ms)
        pc_all.append(torch.cat([pc, torch.full((pc.shape[0], 1), b, dtype=torch.int32, device=pc.device)], 1))
        n_cur.append(int(cur.shape[0]))                                                # key-frame points come first and all survive the clamp
        cur_off.append(off)
        off += ms.shape[0]
    point_ms, pc_ms = torch.cat(pts_all, 0).contiguous(), torch.cat(pc_all, 0).contiguous()
    vox, first, inverse = ops.unique_coords(pc_ms, want_index=True, want_inverse=True)
    vfeat = ops.gather_rows(point_ms, first)
    cur_idx = torch.cat([torch.arange(o, o + n, device=pc_ms.device) for o, n in zip(cur_off, n_cur)])
    return dict(coords=vox, feats=vfeat, inverse=inverse, cur_rows=inverse[cur_idx], point_ms=point_ms, pc_ms=pc_ms, inds=first,
                n_cur=n_cur)


def as_lidar_ms(out: dict) -> SparseTensor:
    return SparseTensor(out["feats"], out["coords"], 1)
