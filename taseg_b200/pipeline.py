"""Sync-free, graph-captured inference pipeline: raw multi-frame points -> per-point logits with NO
host read-back of a data-dependent size and, after capture, ONE graph launch per batch.

The reference's forward (R/pcseg/model/segmentor/voxel/minkunet/minkunet_ms.py:385-420 over torchsparse) returns to the
host for every data-dependent size: `np.unique` in sparse_quantize, `torch.unique` + `.item()` in spdownsample
(TS/nn/functional/downsample.py:47-52), `nonzero` / `sum` in the kernel-map builder (TS/nn/functional/conv.py:156-176).
Here every such size stays in a device counter:
  * buffers and grids are sized by CAPACITIES learnt from one eager calibration pass (x margin, rounded to 256 rows);
  * kernels process min(*counter, capacity) rows (`*_dev` entry points of include/taseg_b200.h);
  * producers clamp their counters to the consumer's capacity and raise a device status word on overflow, which the host
    inspects together with the result (one 4-byte copy) — an overflowing batch is re-run eagerly and re-calibrates.
With fixed addresses, grids and launch parameters the whole forward (about 250 kernels) is captured once into a CUDA
graph; replaying it removes the Python / ctypes launch path and the five size read-backs from the step.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import frontend, ops
from .engine import Engine, Level
from .nn.utils.kernel import kernel_offsets_np

OVERFLOW = 2      # status bit: a device counter exceeded the capacity of its consumer
RANGE = 1         # status bit: coordinate outside the packable range / the promised field bits


def round_cap(n: int, margin: float, quantum: int = 256) -> int:
    return max(quantum, (int(n * margin) + quantum - 1) // quantum * quantum)


class GeometryDev:
    """Voxel pyramid whose row counts live on the device: level l has capacity caps[l] and counter counts[l] (int32 views
    of one tensor).  Same attributes as engine.Geometry's levels, with the mask-sorted conv maps prebuilt."""

    def __init__(self, coords0: torch.Tensor, counters: Sequence[torch.Tensor], caps: Sequence[int], field_bits, status):
        assert coords0.shape[0] == caps[0]
        self.levels: List[Level] = []
        n_levels = len(caps)
        c = coords0
        coords = [c]
        for l in range(n_levels - 1):      # the down-sampling chain: unique((coords >> 1) << 1) level by level
            c = ops.unique_coords_dev(c, counters[l], caps[l + 1], counters[l + 1], status, trunc_stride=2 ** (l + 1),
                                      field_bits=field_bits)
            coords.append(c)
        for l in range(n_levels):
            lv = Level()
            lv.stride, lv.coords, lv.n, lv.n_dev = 2 ** l, coords[l], caps[l], counters[l]
            lv.km3 = lv.km2 = lv.m2 = lv.m2t = None
            lv.table = ops.table_from_coords_dev(coords[l], counters[l], status)
            nbr3, keys3 = ops.build_kmap_dev(lv.table, coords[l], counters[l], kernel_offsets_np(3, lv.stride), want_keys=True)
            lv.m3 = ops.kmap_sort_rows_dev(nbr3, counters[l], row_keys=keys3)
            if ops.SplitItems.wanted(caps[l], 27):      # few tiles per SM: K-split work items balance the launch
                lv.m3 = lv.m3 + (ops.SplitItems(lv.m3[1], caps[l], 27, n_dev=counters[l]),)
            if l + 1 < n_levels:
                nbr2, keys2 = ops.build_kmap_dev(lv.table, coords[l + 1], counters[l + 1], kernel_offsets_np(2, lv.stride),
                                                 want_keys=True)
                lv.m2 = ops.kmap_sort_rows_dev(nbr2, counters[l + 1], row_keys=keys2)
                nbr2t = ops.kmap_transpose_dev(nbr2, counters[l + 1], caps[l])
                lv.m2t = ops.kmap_sort_rows_dev(nbr2t, counters[l])
            self.levels.append(lv)


class Pipeline:
    """One batch SHAPE (frame layout of a frontend.MultiFrameBatch) of the multi-frame voxel backbone, captured once.

        pipe = Pipeline(engine, mfb, voxel_size)
        pipe.calibrate(points)            # one eager pass: learns capacities and key widths
        pipe.capture()                    # optional: CUDA graph
        logits = pipe(points)             # (sum n_cur, num_class) fp32, a STATIC buffer overwritten by the next call

    `points` (sum n, 4) fp32 device tensor in the layout of `mfb`; poses are taken from `mfb` (update with set_poses)."""

    N_COUNTERS = 8     # [kept points, n_0 .. n_4, spare, spare]

    def __init__(self, engine: Engine, mfb: frontend.MultiFrameBatch, voxel_size: float, keep: Optional[torch.Tensor] = None,
                 margin: float = 1.25, n_levels: int = 5):
        assert not engine.spv and not engine.voxelize_input, "the captured pipeline covers the voxel backbones fed by the device front end"
        self.engine, self.mfb, self.voxel, self.margin, self.n_levels = engine, mfb, float(voxel_size), margin, n_levels
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.n_raw = int(mfb.total)
        self.n_cur = int(sum(mfb.n_cur))
        self.points = torch.empty((self.n_raw, 4), dtype=torch.float32, device=dev)      # static input buffer
        self.cur_idx = torch.from_numpy(mfb.cur_idx).to(dev)
        self.keep = keep
        self.max_count = max(int(f["count"]) for f in mfb.frames)
        self.frames_host = torch.empty(ctypes.sizeof(L.Frame) * len(mfb.frames), dtype=torch.uint8).pin_memory()
        self.frames_dev = torch.empty_like(self.frames_host, device=dev)
        self.set_poses(mfb.frames)
        self.counters = torch.zeros(self.N_COUNTERS, dtype=torch.int32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sched = torch.zeros(2, dtype=torch.int32, device=dev)     # this pipeline's dynamic-tile-scheduler counters
        self.caps: Optional[List[int]] = None
        self.cap_pts = self.n_raw
        self.field_bits = None
        self.graph = None
        self.logits = None
        self.sizes = None

    # -- inputs --------------------------------------------------------------------------------------------------
    def set_poses(self, frames: Sequence[dict]) -> None:
        """Refresh the device frame table (poses change every step; offsets / counts are the captured shape)."""
        arr = ops.frames_to_array(frames)
        ctypes.memmove(self.frames_host.data_ptr(), ctypes.addressof(arr), ctypes.sizeof(arr))
        self.frames_dev.copy_(self.frames_host, non_blocking=True)

    # -- calibration ---------------------------------------------------------------------------------------------
    def calibrate(self, points: torch.Tensor) -> dict:
        """One eager (synchronising) pass over a representative batch: voxel counts per level and key widths."""
        out = frontend.aggregate_voxelize(points, self.mfb, self.voxel, self.cur_idx, keep=self.keep)
        sizes = [int(out["coords"].shape[0])]
        c = out["coords"]
        for l in range(self.n_levels - 1):
            c = ops.unique_coords(c, trunc_stride=2 ** (l + 1), field_bits=out["field_bits"])
            sizes.append(int(c.shape[0]))
        self.sizes = dict(kept=int(out["point_ms"].shape[0]), levels=sizes)
        self.caps = [min(self.n_raw, round_cap(sizes[0], self.margin))]
        for l in range(1, self.n_levels):
            self.caps.append(min(self.caps[l - 1], round_cap(sizes[l], self.margin)))
        self.field_bits = list(out["field_bits"])
        self.graph = None
        return self.sizes

    # -- the sync-free forward -------------------------------------------------------------------------------------
    def _forward(self) -> torch.Tensor:
        prev, ops.SCHED_OVERRIDE = ops.SCHED_OVERRIDE, self.sched
        try:
            return self._forward_impl()
        finally:
            ops.SCHED_OVERRIDE = prev

    def _forward_impl(self) -> torch.Tensor:
        cnt = self.counters
        ctr = [cnt[i:i + 1] for i in range(self.N_COUNTERS)]
        self.status.zero_()
        feats, coords, flags = ops.aggregate_quantize_dev(self.points, self.frames_dev, len(self.mfb.frames), self.max_count,
                                                          self.mfb.n_samples, self.voxel, self.keep)
        point_ms, pc_ms, pos, _ = ops.compact_rows(flags, feats, coords, want_pos=True, sync=False, m_dev=ctr[0])
        vox, first, inverse = ops.unique_coords_dev(pc_ms, ctr[0], self.caps[0], ctr[1], self.status,
                                                    field_bits=self.field_bits, want_index=True, want_inverse=True)
        vfeat = ops.gather_rows_dev(point_ms, first, ctr[1])
        cur_rows = inverse[pos[self.cur_idx].long()]
        geo = GeometryDev(vox, ctr[1:1 + self.n_levels], self.caps, self.field_bits, self.status)
        x_f = vfeat[:, :self.engine.in_dim].contiguous()
        return self.engine.run(geo, x_f, dict(zc=vox.float()), out_rows=cur_rows)

    @torch.no_grad()
    def capture(self, warmup: int = 2) -> None:
        """Capture the forward into a CUDA graph (static input buffer `points`, static output `logits`)."""
        assert self.caps is not None, "calibrate() first"
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):        # warm the allocator and the lazily initialised kernels outside the capture
                self._forward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.logits = self._forward()
        self.graph = g

    @torch.no_grad()
    def __call__(self, points: Optional[torch.Tensor] = None) -> torch.Tensor:
        """points None: the caller has already filled `self.points` (e.g. with its own host-to-device copy)."""
        assert self.caps is not None, "calibrate() first"
        if points is not None:
            self.points.copy_(points, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.logits = self._forward()
        return self.logits

    def check(self) -> int:
        """Status word of the last forward (synchronises): 0, or RANGE / OVERFLOW bits — the batch must then be re-run
        through the eager path (frontend.aggregate_voxelize + Engine) and the pipeline re-calibrated."""
        return int(self.status.item())

    def level_counts(self) -> List[int]:
        """Device counters of the last forward: [kept points, n_0, ..., n_4] (synchronises; for tests and reports)."""
        return self.counters[:1 + self.n_levels].tolist()
