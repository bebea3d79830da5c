"""Seeded synthetic LiDAR worlds for benchmarks and tests (SURVEY.md §8d "synthetic inputs").

There is no dataset on the benchmark box, so inputs are ray-cast: a spinning multi-ring sensor in
a world made of a ground plane and seeded axis-aligned boxes.  Ring structure matters — i.i.d. random
points give ~2.3 kernel-map pairs per voxel instead of the 6-10 a real scan has, which would
mis-state the convolution work.  Multi-frame samples re-cast the SAME world from each past ego pose so
that warping history into the current frame (fuse_multi_scan) makes static structure overlap.

Pure numpy, host side, input generation only — not a fallback for any device op.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

F32 = np.float32


@dataclass(frozen=True)
class SensorSpec:
    rings: int
    elev_lo_deg: float
    elev_hi_deg: float
    azimuth_bins: int
    height: float
    max_range: float


KITTI = SensorSpec(64, -24.8, 2.0, 2048, 1.73, 80.0)
NUSCENES = SensorSpec(32, -30.67, 10.67, 1090, 1.84, 70.0)


def make_world(rng: np.random.Generator, n_boxes: int = 60):
    """Boxes (n,6) = [xmin,ymin,zmin,xmax,ymax,zmax] in the world frame of the current pose (sensor at origin)."""
    c = rng.uniform(-50, 50, (n_boxes, 2))
    fp = rng.uniform(1.5, 12, (n_boxes, 2))
    h = rng.uniform(1.4, 6, n_boxes)
    keep = np.linalg.norm(c, axis=1) > 4.0 + 0.5 * np.linalg.norm(fp, axis=1)
    c, fp, h = c[keep], fp[keep], h[keep]
    return c, fp, h


def yaw_pose(x: float, y: float, yaw_deg: float) -> np.ndarray:
    """4x4 world-from-sensor matrix (float64)."""
    a = np.deg2rad(yaw_deg)
    P = np.eye(4)
    P[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    P[0, 3], P[1, 3] = x, y
    return P


def cast_scan(spec: SensorSpec, world, pose: np.ndarray, rng: np.random.Generator, n_feat: int = 4) -> np.ndarray:
    """One sweep as seen from `pose` (world-from-sensor); returns (N, n_feat) fp32 in the SENSOR frame,
    columns x,y,z,intensity[,0...]."""
    c, fp, h = world
    elev = np.deg2rad(np.linspace(spec.elev_lo_deg, spec.elev_hi_deg, spec.rings))
    az_idx = np.arange(spec.azimuth_bins)[None, :] + rng.uniform(-0.5, 0.5, (spec.rings, spec.azimuth_bins))
    az = az_idx * (2 * np.pi / spec.azimuth_bins)
    ce = np.cos(elev)[:, None]
    d_s = np.stack([ce * np.cos(az), ce * np.sin(az), np.broadcast_to(np.sin(elev)[:, None], az.shape)], -1).reshape(-1, 3)
    R, o = pose[:3, :3], pose[:3, 3]
    d = d_s @ R.T
    zg = -spec.height
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(d[:, 2] < -1e-6, (zg - o[2]) / d[:, 2], np.inf)
        lo = np.concatenate([c - fp / 2, np.full((len(c), 1), zg)], 1)
        hi = np.concatenate([c + fp / 2, (zg + h)[:, None]], 1)
        inv = 1.0 / d
        for b in range(len(c)):
            t0 = (lo[b] - o) * inv
            t1 = (hi[b] - o) * inv
            tn = np.minimum(t0, t1).max(1)
            tf = np.maximum(t0, t1).min(1)
            hit = (tn <= tf) & (tn > 0)
            t = np.where(hit & (tn < t), tn, t)
    ok = np.isfinite(t) & (t < spec.max_range)
    ok &= rng.uniform(size=t.shape) >= 0.08
    t = t[ok] + rng.normal(0, 0.01, int(ok.sum()))
    pts = d_s[ok] * t[:, None]
    out = np.zeros((len(pts), n_feat), F32)
    out[:, :3] = pts
    out[:, 3] = rng.uniform(0, 1, len(pts))
    return out


def kitti_sample(seed: int, n_frames: int = 1, spec: SensorSpec = KITTI, n_boxes: int = 60):
    """frames[0] = current scan, frames[j] = scan j steps in the past ((N_j,4) fp32, sensor frame);
    poses[j] = float32 4x4 world-from-sensor (ego moved 1.2 m/frame along +x, yaw +0.5 deg/frame)."""
    rng = np.random.default_rng(seed)
    world = make_world(rng, n_boxes)
    frames, poses = [], []
    for j in range(n_frames):
        P = yaw_pose(-1.2 * j, 0.0, -0.5 * j)
        frames.append(cast_scan(spec, world, P, np.random.default_rng([seed, j])))
        poses.append(P.astype(F32))
    return frames, poses


def nus_sample(seed: int, n_sweeps: int = 10, spec: SensorSpec = NUSCENES, n_boxes: int = 60):
    """sweeps[k] (N_k,5) fp32 [x,y,z,intensity,0] in sweep k's frame (k=0 key frame); Rs/Ts float64 with
    p_key = p_k @ R_k + T_k (the form nuscenes_ms.py:371 applies); dts[k] = 0.05*k seconds."""
    rng = np.random.default_rng(seed)
    world = make_world(rng, n_boxes)
    sweeps, Rs, Ts, dts = [], [], [], []
    for k in range(n_sweeps):
        P = yaw_pose(-0.5 * k, 0.0, -0.2 * k)
        sweeps.append(cast_scan(spec, world, P, np.random.default_rng([seed, k]), n_feat=5))
        Rs.append(P[:3, :3].T.copy())
        Ts.append(P[:3, 3].copy())
        dts.append(0.05 * k)
    return sweeps, Rs, Ts, dts


def sector(points: np.ndarray, frac: float) -> np.ndarray:
    """Keep the azimuth sector [0, 2*pi*frac) of a scan — the bounded sample the CPU baseline times."""
    a = np.arctan2(points[:, 1], points[:, 0]) % (2 * np.pi)
    return points[a < 2 * np.pi * frac]
