"""CPU oracle for the data side of the hot path: multi-frame aggregation + voxel quantization.

TEST INFRASTRUCTURE ONLY (see oracle/ts_oracle.py header).  R/ = /root/reference/.

`multiscan_fuse` itself cannot be imported under numpy 2 (np.bool / np.float, SURVEY §8c caveat ii),
so the FSA mask + concat are restated here; `fuse_multi_scan` IS importable and
tests/golden/make_golden.py checks this restatement against it bit for bit.
"""
from __future__ import annotations

import numpy as np

from .ts_oracle import sparse_quantize

F32 = np.float32


def fuse_multi_scan(points: np.ndarray, pose0: np.ndarray, pose: np.ndarray) -> np.ndarray:
    """R/pcseg/data/dataset/semantickitti/semantickitti_ms.py:403-417.
    p' = ((P.[p;1])_xyz - t0) . R0, fp32, every product rounded before the left-to-right sums
    (numpy broadcast-multiply then np.sum over a 4- / 3-long axis: no FMA, no pairwise tree)."""
    pts = np.asarray(points, dtype=F32)
    pose0 = np.asarray(pose0, dtype=F32)
    pose = np.asarray(pose, dtype=F32)
    h = [pts[:, 0], pts[:, 1], pts[:, 2], np.ones_like(pts[:, 0])]
    new = []
    for j in range(3):
        acc = h[0] * pose[j, 0]
        for i in range(1, 4):
            acc = acc + h[i] * pose[j, i]
        new.append(acc - pose0[j, 3])
    out = []
    for j in range(3):
        acc = new[0] * pose0[0, j]
        for i in range(1, 3):
            acc = acc + new[i] * pose0[i, j]
        out.append(acc)
    return np.concatenate([np.stack(out, 1), pts[:, 3:]], axis=1).astype(F32)


def fsa_mask(pseudo_labels: np.ndarray, delta_idx: int, flexible_steps, class_ids) -> np.ndarray:
    """semantickitti_ms.py:303-308 (Flexible Step Aggregation): keep a past point of class c iff
    step_c != 0 and |delta| % step_c == 0.  class_ids[c] is the raw label of train class c
    (LEARNING_MAP_INV[c] for KITTI, c for nuScenes, nuscenes_ms.py:329-334)."""
    m = np.zeros(len(pseudo_labels), dtype=bool)
    for c, step in enumerate(flexible_steps):
        if step == 0:
            continue
        if abs(delta_idx) % step == 0:
            m |= pseudo_labels == class_ids[c]
    return m


def aggregate_kitti(frames, poses, pseudo=None, flexible_steps=None, class_ids=None):
    """semantickitti_ms.py:140-149,253-257,263-320 (ONLY_HISTORY).
    frames[0]/poses[0] = current scan; frames[j] = scan at delta=-j... given oldest-first as the
    reference iterates delta = -MULTISCAN..-1.  Returns (raw (N0,5... ) current with time flag handled by
    caller) -> here: xyzret_ms (sum N,5) fp32 with col 4 = time flag (1 current, 0 past), n_current."""
    cur = np.asarray(frames[0], F32)
    past = []
    n_hist = len(frames) - 1
    for d in range(-n_hist, 0):                       # delta_idx ascending, like the reference loop
        j = -d
        w = fuse_multi_scan(frames[j], poses[0], poses[j])
        if flexible_steps is not None:
            w = w[fsa_mask(pseudo[j], d, flexible_steps, class_ids)]
        past.append(w)
    ms = np.concatenate([cur] + past, 0) if past else cur
    flag = np.zeros((len(ms), 1), F32)
    flag[:len(cur), 0] = 1
    return np.concatenate([ms[:, :4], flag, ms[:, 4:]], 1), len(cur)


def quat_to_mat(q) -> np.ndarray:
    """pyquaternion.Quaternion(q).rotation_matrix (absent here; nuscenes_ms.py:354-362 uses it):
    q = (w,x,y,z), normalised, standard Hamilton convention, float64."""
    w, x, y, z = np.asarray(q, np.float64) / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


def nus_RT(info0: dict, info: dict):
    """nuscenes_ms.py:348-371: R (3,3), T (3,) float64 such that p' = p @ R + T."""
    l2e_r_mat, e2g_r_mat = quat_to_mat(info0["lidar2ego_rotation"]), quat_to_mat(info0["ego2global_rotation"])
    l2e_t, e2g_t = np.asarray(info0["lidar2ego_translation"], np.float64), np.asarray(info0["ego2global_translation"], np.float64)
    l2e_r_s_mat, e2g_r_s_mat = quat_to_mat(info["lidar2ego_rotation"]), quat_to_mat(info["ego2global_rotation"])
    l2e_t_s, e2g_t_s = np.asarray(info["lidar2ego_translation"], np.float64), np.asarray(info["ego2global_translation"], np.float64)
    A = np.linalg.inv(e2g_r_mat).T @ np.linalg.inv(l2e_r_mat).T
    R = (l2e_r_s_mat.T @ e2g_r_s_mat.T) @ A
    T = (l2e_t_s @ e2g_r_s_mat.T + e2g_t_s) @ A
    T -= e2g_t @ A + l2e_t @ np.linalg.inv(l2e_r_mat).T
    return R, T


def transform_point(raw: np.ndarray, R: np.ndarray, T: np.ndarray) -> np.ndarray:
    """nuscenes_ms.py:371: raw[:, :3] = raw[:, :3] @ R + T  (fp32 points promoted to fp64, stored back fp32)."""
    out = np.asarray(raw, F32).copy()
    out[:, :3] = out[:, :3] @ R + T
    return out


def aggregate_nus(sweeps, Rs, Ts, dts):
    """nuscenes_ms.py:284-341 (LiDAR-only part): per sweep drop the ego box (|x|<1 & |y|<1.5, tested
    BEFORE the warp), write dt into col 4, warp; sweeps[0] is the key frame (identity, dt=0)."""
    out = []
    for k, s in enumerate(sweeps):
        s = np.asarray(s, F32).copy()
        no_ego = ~((np.abs(s[:, 0]) < 1.0) & (np.abs(s[:, 1]) < 1.5))
        s[:, 4] = dts[k]
        if k > 0:
            s = transform_point(s, Rs[k], Ts[k])
        out.append(s[no_ego])
    return np.concatenate(out, 0), len(out[0])


def quantize_ms(point: np.ndarray, point_ms: np.ndarray, voxel_size: float):
    """R/pcseg/data/dataset/semantickitti/semantickitti_voxel_ms.py:121-165 (eval path, no aug):
    clamp past points to the current frame's min corner, round-half-even(xyz/voxel) -> int32, shift to
    the ms min corner, dedup with sparse_quantize (first point wins, lexicographic voxel order).
    Returns dict with pc_ms (M,3), feat_ms (M,C), inds_ms, inverse_map_ms, point_ms (clamped), pc_ms_ (per point)."""
    point = np.asarray(point, F32)
    point_ms = np.asarray(point_ms, F32)
    clamp = ((point_ms[:, 0] >= point[:, 0].min()) & (point_ms[:, 1] >= point[:, 1].min())
             & (point_ms[:, 2] >= point[:, 2].min()))
    point_ms = point_ms[clamp]
    pc_ms_ = np.round(point_ms[:, :3] / F32(voxel_size)).astype(np.int32)
    pc_ms_ = pc_ms_ - pc_ms_.min(0, keepdims=True)
    _, inds, inv = sparse_quantize(pc_ms_.copy(), return_index=True, return_inverse=True)
    return dict(pc_ms=pc_ms_[inds], feat_ms=point_ms[inds], inds_ms=inds, inverse_map_ms=inv,
                point_ms=point_ms, pc_ms_=pc_ms_, clamp_mask=clamp)


def quantize_single(point: np.ndarray, voxel_size: float):
    """R/pcseg/data/dataset/semantickitti/semantickitti_voxel.py:119-133 (single-frame loader)."""
    point = np.asarray(point, F32)
    pc_ = np.round(point[:, :3] / F32(voxel_size)).astype(np.int32)
    pc_ = pc_ - pc_.min(0, keepdims=True)
    _, inds, inv = sparse_quantize(pc_, return_index=True, return_inverse=True)
    return dict(pc=pc_[inds], feat=point[inds], inds=inds, inverse_map=inv, pc_=pc_)
