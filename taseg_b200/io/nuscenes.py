"""nuScenes sweep metadata -> the inputs of `frontend.aggregate_voxelize_nus`.

Host-side restatement of what R/pcseg/data/dataset/nuscenes/nuscenes_ms.py does with the sweep info dictionaries
(`nusc_infos_sweep` entries: `lidar2ego_rotation/translation`, `ego2global_rotation/translation`, `timestamp`,
`lidar_path` for key frames or `data_path` + `sensor2lidar_*` for intermediate sweeps):
  * `sweep_transform(info0, info)`: R (3,3), T (3,) float64 with p_key = p_sweep @ R + T          (:348-371)
  * `sweep_dt(info0, info)`: the value written to column 4                                      (:289, :315)
  * `select_sweeps(...)`: the distance-based choice of `multiscan` history sweeps `step` metres apart plus every key
    frame on the way                                                                            (:237-276)
  * `read_sweep(path)`: float32 x 5 [x, y, z, intensity, ring/time]                              (:287, :309)
pyquaternion is not a dependency: `quaternion_rotation_matrix` is its `Quaternion(q).rotation_matrix` (unit-normalised
Hamilton convention, q = (w, x, y, z)).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def quaternion_rotation_matrix(q: Sequence[float]) -> np.ndarray:
    w, x, y, z = np.asarray(q, np.float64) / np.linalg.norm(np.asarray(q, np.float64))
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


def sweep_transform(info0: dict, info: dict) -> Tuple[np.ndarray, np.ndarray]:
    """Sweep `info`'s lidar frame -> key frame `info0`'s lidar frame, through ego and global poses."""
    l2e_r_mat = quaternion_rotation_matrix(info0["lidar2ego_rotation"])
    e2g_r_mat = quaternion_rotation_matrix(info0["ego2global_rotation"])
    l2e_t, e2g_t = np.asarray(info0["lidar2ego_translation"], np.float64), np.asarray(info0["ego2global_translation"], np.float64)
    l2e_r_s_mat = quaternion_rotation_matrix(info["lidar2ego_rotation"])
    e2g_r_s_mat = quaternion_rotation_matrix(info["ego2global_rotation"])
    l2e_t_s, e2g_t_s = np.asarray(info["lidar2ego_translation"], np.float64), np.asarray(info["ego2global_translation"], np.float64)
    back = np.linalg.inv(e2g_r_mat).T @ np.linalg.inv(l2e_r_mat).T
    R = (l2e_r_s_mat.T @ e2g_r_s_mat.T) @ back
    T = (l2e_t_s @ e2g_r_s_mat.T + e2g_t_s) @ back
    T -= e2g_t @ back + l2e_t @ np.linalg.inv(l2e_r_mat).T
    return R, T


def sweep_dt(info0: dict, info: dict) -> float:
    return info0["timestamp"] / 1e6 - info["timestamp"] / 1e6


def read_sweep(path: str) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.float32, count=-1)
    if raw.size % 5:
        raise ValueError(f"{path}: size is not a multiple of 5 float32 values")
    return raw.reshape([-1, 5])


def select_sweeps(dists: Sequence[float], is_key_frame: Sequence[bool], multiscan: int, step: float) -> List[int]:
    """Which history sweeps to aggregate.  dists[i] = planar distance of sweep delta = -(i+1) from the key frame, in
    walking order, ending with the first entry beyond multiscan*step (or 1000 at a scene boundary); is_key_frame[i] for
    the same sweeps (entries that exist).  Returns sorted negative deltas."""
    n = len(is_key_frame)
    cur_scan, chosen = 1, []
    for idx in range(n):
        d = dists[idx]
        if d - cur_scan * step > 0 or (d < dists[idx + 1] and abs(d - cur_scan * step) < abs(dists[idx + 1] - cur_scan * step)):
            chosen.append(-(idx + 1))
            cur_scan += 1
        if cur_scan > multiscan:
            break
    chosen += [-(i + 1) for i in range(n) if is_key_frame[i]]
    return sorted(set(chosen))
