"""SemanticKITTI on-disk formats -> the inputs of `frontend.MultiFrameBatch`.

What the reference reads (`R/` = /root/reference):
  * `velodyne/NNNNNN.bin`   float32 x 4 [x, y, z, intensity]         (R/pcseg/data/dataset/semantickitti/semantickitti_ms.py:121)
  * `labels/NNNNNN.label`   uint32, semantic id = value & 0xFFFF      (semantickitti_ms.py:133-138)
  * `calib.txt`             `key: 12 floats` rows of 3x4 matrices     (semantickitti_ms.py:349-373)
  * `poses.txt`             12 floats per scan; pose = Tr^-1 . T . Tr in float64, then cast to float32
                            (semantickitti_ms.py:375-401, :346)
  * pseudo-label dumps written by R/train.py:503-508 use the `.label` format (uint32), read the same way.
The history of scan n is the scans n-MULTISCAN .. n-1 that exist (semantickitti_ms.py:271-281, `only_history`),
each with the FSA keep mask `OR_c(|delta| % step_c == 0 and pseudo == c)` (semantickitti_ms.py:303-308).
File I/O stays on the host; everything after it (warp, time flag, clamp, quantise, dedup) is the device front end.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def read_scan(path: str) -> np.ndarray:
    """(N, 4) float32 [x, y, z, intensity]."""
    raw = np.fromfile(path, dtype=np.float32)
    if raw.size % 4:
        raise ValueError(f"{path}: size is not a multiple of 4 float32 values")
    return raw.reshape((-1, 4))


def read_labels(path: str, learning_map: Optional[Dict[int, int]] = None) -> np.ndarray:
    """(N,) int64 semantic ids (`value & 0xFFFF`), optionally passed through the dataset's learning map."""
    sem = (np.fromfile(path, dtype=np.uint32) & 0xFFFF).astype(np.int64)
    if learning_map is not None:
        lut = np.zeros(max(max(learning_map) + 1, int(sem.max(initial=0)) + 1), dtype=np.int64)
        for k, v in learning_map.items():
            lut[k] = v
        sem = lut[sem]
    return sem


def write_labels(path: str, labels: np.ndarray) -> None:
    """The format of the reference's prediction / pseudo-label dumps (R/train.py:503-508)."""
    np.asarray(labels).astype(np.uint32).tofile(path)


def _row_to_pose(values: Sequence[float]) -> np.ndarray:
    pose = np.zeros((4, 4))
    pose[0, 0:4] = values[0:4]
    pose[1, 0:4] = values[4:8]
    pose[2, 0:4] = values[8:12]
    pose[3, 3] = 1.0
    return pose


def parse_calibration(path: str) -> Dict[str, np.ndarray]:
    calib = {}
    with open(path) as f:
        for line in f:
            if ":" not in line:
                continue
            key, content = line.strip().split(":")
            calib[key] = _row_to_pose([float(v) for v in content.strip().split()])
    return calib


def parse_poses(path: str, calibration: Dict[str, np.ndarray]) -> List[np.ndarray]:
    """Per-scan sensor poses as float32 4x4 (float64 product, then the cast of semantickitti_ms.py:346)."""
    tr = calibration["Tr"]
    tr_inv = np.linalg.inv(tr)
    poses = []
    with open(path) as f:
        for line in f:
            values = [float(v) for v in line.strip().split()]
            if len(values) < 12:
                continue
            poses.append(np.matmul(tr_inv, np.matmul(_row_to_pose(values), tr)).astype(np.float32))
    return poses


class KittiSequence:
    """One `sequences/SS` directory."""

    def __init__(self, seq_dir: str):
        self.dir = seq_dir
        self.calibration = parse_calibration(os.path.join(seq_dir, "calib.txt"))
        self.poses = parse_poses(os.path.join(seq_dir, "poses.txt"), self.calibration)
        vel = os.path.join(seq_dir, "velodyne")
        self.scan_ids = sorted(int(f[:-4]) for f in os.listdir(vel) if f.endswith(".bin"))

    def scan_path(self, n: int) -> str:
        return os.path.join(self.dir, "velodyne", "%06d.bin" % n)

    def label_path(self, n: int, folder: str = "labels") -> str:
        return os.path.join(self.dir, folder, "%06d.label" % n)


def load_sample(seq: KittiSequence, n: int, multiscan: int, flexible_steps: Optional[Sequence[int]] = None,
                pseudo_folder: Optional[str] = None, learning_map_inv: Optional[Dict[int, int]] = None
                ) -> Tuple[List[np.ndarray], List[np.ndarray], Optional[np.ndarray]]:
    """Scan n with its history, in the layout `MultiFrameBatch` expects: frames[0] = current scan, frames[j] = scan n - j
    (the batch object replays them oldest first, the order of semantickitti_ms.py:271), poses alike, and — with
    `flexible_steps` — the FSA keep mask over [current, history oldest first] (current scan all ones)."""
    frames, poses = [read_scan(seq.scan_path(n))], [seq.poses[n]]
    deltas = []
    for j in range(1, multiscan + 1):
        m = n - j
        if m < 0 or m >= len(seq.poses) or not os.path.exists(seq.scan_path(m)):
            continue                                      # the reference's try/except: a missing scan is skipped
        frames.append(read_scan(seq.scan_path(m)))
        poses.append(seq.poses[m])
        deltas.append(j)
    keep = None
    if flexible_steps is not None:
        assert pseudo_folder is not None and learning_map_inv is not None
        parts = [np.ones(len(frames[0]), np.uint8)]
        for idx in range(len(deltas) - 1, -1, -1):        # oldest first
            j = deltas[idx]
            pseudo = read_labels(seq.label_path(n - j, pseudo_folder))
            mask = np.zeros(len(pseudo), dtype=bool)
            for class_idx, class_step in enumerate(flexible_steps):
                if class_step == 0:
                    continue
                if j % class_step == 0:
                    mask |= pseudo == learning_map_inv[class_idx]
            parts.append(mask.astype(np.uint8))
        keep = np.concatenate(parts)
    return frames, poses, keep
