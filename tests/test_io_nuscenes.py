"""nuScenes sweep metadata handling against the reference's own `transform_point` source (build container) and the data
oracle."""
import os

import numpy as np
import pytest

from oracle import data_oracle as D
from taseg_b200.io import quaternion_rotation_matrix, select_sweeps, sweep_dt, sweep_transform

REF = "/root/reference/pcseg/data/dataset/nuscenes/nuscenes_ms.py"


def _info(rng, ts):
    q = lambda: (rng.normal(size=4) * [1.0, 0.05, 0.05, 0.3]).tolist()
    return dict(lidar2ego_rotation=q(), lidar2ego_translation=(rng.normal(size=3) * [1, 0.1, 1.8]).tolist(),
                ego2global_rotation=q(), ego2global_translation=(rng.normal(size=3) * 300).tolist(), timestamp=ts)


def test_rotation_matrix_is_proper():
    rng = np.random.default_rng(0)
    for _ in range(10):
        m = quaternion_rotation_matrix(rng.normal(size=4))
        assert np.allclose(m @ m.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(m), 1.0)
    assert np.array_equal(quaternion_rotation_matrix([1, 0, 0, 0]), np.eye(3))
    assert np.allclose(quaternion_rotation_matrix([np.cos(0.2), 0, 0, np.sin(0.2)]),
                       [[np.cos(0.4), -np.sin(0.4), 0], [np.sin(0.4), np.cos(0.4), 0], [0, 0, 1]])


def test_sweep_transform_matches_oracle_and_geometry():
    rng = np.random.default_rng(1)
    info0, info = _info(rng, 1.6e15), _info(rng, 1.6e15 - 150000.0)
    R, T = sweep_transform(info0, info)
    R2, T2 = D.nus_RT(info0, info)
    assert np.array_equal(R, R2) and np.array_equal(T, T2)
    assert np.isclose(sweep_dt(info0, info), 0.15)
    # geometry: a point of the sweep, taken to global coordinates and back into the key frame's lidar frame
    p = rng.normal(size=(5, 3)) * 20
    ls, es = quaternion_rotation_matrix(info["lidar2ego_rotation"]), quaternion_rotation_matrix(info["ego2global_rotation"])
    l0, e0 = quaternion_rotation_matrix(info0["lidar2ego_rotation"]), quaternion_rotation_matrix(info0["ego2global_rotation"])
    g = (p @ ls.T + info["lidar2ego_translation"]) @ es.T + info["ego2global_translation"]
    back = ((g - info0["ego2global_translation"]) @ e0 - info0["lidar2ego_translation"]) @ l0
    assert np.allclose(p @ R + T, back, atol=1e-8)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree only exists in the build container")
def test_sweep_transform_matches_reference_source():
    src = open(REF).read()
    a = src.index("    def transform_point(self, raw_data, info0, info):")
    b = src.index("    @staticmethod", a)

    class Quaternion:          # pyquaternion is not installed: same rotation_matrix, so the test pins the pose algebra
        def __init__(self, q):
            self.rotation_matrix = quaternion_rotation_matrix(q)
    ns = {"np": np, "Quaternion": Quaternion}
    exec("class _P:\n" + src[a:b], ns)
    rng = np.random.default_rng(2)
    info0, info = _info(rng, 1.0), _info(rng, 0.5)
    pts = (rng.normal(size=(100, 5)) * 20).astype(np.float32)
    want = ns["_P"]().transform_point(pts.copy(), info0, info)
    R, T = sweep_transform(info0, info)
    assert np.array_equal(D.transform_point(pts, R, T), want)


def test_select_sweeps():
    # sweeps 0.6 m apart, want 3 history scans 1.0 m apart: the closest sweep to 1, 2, 3 m each, plus key frames
    dists = [0.6, 1.2, 1.8, 2.4, 3.0, 3.6]
    key = [False, False, True, False, False]
    assert select_sweeps(dists, key, multiscan=3, step=1.0) == [-5, -3, -2]
    assert select_sweeps([1000], [], 3, 1.0) == []                       # scene boundary right behind the key frame
    assert select_sweeps([0.2, 1000], [False], 2, 0.5) == [-1]           # closer to the 0.5 m target than the boundary entry
    assert select_sweeps([0.7, 1000], [False], 2, 0.5) == [-1]           # already beyond the target
    assert select_sweeps([0.1, 0.2, 0.3, 1000], [False, True, False], 1, 5.0) == [-3, -2]   # last sweep before the boundary + key frame
