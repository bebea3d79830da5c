"""SemanticKITTI on-disk formats: our readers against the reference's own parsing code where
/root/reference exists (build container), and against hand-built files everywhere."""
import os
import sys
import types

import numpy as np
import pytest

from taseg_b200.io import KittiSequence, load_sample, parse_calibration, parse_poses, read_labels, read_scan, write_labels

REF = "/root/reference"


def _make_sequence(root, n_scans=5, seed=0):
    rng = np.random.default_rng(seed)
    seq = os.path.join(root, "sequences", "00")
    os.makedirs(os.path.join(seq, "velodyne"))
    os.makedirs(os.path.join(seq, "labels"))
    os.makedirs(os.path.join(seq, "predictions"))
    tr = np.array([[4.27680239e-04, -9.99967248e-01, -8.08449168e-03, -1.19845993e-02],
                   [-7.21062651e-03, 8.08119847e-03, -9.99941316e-01, -5.40398473e-02],
                   [9.99973865e-01, 4.85948581e-04, -7.20693369e-03, -2.92196865e-01]])
    with open(os.path.join(seq, "calib.txt"), "w") as f:
        for key in ("P0", "P1", "P2", "P3"):
            f.write(key + ": " + " ".join("%.12e" % v for v in rng.normal(size=12)) + "\n")
        f.write("Tr: " + " ".join("%.12e" % v for v in tr.reshape(-1)) + "\n")
    scans, labels, rows = [], [], []
    for i in range(n_scans):
        n = 200 + 10 * i
        pts = rng.normal(size=(n, 4)).astype(np.float32)
        pts.tofile(os.path.join(seq, "velodyne", "%06d.bin" % i))
        lab = (rng.integers(0, 20, n).astype(np.uint32) | (rng.integers(0, 500, n).astype(np.uint32) << 16))
        lab.tofile(os.path.join(seq, "labels", "%06d.label" % i))
        write_labels(os.path.join(seq, "predictions", "%06d.label" % i), lab & 0xFFFF)
        c, s = np.cos(0.01 * i), np.sin(0.01 * i)
        t = np.array([[c, 0, s, 0.1 * i], [0, 1, 0, 0.01 * i], [-s, 0, c, 1.2 * i]])
        rows.append(t.reshape(-1))
        scans.append(pts)
        labels.append(lab)
    with open(os.path.join(seq, "poses.txt"), "w") as f:
        for r in rows:
            f.write(" ".join("%.9e" % v for v in r) + "\n")
    return seq, scans, labels


def test_roundtrip_and_history(tmp_path):
    seq_dir, scans, labels = _make_sequence(str(tmp_path))
    seq = KittiSequence(seq_dir)
    assert seq.scan_ids == list(range(5)) and len(seq.poses) == 5 and seq.poses[0].dtype == np.float32
    assert np.array_equal(read_scan(seq.scan_path(3)), scans[3])
    assert np.array_equal(read_labels(seq.label_path(3)), (labels[3] & 0xFFFF).astype(np.int64))
    lm = {i: (i + 1) % 20 for i in range(20)}
    assert np.array_equal(read_labels(seq.label_path(2), lm), ((labels[2] & 0xFFFF).astype(np.int64) + 1) % 20)
    frames, poses, keep = load_sample(seq, 3, 2)
    assert keep is None and len(frames) == 3
    assert np.array_equal(frames[1], scans[2]) and np.array_equal(frames[2], scans[1])     # frames[j] = scan n - j
    assert np.array_equal(poses[2], seq.poses[1])
    frames, poses, _ = load_sample(seq, 1, 3)                                                # history clipped at scan 0
    assert len(frames) == 2
    # FSA keep mask: class c kept in history scan at distance j iff j % step_c == 0
    steps = [0] + [1 if c % 2 else 2 for c in range(1, 20)]
    inv = {c: c for c in range(20)}
    frames, poses, keep = load_sample(seq, 4, 2, flexible_steps=steps, pseudo_folder="predictions", learning_map_inv=inv)
    want = [np.ones(len(scans[4]), np.uint8)]
    for j in (2, 1):
        lab = (labels[4 - j] & 0xFFFF).astype(np.int64)
        m = np.zeros(len(lab), bool)
        for c, st in enumerate(steps):
            if st and j % st == 0:
                m |= lab == c
        want.append(m.astype(np.uint8))
    assert np.array_equal(keep, np.concatenate(want))
    # the sample drops straight into the device front end's batch description
    from taseg_b200.frontend import MultiFrameBatch
    mfb = MultiFrameBatch([frames], [poses])
    assert mfb.total == sum(len(f) for f in frames) == len(keep) and mfb.n_cur == [len(scans[4])]
    assert np.array_equal(mfb.points[:len(scans[4])], scans[4]) and np.array_equal(mfb.points[len(scans[4]):][:len(scans[2])], scans[2])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_calib_and_poses_match_reference_parser(tmp_path):
    seq_dir, _, _ = _make_sequence(str(tmp_path), seed=3)
    sys.modules.setdefault("petrel_client", types.ModuleType("petrel_client"))
    sys.modules.setdefault("petrel_client.client", types.ModuleType("petrel_client.client"))
    setattr(sys.modules["petrel_client.client"], "Client", object)
    src = open(os.path.join(REF, "pcseg/data/dataset/semantickitti/semantickitti_ms.py")).read()
    # take the two parser methods out of the reference class without importing the dataset package (heavy deps)
    a, b = src.index("    def parse_calibration(self, filename):"), src.index("    def fuse_multi_scan(")
    ns = {"np": np}
    exec("class _P:\n" + src[a:b], ns)
    ref = ns["_P"]()
    calib_ref = ref.parse_calibration(os.path.join(seq_dir, "calib.txt"))
    calib = parse_calibration(os.path.join(seq_dir, "calib.txt"))
    assert calib.keys() == calib_ref.keys() and all(np.array_equal(calib[k], calib_ref[k]) for k in calib)
    poses_ref = [p.astype(np.float32) for p in ref.parse_poses(os.path.join(seq_dir, "poses.txt"), calib_ref)]
    poses = parse_poses(os.path.join(seq_dir, "poses.txt"), calib)
    assert len(poses) == len(poses_ref) and all(np.array_equal(a_, b_) for a_, b_ in zip(poses, poses_ref))


def test_tta_vote_and_label_dump(tmp_path):
    """Vote accumulation of R/train.py:471-503 (sum of the votes' logits, arg-max, uint32 .label file)."""
    import torch
    from taseg_b200.engine import tta_vote
    rng = np.random.default_rng(0)
    votes = [torch.from_numpy(rng.normal(size=(300, 20)).astype(np.float32)) for _ in range(4)]
    want = votes[0].numpy().copy()
    for v in votes[1:]:
        want += v.numpy()
    labels = tta_vote(votes)
    assert np.array_equal(labels.numpy(), want.argmax(1)) and np.array_equal(tta_vote(torch.stack(votes)).numpy(), want.argmax(1))
    assert np.allclose(tta_vote(votes, save_score=True).numpy(), want, atol=1e-6)
    path = str(tmp_path / "000000.label")
    write_labels(path, labels.numpy())
    assert np.array_equal(read_labels(path), want.argmax(1))
