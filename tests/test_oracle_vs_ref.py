"""The numpy oracle against oracle/_ref — the reference's own CPU backend compiled by oracle/build_ref.py —
on fresh seeded inputs (sizes/shapes the fixtures do not cover).  Skipped where the .so is absent."""
import numpy as np
import pytest

from oracle import net_oracle as N
from oracle import ref_backend as RB
from oracle import ts_oracle as T
from conftest import rel_err

pytestmark = pytest.mark.skipif(not RB.available(), reason="oracle/_ref not built")


def _cloud(seed, n=3000, span=40):
    rng = np.random.default_rng(seed)
    c = np.unique(rng.integers(0, span, (n, 3)).astype(np.int32), axis=0)
    return np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1), rng


def test_native_entry_points():
    R = RB.RefOps
    c, rng = _cloud(0)
    assert np.array_equal(R.sphash(c), T.sphash(c))
    for ks, s in [(3, 1), (2, 2), (3, 4)]:
        off = T.get_kernel_offsets(ks, s)
        assert np.array_equal(R.sphash(c, off), T.sphash(c, off))
    q = T.sphash(c, T.get_kernel_offsets(3, 1))
    assert np.array_equal(R.sphashquery(q, T.sphash(c)), T.sphashquery(q, T.sphash(c)))
    idx = rng.integers(-1, 50, 5000).astype(np.int32)
    assert np.array_equal(R.spcount(idx, 50), T.spcount(idx, 50))
    idx = idx[idx >= 0]
    cnt = T.spcount(idx, 50)
    f = rng.normal(size=(len(idx), 7)).astype(np.float32)
    assert rel_err(T.spvoxelize(f, idx, cnt), R.spvoxelize(f, idx, cnt)) < 1e-6
    iq = rng.integers(-1, 50, (400, 8)).astype(np.int32)
    w = rng.uniform(size=(400, 8)).astype(np.float32)
    vf = rng.normal(size=(50, 9)).astype(np.float32)
    assert rel_err(T.spdevoxelize(vf, iq, w), R.spdevoxelize(vf, iq, w)) < 1e-6


@pytest.mark.parametrize("cin,cout", [(4, 16), (32, 32), (48, 24)])
def test_conv_forward_backward(cin, cout):
    R = RB.RefOps
    c, rng = _cloud(cin)
    x = rng.normal(size=(len(c), cin)).astype(np.float32)
    w = (rng.normal(size=(27, cin, cout)) * 0.1).astype(np.float32)
    nb, ns = T.build_kmap(c, c, 3, 1)
    sz = (len(c), len(c))
    assert rel_err(T.conv_forward(x, w, nb, ns, sz), R.conv_forward(x, w, nb, ns, sz)) < 1e-5
    gy = rng.normal(size=(len(c), cout)).astype(np.float32)
    a, b = T.conv_backward(x, gy, w, nb, ns), R.conv_backward(x, gy, w, nb, ns)
    assert rel_err(a[0], b[0]) < 1e-5 and rel_err(a[1], b[1]) < 1e-5
    oc = T.spdownsample(c, 2, 2, 1)
    nb2, ns2 = T.build_kmap(c, oc, 2, 1)
    w2 = (rng.normal(size=(8, cin, cout)) * 0.1).astype(np.float32)
    sz2 = (len(c), len(oc))
    y = T.conv_forward(x, w2, nb2, ns2, sz2)
    assert rel_err(y, R.conv_forward(x, w2, nb2, ns2, sz2)) < 1e-5
    wt = (rng.normal(size=(8, cout, cin)) * 0.1).astype(np.float32)
    assert rel_err(T.conv_forward(y, wt, nb2, ns2, sz2, True), R.conv_forward(y, wt, nb2, ns2, sz2, True)) < 1e-5


def test_network_on_reference_backend(golden):
    g = golden("net_spvcnn")
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    a = N.Net(sd, ops=RB.RefOps).spvcnn(g["coords"], g["feats"])
    assert rel_err(a[::4], g["voxel_logits"]) < 1e-4
