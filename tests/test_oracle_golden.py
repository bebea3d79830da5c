"""The numpy oracle against fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py).
CPU only; this is what pins the oracle (the reference itself has no tests)."""
import hashlib

import numpy as np
import pytest

from oracle import data_oracle as D
from oracle import net_oracle as N
from oracle import ts_oracle as T
from conftest import rel_err


def sha(a, dtype=None):
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a).astype(dtype))
    return hashlib.sha256(a.tobytes()).hexdigest()


def test_hash_known_answers(golden):
    g = golden("ops_kat")
    assert np.array_equal(T.sphash(g["hash_kat_in"]), g["hash_kat"])
    assert T.sphash(np.array([[0, 0, 0, 0]], np.int32))[0] == 947293587111810033      # SURVEY §8c [probe]
    assert T.sphash(np.array([[1, 0, 0, 0]], np.int32))[0] == 948793285165995886
    assert np.array_equal(T.sphash(g["coords"]), g["hash"])
    assert np.array_equal(T.sphash(g["coords"], g["offsets_k3_s1"]), g["khash27"])


def test_offsets_and_query(golden):
    g = golden("ops_kat")
    for ks, st in [(3, 1), (2, 1), (2, 2), (3, 2)]:
        assert np.array_equal(T.get_kernel_offsets(ks, st), g[f"offsets_k{ks}_s{st}"])
    assert np.array_equal(T.sphashquery(g["khash27"], g["hash"]), g["query27"])


def test_sparse_quantize(golden):
    g = golden("ops_kat")
    c, i, v = T.sparse_quantize(g["q_in"], 1, return_index=True, return_inverse=True)
    assert np.array_equal(c, g["q_coords"]) and np.array_equal(i, g["q_inds"]) and np.array_equal(v, g["q_inv"])


def test_kernel_maps_and_convs(golden):
    g = golden("ops_kat")
    x = T.SparseTensor(g["conv_in"], g["coords"], 1)
    y = T.conv3d(x, g["conv_w3"], 3)
    nb, ns, _ = x.kmaps[((1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    assert np.array_equal(nb, g["kmap3_nbmaps"]) and np.array_equal(ns, g["kmap3_nbsizes"])
    assert rel_err(y.F, g["conv_out3"]) < 1e-5
    y2 = T.conv3d(y, g["conv_w2"], 2, stride=2)
    assert np.array_equal(y2.C, g["coords_s2"])
    nb, ns, _ = x.kmaps[((1, 1, 1), (2, 2, 2), (2, 2, 2), (1, 1, 1))]
    assert np.array_equal(nb, g["kmap2_nbmaps"]) and np.array_equal(ns, g["kmap2_nbsizes"])
    assert rel_err(y2.F, g["conv_out2"]) < 1e-5
    y3 = T.conv3d(y2, g["conv_w3b"], 3)
    nb, ns, _ = x.kmaps[((2, 2, 2), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    assert np.array_equal(nb, g["kmap3s2_nbmaps"]) and np.array_equal(ns, g["kmap3s2_nbsizes"])
    assert rel_err(y3.F, g["conv_out3b"]) < 1e-5
    y4 = T.conv3d(y3, g["conv_w1"], 1)
    assert rel_err(y4.F, g["conv_out1"]) < 1e-5
    y5 = T.conv3d(y4, g["conv_wt"], 2, stride=2, transposed=True)
    assert y5.C is x.C and rel_err(y5.F, g["conv_outT"]) < 1e-5
    with pytest.raises(KeyError):
        T.conv3d(T.SparseTensor(g["conv_in"], g["coords"], 2), g["conv_wt"][:, :5], 2, stride=2, transposed=True)
    with pytest.raises(ValueError):
        T.conv_forward(g["conv_in"][:, :3], g["conv_w3"], g["kmap3_nbmaps"], g["kmap3_nbsizes"], (1, 1))


def test_conv_backward(golden):
    g = golden("ops_kat")
    n = len(g["coords"])
    gx, gw = T.conv_backward(g["conv_in"], g["bwd_gy"], g["conv_w3"], g["kmap3_nbmaps"], g["kmap3_nbsizes"])
    assert rel_err(gx, g["bwd_gx"]) < 1e-5 and rel_err(gw, g["bwd_gw"]) < 1e-5
    gx, gw = T.conv_backward(g["conv_out3"], g["bwd2_gy"], g["conv_w2"], g["kmap2_nbmaps"], g["kmap2_nbsizes"])
    assert gx.shape[0] == n and rel_err(gx, g["bwd2_gx"]) < 1e-5 and rel_err(gw, g["bwd2_gw"]) < 1e-5


def test_point_voxel_ops(golden):
    g = golden("ops_kat")
    p = g["pv_points"]
    for s, cs in [(1, g["coords"]), (2, g["coords_s2"])]:
        fl = T._floor_to_stride(p, s)
        idx = T.sphashquery(T.sphash(fl, T.get_kernel_offsets(2, s)), T.sphash(cs))
        w = T.calc_ti_weights(p, idx, s).T
        assert np.array_equal(idx.T, g[f"dv{s}_idx"])
        assert np.abs(w - g[f"dv{s}_w"]).max() < 1e-6
        assert rel_err(T.spdevoxelize(g[f"dv{s}_feat"], g[f"dv{s}_idx"], g[f"dv{s}_w"]), g[f"dv{s}_out"]) < 1e-6
        iq = T.sphashquery(T.sphash(fl), T.sphash(cs))
        assert np.array_equal(T.spcount(iq, len(cs)), g[f"vx{s}_cnt"])
        assert rel_err(T.spvoxelize(g[f"vx{s}_feat"], g[f"vx{s}_idx"], g[f"vx{s}_cnt"]), g[f"vx{s}_out"]) < 1e-6


def test_fuse_multi_scan_bit_exact(golden):
    g = golden("fuse_kat")
    for sfx in ["", "2"]:
        pts, p0, p1, ref = g["points" + sfx], g["pose0" + ("_2" if sfx else "")], g["pose" + ("_2" if sfx else "")], g["fused" + sfx]
        out = D.fuse_multi_scan(pts, p0, p1)
        assert out.dtype == np.float32 and np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def _check_net(g, logits_vox, feats):
    assert rel_err(logits_vox[::4], g["voxel_logits"]) < 1e-4
    for name, st in feats.items():
        step = int(g[f"Fstep_{name}"])
        assert int(g[f"Cn_{name}"]) == st.C.shape[0]
        assert sha(st.C, np.int32) == str(g[f"Csha_{name}"]), name
        assert rel_err(st.F[::step], g[f"F_{name}"]) < 1e-4, name


class _Tap(N.Net):
    def __init__(self, sd):
        super().__init__(sd)
        self.taps = {}

    def stem(self, x):
        self.taps["stem"] = r = super().stem(x)
        return r

    def stage(self, x, p):
        self.taps[p] = r = super().stage(x, p)
        return r

    def up(self, x, skip, p):
        self.taps[p] = r = super().up(x, skip, p)
        return r


def _sd(g):
    return {k[3:]: g[k] for k in g.files if k.startswith("sd/")}


def test_minkunet_ms_network(golden):
    g = golden("net_minkunet_ms")
    sizes = g["frame_sizes"]
    frames = np.split(g["frames"], np.cumsum(sizes)[:-1])
    ms, n0 = D.aggregate_kitti(frames, list(g["poses"]))
    assert sha(ms) == str(g["xyzret_ms_sha"]) and n0 == int(g["n_current"])
    q = D.quantize_ms(ms[:n0], ms, 0.05)
    coords, feats = T.sparse_collate([q["pc_ms"]], [q["feat_ms"]])
    assert np.array_equal(coords, g["coords"]) and np.array_equal(feats, g["feats"])
    assert np.array_equal(q["inds_ms"], g["inds"]) and np.array_equal(q["inverse_map_ms"], g["inverse"])
    net = _Tap(_sd(g))
    logits = net.minkunet_ms(coords, feats)
    _check_net(g, logits, net.taps)
    pm = np.zeros(len(q["pc_ms_"]), bool)
    pm[:n0] = True
    assert rel_err(N.gather_points(logits, q["inverse_map_ms"], pm, n0), g["point_logits"]) < 1e-4
    x = net.taps["up4"]
    for key, (nb, ns, sz) in x.kmaps.items():
        tag = "kmap_s%d_k%d_st%d" % (key[0][0], key[1][0], key[2][0])
        assert sha(nb, np.int64) == str(g[tag + "_nbmaps_sha"]), tag
        assert np.array_equal(ns, g[tag + "_nbsizes"]) and tuple(g[tag + "_sizes"]) == tuple(sz)


@pytest.mark.parametrize("kind", ["minkunet", "spvcnn"])
def test_single_frame_networks(golden, kind):
    g = golden("net_" + kind)
    q = D.quantize_single(g["points"], 0.05)
    coords, feats = T.sparse_collate([q["pc"]], [q["feat"]])
    assert np.array_equal(coords, g["coords"]) and np.array_equal(q["inverse_map"], g["inverse"])
    net = _Tap(_sd(g))
    logits = getattr(net, kind)(coords, feats)
    _check_net(g, logits, net.taps)
    assert rel_err(N.gather_points(logits, q["inverse_map"]), g["point_logits"]) < 1e-4
