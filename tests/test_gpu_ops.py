"""GPU parity tests, op level: every C-ABI entry point against the oracle / golden fixtures.
Bit-exact for hashes, tables, kernel maps, voxel sets, indices; fp32 features within 1e-3 relative
(max|a-b|/max|b|); bf16 tensor-core path within 2e-2."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import data_oracle as D
from oracle import ts_oracle as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts():
    import taseg_b200
    taseg_b200.install_as_torchsparse()
    return taseg_b200


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def npy(t):
    return t.detach().float().cpu().numpy() if t.is_floating_point() else t.detach().cpu().numpy()


def random_cloud(seed, n=20000, span=60, batches=3, neg=True):
    rng = np.random.default_rng(seed)
    c = rng.integers(-span if neg else 0, span, (n, 3)).astype(np.int32)
    b = rng.integers(0, batches, (n, 1)).astype(np.int32)
    return np.concatenate([c, b], 1), rng


def test_hash_and_query(ts, golden):
    from taseg_b200.nn import functional as F
    g = golden("ops_kat")
    assert np.array_equal(npy(F.sphash(cu(g["hash_kat_in"]))), g["hash_kat"])
    assert np.array_equal(npy(F.sphash(cu(g["coords"]))), g["hash"])
    kh = F.sphash(cu(g["coords"]), cu(g["offsets_k3_s1"]))
    assert np.array_equal(npy(kh), g["khash27"])
    assert np.array_equal(npy(F.sphashquery(kh, cu(g["hash"]))), g["query27"])
    # batch > 0 rows keep their own batch index (hash_cuda.cu:42-46)
    c, _ = random_cloud(1)
    off = T.get_kernel_offsets(2, 4)
    assert np.array_equal(npy(F.sphash(cu(c), cu(off))), T.sphash(c, off))
    # duplicates in references: first position wins; empty query
    ref = np.array([5, 7, 5, 9, 7], np.int64)
    q = np.array([[7, 5], [1, 9]], np.int64)
    assert np.array_equal(npy(F.sphashquery(cu(q), cu(ref))), T.sphashquery(q, ref))
    assert F.sphashquery(cu(np.zeros((0,), np.int64)), cu(ref)).numel() == 0


def test_sort_pairs(ts):
    from taseg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    for n, bits in [(1, 64), (257, 64), (5000, 40), (300001, 64)]:
        keys = torch.randint(0, 2 ** 62, (n,), generator=g, device="cuda", dtype=torch.int64)
        if bits < 64:
            keys = keys & ((1 << bits) - 1)
        keys[::7] = keys[0].clone()                                   # duplicates: stability matters
        ko, vo = ops.sort_pairs(keys, None, 0, bits)
        ref_k, ref_i = torch.sort(keys, stable=True)
        assert torch.equal(ko, ref_k) and torch.equal(vo.long(), ref_i), (n, bits)


def test_unique_and_quantize(ts, golden):
    from taseg_b200 import ops
    from taseg_b200.nn import functional as F
    from taseg_b200.utils.quantize import sparse_quantize
    g = golden("ops_kat")
    c, i, v = sparse_quantize(g["q_in"], 1, return_index=True, return_inverse=True)   # numpy in -> numpy out
    assert np.array_equal(c, g["q_coords"]) and np.array_equal(i, g["q_inds"]) and np.array_equal(v, g["q_inv"])
    c2 = sparse_quantize(cu(g["q_in"]))                                               # tensor in -> tensor out
    assert c2.is_cuda and np.array_equal(npy(c2), g["q_coords"])
    fc = np.random.default_rng(0).uniform(-3, 3, (5000, 3))
    a = sparse_quantize(fc, 0.25, return_index=True, return_inverse=True)
    b = T.sparse_quantize(fc, 0.25, return_index=True, return_inverse=True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # spdownsample: batches, negative coordinates (trunc toward zero), all strides on the path
    c, _ = random_cloud(2)
    cur = T.sparse_collate([np.unique(c[c[:, 3] == b][:, :3], axis=0) for b in range(3)], [np.zeros((1, 1))] * 3)[0]
    for s in (1, 2, 4, 8):
        want = T.spdownsample(cur, 2, 2, s)
        got = F.spdownsample(cu(cur), 2, 2, s)
        assert np.array_equal(npy(got), want), s
        cur = want
    # hash order (initial_voxelize)
    uc, first, inv = ops.unique_coords(cu(c), want_index=True, want_inverse=True, by_hash=True)
    h = T.sphash(c)
    uh, fi, iv = np.unique(h, return_index=True, return_inverse=True)
    assert np.array_equal(npy(uc), c[fi]) and np.array_equal(npy(first), fi) and np.array_equal(npy(inv), iv)
    with pytest.raises(RuntimeError):
        ops.unique_coords(cu(np.array([[1 << 20, 0, 0, 0]], np.int32)))
    # dense sort keys (caller-promised coordinate widths): same result, fewer radix passes; a broken promise raises
    pos = np.abs(c).astype(np.int32)
    bits = [int(pos[:, j].max()).bit_length() for j in range(4)]
    a = ops.unique_coords(cu(pos), trunc_stride=4, want_index=True, want_inverse=True, field_bits=bits)
    b = ops.unique_coords(cu(pos), trunc_stride=4, want_index=True, want_inverse=True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    with pytest.raises(RuntimeError):
        ops.unique_coords(cu(pos), field_bits=[bits[0] - 1, bits[1], bits[2], bits[3]])
    assert ops.unique_coords(cu(np.zeros((0, 4), np.int32))).shape == (0, 4)


def test_kernel_maps_and_convs(ts, golden):
    from taseg_b200 import SparseTensor
    from taseg_b200.nn import functional as F
    g = golden("ops_kat")
    x = SparseTensor(cu(g["conv_in"]), cu(g["coords"]), 1)
    y = F.conv3d(x, cu(g["conv_w3"]), 3)
    nb, ns, sz = x.kmaps[((1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    assert nb.dtype == torch.int64 and np.array_equal(npy(nb), g["kmap3_nbmaps"]) and np.array_equal(npy(ns), g["kmap3_nbsizes"])
    assert sz == (len(g["coords"]),) * 2
    assert rel_err(npy(y.F), g["conv_out3"]) < 1e-3
    y2 = F.conv3d(y, cu(g["conv_w2"]), 2, stride=2)
    assert np.array_equal(npy(y2.C), g["coords_s2"]) and y2.s == (2, 2, 2)
    nb, ns, _ = x.kmaps[((1, 1, 1), (2, 2, 2), (2, 2, 2), (1, 1, 1))]
    assert np.array_equal(npy(nb), g["kmap2_nbmaps"]) and np.array_equal(npy(ns), g["kmap2_nbsizes"])
    assert rel_err(npy(y2.F), g["conv_out2"]) < 1e-3
    y3 = F.conv3d(y2, cu(g["conv_w3b"]), 3)
    nb, ns, _ = x.kmaps[((2, 2, 2), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    assert np.array_equal(npy(nb), g["kmap3s2_nbmaps"]) and np.array_equal(npy(ns), g["kmap3s2_nbsizes"])
    assert rel_err(npy(y3.F), g["conv_out3b"]) < 1e-3
    y4 = F.conv3d(y3, cu(g["conv_w1"]), 1)
    assert rel_err(npy(y4.F), g["conv_out1"]) < 1e-3
    y5 = F.conv3d(y4, cu(g["conv_wt"]), 2, stride=2, transposed=True)
    assert y5.C is x.C and y5.s == (1, 1, 1) and rel_err(npy(y5.F), g["conv_outT"]) < 1e-3
    with pytest.raises(KeyError):
        F.conv3d(SparseTensor(cu(g["conv_in"]), cu(g["coords"]), 2), cu(g["conv_wt"][:, :5]), 2, stride=2, transposed=True)
    with pytest.raises(ValueError):
        F.conv3d(SparseTensor(cu(g["conv_in"][:, :3]), cu(g["coords"]), 1), cu(g["conv_w3"]), 3)
    b = F.conv3d(x, cu(g["conv_w3"]), 3, bias=torch.ones(8, device="cuda"))
    assert rel_err(npy(b.F), g["conv_out3"] + 1) < 1e-3


@pytest.mark.parametrize("stride,ks,tstride", [((2, 2, 1), (2, 2, 1), 1), (2, 3, 1), ((2, 2, 1), 3, 1), ((2, 2, 1), 3, (2, 2, 1)),
                                               ((1, 2, 2), (3, 1, 3), 1), (1, (3, 1, 3), 1)])
def test_anisotropic_and_expanding_convs(ts, stride, ks, tstride):
    """Cylinder3D-style geometries: anisotropic kernels / strides and the offset-expansion branch of
    spdownsample — coarse coordinates and kernel maps bit-exact, features within fp32 tolerance of the oracle."""
    from taseg_b200 import SparseTensor
    from taseg_b200.nn import functional as F
    from taseg_b200.utils import make_ntuple
    rng = np.random.default_rng(11)
    tsn = make_ntuple(tstride, 3)
    c = np.unique(rng.integers(0, 28, (4000, 3)).astype(np.int32), axis=0) * np.asarray(tsn, np.int32)
    c = np.concatenate([c, rng.integers(0, 2, (len(c), 1)).astype(np.int32)], 1)
    c = c[np.lexsort((c[:, 2], c[:, 1], c[:, 0], c[:, 3]))]
    feats = rng.normal(size=(len(c), 8)).astype(np.float32)
    kvol = int(np.prod(make_ntuple(ks, 3)))
    w = (rng.normal(size=(kvol, 8, 16)) * 0.2).astype(np.float32)
    x = SparseTensor(cu(feats), cu(c), tsn)
    y = F.conv3d(x, cu(w), ks, stride=stride)
    strided = any(v != 1 for v in make_ntuple(stride, 3))
    out_c = T.spdownsample(c, stride, ks, tstride) if strided else c
    assert np.array_equal(npy(y.C), out_c)
    assert y.s == tuple(a * b for a, b in zip(tsn, make_ntuple(stride, 3)))
    nbmaps, nbsizes = T.build_kmap(c, out_c, ks, tsn)
    nb, ns, sz = x.kmaps[(tsn, make_ntuple(ks, 3), make_ntuple(stride, 3), (1, 1, 1))]
    assert np.array_equal(npy(nb), nbmaps) and np.array_equal(npy(ns), nbsizes) and sz == (len(c), len(out_c))
    want = T.conv_forward(feats, w, nbmaps, nbsizes, (len(c), len(out_c)))
    assert rel_err(npy(y.F), want) < 1e-3


def test_conv_backward(ts, golden):
    from taseg_b200 import SparseTensor
    from taseg_b200.nn import functional as F
    g = golden("ops_kat")
    xi = cu(g["conv_in"]).requires_grad_(True)
    w = cu(g["conv_w3"]).requires_grad_(True)
    x = SparseTensor(xi, cu(g["coords"]), 1)
    y = F.conv3d(x, w, 3)
    y.F.backward(cu(g["bwd_gy"]))
    assert rel_err(npy(xi.grad), g["bwd_gx"]) < 1e-3 and rel_err(npy(w.grad), g["bwd_gw"]) < 1e-3
    yi = cu(g["conv_out3"]).requires_grad_(True)
    w2 = cu(g["conv_w2"]).requires_grad_(True)
    yy = x.derive(yi)
    F.conv3d(yy, w2, 2, stride=2).F.backward(cu(g["bwd2_gy"]))
    assert rel_err(npy(yi.grad), g["bwd2_gx"]) < 1e-3 and rel_err(npy(w2.grad), g["bwd2_gw"]) < 1e-3
    # transposed conv backward against the oracle
    rng = np.random.default_rng(3)
    nb, ns = g["kmap2_nbmaps"], g["kmap2_nbsizes"]
    n_c = len(g["coords_s2"])
    xin = rng.normal(size=(n_c, 16)).astype(np.float32)
    wt = (rng.normal(size=(8, 16, 8)) * 0.2).astype(np.float32)
    gy = rng.normal(size=(len(g["coords"]), 8)).astype(np.float32)
    gx_ref, gw_ref = T.conv_backward(xin, gy, wt, nb, ns, transposed=True)
    xt = cu(xin).requires_grad_(True)
    wtt = cu(wt).requires_grad_(True)
    coarse = SparseTensor(xt, cu(g["coords_s2"]), 2)
    coarse.cmaps, coarse.kmaps = x.cmaps, x.kmaps
    out = F.conv3d(coarse, wtt, 2, stride=2, transposed=True)
    assert rel_err(npy(out.F), T.conv_forward(xin, wt, nb, ns, (len(g["coords"]), n_c), True)) < 1e-3
    out.F.backward(cu(gy))
    assert rel_err(npy(xt.grad), gx_ref) < 1e-3 and rel_err(npy(wtt.grad), gw_ref) < 1e-3


def test_point_voxel_ops(ts, golden):
    from taseg_b200 import ops
    from taseg_b200.nn import functional as F
    g = golden("ops_kat")
    p = cu(g["pv_points"])
    for s, cs in [(1, g["coords"]), (2, g["coords_s2"])]:
        fl = torch.cat([torch.floor(p[:, :3] / s).int() * s, p[:, -1].int().view(-1, 1)], 1)
        off = cu(T.get_kernel_offsets(2, s))
        idx = F.sphashquery(F.sphash(fl, off), F.sphash(cu(cs)))
        w = F.calc_ti_weights(p, idx, scale=s).transpose(0, 1).contiguous()
        idx = idx.transpose(0, 1).contiguous()
        assert np.array_equal(npy(idx), g[f"dv{s}_idx"]) and np.abs(npy(w) - g[f"dv{s}_w"]).max() < 1e-6
        out = F.spdevoxelize(cu(g[f"dv{s}_feat"]), idx, w)
        assert rel_err(npy(out), g[f"dv{s}_out"]) < 1e-5
        # fused query (one kernel instead of hash+hash+table+~30 elementwise launches)
        tab = ops.Table.from_coords(cu(cs))
        i8, w8 = ops.trilinear_query(tab, p, s)
        assert np.array_equal(npy(i8), g[f"dv{s}_idx"]) and np.abs(npy(w8) - g[f"dv{s}_w"]).max() < 1e-6
        i8n, w8n = ops.trilinear_query(tab, p, s, nearest=True)
        assert (npy(i8n)[:, 1:] == -1).all() and (npy(w8n)[:, 1:] == 0).all() and np.array_equal(npy(i8n)[:, 0], g[f"dv{s}_idx"][:, 0])
        iq = F.sphashquery(F.sphash(fl), F.sphash(cu(cs)))
        assert np.array_equal(npy(ops.point_query(tab, p, s)), npy(iq))
        assert np.array_equal(npy(F.spcount(iq.int(), len(cs))), g[f"vx{s}_cnt"])
        vo = F.spvoxelize(cu(g[f"vx{s}_feat"]), cu(g[f"vx{s}_idx"]), cu(g[f"vx{s}_cnt"]))
        assert rel_err(npy(vo), g[f"vx{s}_out"]) < 1e-5
        # backward of both against the oracle
        gyv = np.random.default_rng(s).normal(size=g[f"vx{s}_out"].shape).astype(np.float32)
        f_in = cu(g[f"vx{s}_feat"]).requires_grad_(True)
        F.spvoxelize(f_in, cu(g[f"vx{s}_idx"]), cu(g[f"vx{s}_cnt"])).backward(cu(gyv))
        assert rel_err(npy(f_in.grad), T.spvoxelize_backward(gyv, g[f"vx{s}_idx"], g[f"vx{s}_cnt"], len(g[f"vx{s}_idx"]))) < 1e-5
        gyd = np.random.default_rng(s + 9).normal(size=g[f"dv{s}_out"].shape).astype(np.float32)
        vf = cu(g[f"dv{s}_feat"]).requires_grad_(True)
        F.spdevoxelize(vf, idx, w).backward(cu(gyd))
        assert rel_err(npy(vf.grad), T.spdevoxelize_backward(gyd, g[f"dv{s}_idx"], g[f"dv{s}_w"], len(cs))) < 1e-5
    # 16-bit storage paths
    hf = F.spdevoxelize(cu(g["dv1_feat"]).bfloat16(), cu(g["dv1_idx"]), cu(g["dv1_w"]))
    assert hf.dtype == torch.bfloat16 and rel_err(npy(hf), g["dv1_out"]) < 2e-2


@pytest.mark.parametrize("c,dtype", [(32, torch.float32), (96, torch.float32), (4, torch.float32), (64, torch.bfloat16),
                                     (48, torch.float16), (5, torch.float32)])
def test_point_voxel_vector_kernels(ts, c, dtype):
    """The lane-group / segmented-reduction kernels of pointvoxel.cu against the numpy oracle (TS voxelize / devoxelize
    semantics) on random maps with empty voxels, unmatched points (-1) and missing corners; the scatter-mean must be
    bit-identical run to run (no atomics).  c = 5 exercises the scalar fall-back (rows not a multiple of 16 bytes)."""
    from taseg_b200 import ops
    rng = np.random.default_rng(c)
    n, m = 50000, 9000
    idx = rng.integers(-1, m - 50, n).astype(np.int32)          # -1 = unmatched; the last 50 voxels stay empty
    counts = T.spcount(idx, m)
    feats = rng.normal(size=(n, c)).astype(np.float32)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    f_t = cu(feats).to(dtype)
    f_ref = f_t.float().cpu().numpy()
    got = ops.voxelize_forward(f_t, cu(idx), cu(counts))
    again = ops.voxelize_forward(f_t, cu(idx), cu(counts))
    assert got.dtype == dtype
    if (c * f_t.element_size()) % 16 == 0:      # segmented reduction (the scalar fall-back keeps the reference's atomics)
        assert torch.equal(got, again), "scatter-mean is not deterministic"
    ok = idx >= 0                                               # the numpy oracle has no "unmatched point" branch
    assert rel_err(npy(got.float()), T.spvoxelize(f_ref[ok], idx[ok], np.maximum(counts, 1))) < tol
    gy = rng.normal(size=(m, c)).astype(np.float32)
    g_t = cu(gy).to(dtype)
    gb = ops.voxelize_backward(g_t, cu(idx), cu(counts), n)
    want_gb = T.spvoxelize_backward(g_t.float().cpu().numpy(), np.where(ok, idx, 0), np.maximum(counts, 1), n)
    want_gb[~ok] = 0
    assert rel_err(npy(gb.float()), want_gb) < tol
    idx8 = rng.integers(-1, m, (n, 8)).astype(np.int32)
    w8 = rng.random((n, 8)).astype(np.float32)
    w8[idx8 < 0] = 0
    vf = cu(gy).to(dtype)
    dv = ops.devoxelize_forward(vf, cu(idx8), cu(w8))
    assert rel_err(npy(dv.float()), T.spdevoxelize(vf.float().cpu().numpy(), idx8, w8)) < tol
    top = cu(feats).to(dtype)
    db = ops.devoxelize_backward(top, cu(idx8), cu(w8), m)
    assert rel_err(npy(db.float()), T.spdevoxelize_backward(top.float().cpu().numpy(), idx8, w8, m)) < (1e-4 if dtype == torch.float32 else 2e-2)


def test_fuse_and_aggregate(ts, golden):
    from taseg_b200 import ops
    g = golden("fuse_kat")
    for sfx, psfx in [("", ""), ("2", "_2")]:
        out = ops.fuse_multi_scan(cu(g["points" + sfx]), g["pose0" + psfx], g["pose" + psfx])
        assert np.array_equal(npy(out).view(np.uint32), g["fused" + sfx].view(np.uint32))
    # nuScenes-style float64 transform
    rng = np.random.default_rng(0)
    R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    Tt = rng.normal(size=3) * 10
    pts = rng.normal(size=(5000, 5)).astype(np.float32) * 30
    got = npy(ops.transform_point(cu(pts), R, Tt))
    want = D.transform_point(pts, R, Tt)
    assert (got.view(np.uint32) != want.view(np.uint32)).mean() < 1e-4 and np.abs(got - want).max() < 1e-5
    # whole front end, batch of 2 samples x 3 frames, against the data oracle
    from taseg_b200 import synth
    spec = synth.SensorSpec(16, -24.8, 2.0, 300, 1.73, 60.0)
    frames_all, descr, want_feats, want_coords, want_n0 = [], [], [], [], []
    off = 0
    for b in range(2):
        frames, poses = synth.kitti_sample(100 + b, 3, spec=spec, n_boxes=30)
        ms, n0 = D.aggregate_kitti(frames, poses)
        q = D.quantize_ms(ms[:n0], ms, 0.05)
        want_feats.append(q["point_ms"])
        want_coords.append(np.concatenate([q["pc_ms_"], np.full((len(q["pc_ms_"]), 1), b, np.int32)], 1))
        order = [0] + list(range(len(frames) - 1, 0, -1))       # current first, then history oldest first
        for j in order:
            descr.append(dict(offset=off, count=len(frames[j]), sample=b, is_cur=int(j == 0), pose0=poses[0], pose=poses[j]))
            frames_all.append(frames[j])
            off += len(frames[j])
    feats, coords, flags, extent = ops.aggregate_quantize(cu(np.concatenate(frames_all)), descr, 2, 0.05)
    f2, c2, _, m = ops.compact_rows(flags, feats, coords)
    wf, wc = np.concatenate(want_feats), np.concatenate(want_coords)
    assert m == len(wf)
    assert np.array_equal(npy(f2).view(np.uint32), wf.view(np.uint32)) and np.array_equal(npy(c2), wc)


def test_aggregate_with_fsa_keep_mask(ts):
    """Flexible Step Aggregation on the device (semantickitti_ms.py:303-308): a per-point keep mask over the history
    scans (class c of scan -j kept iff j % step_c == 0) drives `tsg_aggregate_quantize`'s device `keep` input; the whole
    front end (warp, FSA drop, clamp, round, shift, dedup) against the data oracle run with the same pseudo labels."""
    from taseg_b200 import frontend, synth
    spec = synth.SensorSpec(16, -24.8, 2.0, 300, 1.73, 60.0)
    steps = [0, 1, 2, 3, 1, 2]                           # class 0 never kept, class 3 only every third scan
    class_ids = list(range(len(steps)))
    samples, pseudos, keeps = [], [], []
    for b in range(2):
        frames, poses = synth.kitti_sample(500 + b, 4, spec=spec, n_boxes=30)
        rng = np.random.default_rng(50 + b)
        pseudo = [rng.integers(0, len(steps), len(f)) for f in frames]
        samples.append((frames, poses))
        pseudos.append(pseudo)
        keep = [np.ones(len(frames[0]), np.uint8)]       # layout of MultiFrameBatch: current, then history oldest first
        for j in range(len(frames) - 1, 0, -1):
            keep.append(D.fsa_mask(pseudo[j], -j, steps, class_ids).astype(np.uint8))
        keeps.append(np.concatenate(keep))
    mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
    keep = np.concatenate(keeps)
    assert len(keep) == mfb.total and 0.2 < keep.mean() < 0.95
    out = frontend.aggregate_voxelize(cu(mfb.points), mfb, 0.05, cu(mfb.cur_idx), keep=cu(keep))
    want_c, want_f, want_pts = [], [], []
    for b, (frames, poses) in enumerate(samples):
        ms, n0 = D.aggregate_kitti(frames, poses, pseudo=pseudos[b], flexible_steps=steps, class_ids=class_ids)
        q = D.quantize_ms(ms[:n0], ms, 0.05)
        want_c.append(np.concatenate([q["pc_ms"], np.full((len(q["pc_ms"]), 1), b, np.int32)], 1))
        want_f.append(q["feat_ms"])
        want_pts.append(q["point_ms"])
    assert np.array_equal(npy(out["point_ms"]).view(np.uint32), np.concatenate(want_pts).view(np.uint32))
    assert np.array_equal(npy(out["coords"]), np.concatenate(want_c))
    assert np.array_equal(npy(out["feats"]).view(np.uint32), np.concatenate(want_f).view(np.uint32))


def _tc_case(seed, n, c0, c1, c_out, ks, relu, residual, out_dtype, dense_span):
    from taseg_b200 import ops
    rng = np.random.default_rng(seed)
    c = np.unique(rng.integers(0, dense_span, (n, 3)).astype(np.int32), axis=0)
    c = np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1)
    n = len(c)
    offs = T.get_kernel_offsets(ks, 1)
    k = len(offs)
    tab = ops.Table.from_coords(cu(c))
    km = ops.build_kmap(tab, n, cu(c), offs)
    x0 = torch.randn(n, c0, device="cuda").bfloat16()
    x1 = torch.randn(n, c1, device="cuda").bfloat16() if c1 else None
    w = (torch.randn(k, c0 + c1, c_out, device="cuda") * 0.1)
    bias = torch.randn(c_out, device="cuda")
    res = torch.randn(n, c_out, device="cuda").bfloat16() if residual else None
    packed = ops.pack_weights(w, c0, c1)
    got = ops.conv_forward_tc(x0, x1, packed, k, c_out, km.nbr, km.tile_mask(), n, bias=bias, residual=res, relu=relu,
                              out_dtype=out_dtype)
    # mask-sorted tile rows: a permutation of the same table, and bit-identical output rows (same MMAs per row)
    nbr_s, mask_s, perm = km.sorted()
    assert np.array_equal(np.sort(npy(perm)), np.arange(n)), "perm is not a permutation"
    assert nbr_s.shape[1] % 256 == 0 and bool((nbr_s[:, n:] == -1).all())   # padded stride, -1 in the padding rows
    assert torch.equal(nbr_s[:, :n], km.nbr[:, perm.long()])
    pos = list(range(k))    # bit position of offset j in the sort key: rare offsets are the most significant (sort.cu)
    if k == 27:
        order = [0, 2, 6, 8, 18, 20, 24, 26, 1, 3, 5, 7, 19, 21, 23, 25, 4, 22, 9, 11, 15, 17, 10, 12, 14, 16, 13]
        for r, j in enumerate(order):
            pos[j] = 26 - r
    keys = ((km.nbr >= 0).long() << torch.tensor(pos, device="cuda").view(-1, 1)).sum(0)
    ks_sorted = keys[perm.long()]
    assert bool((ks_sorted[1:] >= ks_sorted[:-1]).all()), "tile rows are not ordered by neighbour mask"
    same = ks_sorted[1:] == ks_sorted[:-1]
    assert bool((perm[1:][same] > perm[:-1][same]).all()), "sort is not stable"
    got_s = ops.conv_forward_tc(x0, x1, packed, k, c_out, nbr_s, mask_s, n, bias=bias, residual=res, relu=relu,
                                out_dtype=out_dtype, perm=perm)
    assert torch.equal(got, got_s), "mask-sorted convolution differs from the unsorted one"
    xin = torch.cat([x0, x1], 1).float() if c1 else x0.float()
    want = ops.conv_forward(xin, w.bfloat16().float(), km.nbr, n, bias=bias, residual=res.float() if residual else None, relu=relu)
    torch.cuda.synchronize()
    return rel_err(npy(got), npy(want)), n


@pytest.mark.parametrize("c0,c1,c_out,ks", [(16, 0, 32, 3), (32, 0, 32, 3), (64, 0, 64, 3), (96, 32, 96, 3), (128, 0, 128, 3),
                                            (256, 128, 256, 3), (32, 0, 32, 2), (128, 64, 128, 1), (256, 0, 256, 3),
                                            (16, 0, 256, 3)])   # last: 3 pipeline stages < 4 producer groups
def test_conv_tensor_core(ts, c0, c1, c_out, ks):
    err, n = _tc_case(c0 + c_out, 30000, c0, c1, c_out, ks, True, True, torch.bfloat16, 40)
    assert err < 2e-2, (err, n)
    err, n = _tc_case(c0 + c_out + 1, 700, c0, c1, c_out, ks, False, False, torch.float32, 12)   # partial last tile
    assert err < 1e-2, (err, n)


@pytest.mark.parametrize("c_mid,s0,s1,c_out,n", [(96, 96, 32, 96, 40000), (64, 32, 0, 64, 40000), (256, 256, 128, 256, 40000),
                                                 (256, 128, 0, 256, 9000), (128, 192, 0, 128, 9000), (32, 16, 0, 32, 900)])
def test_conv_tensor_core_folded_shortcut(ts, c_mid, s0, s1, c_out, n):
    """Second K phase (tsg_conv_fwd_tc2): out = relu(conv3(h) + [x0|x1] @ Ws + bias) in one launch, sorted and unsorted
    tile rows, against the fp32 kernel + a torch matmul on the same bf16-representable operands.  The 9000-row cases have
    fewer tiles than SMs (N-split work items)."""
    from taseg_b200 import ops
    rng = np.random.default_rng(c_mid + s0 + n)
    c = np.unique(rng.integers(0, 44, (n, 3)).astype(np.int32), axis=0)
    c = np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1)
    n = len(c)
    km = ops.build_kmap(ops.Table.from_coords(cu(c)), n, cu(c), T.get_kernel_offsets(3, 1))
    h = torch.randn(n, c_mid, device="cuda").bfloat16()
    x0 = torch.randn(n, s0, device="cuda").bfloat16()
    x1 = torch.randn(n, s1, device="cuda").bfloat16() if s1 else None
    w = torch.randn(27, c_mid, c_out, device="cuda") * 0.1
    ws = torch.randn(1, s0 + s1, c_out, device="cuda") * 0.1
    bias = torch.randn(c_out, device="cuda")
    packed, packed_s = ops.pack_weights(w, c_mid), ops.pack_weights(ws, s0, s1)
    xin = torch.cat([x0, x1], 1).float() if s1 else x0.float()
    want = torch.relu(ops.conv_forward(h.float(), w.bfloat16().float(), km.nbr, n) + xin @ ws[0].bfloat16().float() + bias)
    got = ops.conv_forward_tc(h, None, packed, 27, c_out, km.nbr, km.tile_mask(), n, bias=bias, relu=True,
                              shortcut=(x0, x1, packed_s, None))
    nbr_s, mask_s, perm = km.sorted()
    got_s = ops.conv_forward_tc(h, None, packed, 27, c_out, nbr_s, mask_s, n, bias=bias, relu=True, perm=perm,
                                shortcut=(x0, x1, packed_s, nbr_s[13]))
    torch.cuda.synchronize()
    assert torch.equal(got, got_s), "sorted tile rows change the result of the folded shortcut"
    assert rel_err(npy(got), npy(want)) < 2e-2


def test_conv_tensor_core_column_split(ts, monkeypatch):
    """TSG_TC_NSPLIT=2: (tile, column half) work items for launches with fewer tiles than SMs must give bit-identical
    results (every output element sees the same MMAs in the same order)."""
    from taseg_b200 import ops
    rng = np.random.default_rng(3)
    c = np.unique(rng.integers(0, 30, (9000, 3)).astype(np.int32), axis=0)
    c = np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1)
    n = len(c)
    km = ops.build_kmap(ops.Table.from_coords(cu(c)), n, cu(c), T.get_kernel_offsets(3, 1))
    nbr_s, mask_s, perm = km.sorted()
    for c_in, c_out in [(256, 256), (64, 128), (32, 96)]:
        x = torch.randn(n, c_in, device="cuda").bfloat16()
        res = torch.randn(n, c_out, device="cuda").bfloat16()
        bias = torch.randn(c_out, device="cuda")
        packed = ops.pack_weights(torch.randn(27, c_in, c_out, device="cuda") * 0.1, c_in)
        outs = []
        for ns in ("1", "2"):
            monkeypatch.setenv("TSG_TC_NSPLIT", ns)
            outs.append(ops.conv_forward_tc(x, None, packed, 27, c_out, nbr_s, mask_s, n, bias=bias, residual=res, relu=True, perm=perm))
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]), (c_in, c_out)


def test_conv_tensor_core_k_split(ts):
    """K-split work items (tsg_conv_split_items + tsg_conv_fwd_tc4): launches with about as many tiles as SMs sum heavy
    tiles with up to four work items on as many SMs.  The list must cover every (tile, offset) exactly once; the result must match the
    oracle, must not depend on which item finishes last (run-to-run torch.equal) and must stay within fp32 summation-order
    round-off of the unsplit launch; folded shortcut, residual, ReLU and fp32 output ride along."""
    from taseg_b200 import ops
    rng = np.random.default_rng(11)
    c = np.unique(rng.integers(0, 28, (14000, 3)).astype(np.int32), axis=0)      # dense block: most rows see all 27 offsets
    c = np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1)
    n = len(c)
    assert 64 * 128 < n <= ops.SPLIT_MAX_TILES * 128
    km = ops.build_kmap(ops.Table.from_coords(cu(c)), n, cu(c), T.get_kernel_offsets(3, 1))
    nbr_s, mask_s, perm = km.sorted()
    sp = ops.SplitItems(mask_s, n, 27, cap=7, max_parts=4)   # (KernelMap.split_items() returns one only when TSG_SPLIT_K=1: off by default)
    n_items = int(sp.n_items.item())
    items = npy(sp.items[:n_items])
    tiles = (n + 127) // 128
    cover = np.zeros(tiles, np.int64)
    for tile, m, part, slot in items:
        assert 0 <= tile < tiles and (cover[tile] & (int(m) & 0x7ffffff)) == 0
        cover[tile] |= int(m) & 0x7ffffff
        parts = part >> 8
        assert 1 <= parts <= sp.max_parts and (part & 0xff) < parts and 0 <= slot < sp.max_slots
        assert bin(int(m) & 0x7ffffff).count("1") <= 7 or (parts == 1 and slot == 0)
    assert np.array_equal(cover, npy(mask_s).astype(np.int64) & 0x7ffffff), "work items do not cover the tile masks"
    assert n_items > tiles, "no tile was split: the case does not exercise the partial-sum path"
    assert int((items[:, 2] >> 8).max()) >= 3, "no tile was split more than two ways"
    nbmaps, nbsizes = npy(km.nbmaps), npy(km.nbsizes)
    for c_in, c_out, od in [(256, 256, torch.bfloat16), (64, 128, torch.float32), (96, 96, torch.bfloat16)]:
        x = torch.randn(n, c_in, device="cuda").bfloat16()
        w = (torch.randn(27, c_in, c_out, device="cuda") * 0.05).bfloat16()
        packed = ops.pack_weights(w.float(), c_in)
        bias = torch.randn(c_out, device="cuda")
        res = torch.randn(n, c_out, device="cuda").bfloat16() if od == torch.bfloat16 else None
        ref = ops.conv_forward_tc(x, None, packed, 27, c_out, nbr_s, mask_s, n, bias=bias, residual=res, relu=True, perm=perm, out_dtype=od)
        outs = [ops.conv_forward_tc(x, None, packed, 27, c_out, nbr_s, mask_s, n, bias=bias, residual=res, relu=True, perm=perm,
                                    out_dtype=od, split=sp) for _ in range(4)]
        torch.cuda.synchronize()
        assert all(torch.equal(outs[0], o) for o in outs[1:]), "K-split result depends on the arrival order"
        assert rel_err(npy(outs[0].float()), npy(ref.float())) < (1e-5 if od == torch.float32 else 4e-3)
        want = T.conv_forward(npy(x.float()), npy(w.float()), nbmaps, nbsizes, (n, n), False) + npy(bias)
        if res is not None:
            want = want + npy(res.float())
        assert rel_err(npy(outs[0].float()), np.maximum(want, 0)) < (1e-3 if od == torch.float32 else 1e-2)
    assert int(sp.state.abs().sum().item()) == 0, "hand-off words were not re-armed"
    # folded shortcut: summed by part 0 only
    h = torch.randn(n, 128, device="cuda").bfloat16()
    x0 = torch.randn(n, 64, device="cuda").bfloat16()
    w = torch.randn(27, 128, 128, device="cuda") * 0.05
    ws = torch.randn(1, 64, 128, device="cuda") * 0.1
    packed, packed_s = ops.pack_weights(w, 128), ops.pack_weights(ws, 64)
    ref = ops.conv_forward_tc(h, None, packed, 27, 128, nbr_s, mask_s, n, relu=True, perm=perm, shortcut=(x0, None, packed_s, nbr_s[13]),
                              out_dtype=torch.float32)
    got = ops.conv_forward_tc(h, None, packed, 27, 128, nbr_s, mask_s, n, relu=True, perm=perm, shortcut=(x0, None, packed_s, nbr_s[13]),
                              out_dtype=torch.float32, split=sp)
    torch.cuda.synchronize()
    assert rel_err(npy(got), npy(ref)) < 1e-5


def test_conv_tensor_core_vs_oracle(ts):
    """The tensor-core forward and the bf16 weight gradient directly against the numpy oracle (TS conv semantics,
    oracle/ts_oracle.py conv_forward / conv_backward) on bf16-representable inputs — no detour through our own fp32 kernel."""
    from taseg_b200 import ops
    rng = np.random.default_rng(7)
    c = np.unique(rng.integers(0, 24, (6000, 3)).astype(np.int32), axis=0)
    c = np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1)
    n = len(c)
    offs = T.get_kernel_offsets(3, 1)
    km = ops.build_kmap(ops.Table.from_coords(cu(c)), n, cu(c), offs)
    nbmaps, nbsizes = npy(km.nbmaps), npy(km.nbsizes)
    x = torch.randn(n, 64, device="cuda").bfloat16()
    w = (torch.randn(27, 64, 96, device="cuda") * 0.1).bfloat16()
    gy = torch.randn(n, 96, device="cuda").bfloat16()
    nbr_s, mask_s, perm = km.sorted()
    got = ops.conv_forward_tc(x, None, ops.pack_weights(w.float(), 64), 27, 96, nbr_s, mask_s, n, perm=perm, out_dtype=torch.float32)
    want = T.conv_forward(npy(x.float()), npy(w.float()), nbmaps, nbsizes, (n, n), False)
    assert rel_err(npy(got), want) < 1e-3          # same bf16 operands, fp32 accumulation on both sides
    gx_ref, gw_ref = T.conv_backward(npy(x.float()), npy(gy.float()), npy(w.float()), nbmaps, nbsizes, False)
    gw = ops.conv_wgrad_bf16(x, gy, km.nbr, 27)
    assert rel_err(npy(gw), gw_ref) < 1e-3
    gw_tc = ops.conv_wgrad_tc(x, gy, km.nbr, 27)          # tcgen05 weight gradient against the oracle's backward
    assert rel_err(npy(gw_tc), gw_ref) < 1e-3
    # bf16 data gradient: tensor-core kernel over the transposed map with W[k]^T (output rounded to bf16)
    packed_t = ops.pack_weights(w.float().transpose(1, 2).contiguous(), 96)
    gx = ops.conv_forward_tc(gy, None, packed_t, 27, 64, km.nbr_t, km.tile_mask(True), n, out_dtype=torch.float32)
    assert rel_err(npy(gx), gx_ref) < 1e-3
    # ... and as the training path stores it (bf16 rows): per-layer bound of the bf16 data gradient, relative l2 <= 1e-2
    gx16 = ops.conv_forward_tc(gy, None, packed_t, 27, 64, km.nbr_t, km.tile_mask(True), n, out_dtype=torch.bfloat16)
    d = npy(gx16.float()).astype(np.float64) - gx_ref
    assert np.linalg.norm(d) / np.linalg.norm(gx_ref) < 1e-2
    # the autograd function end to end (padded 5-channel stem, 384-channel column blocks) against the oracle's backward
    from taseg_b200.nn.functional.conv import ConvolutionFunction
    for c_in, c_out in [(5, 32), (384, 256)]:
        xi = torch.randn(n, c_in, device="cuda").bfloat16().requires_grad_(True)
        wi = (torch.randn(27, c_in, c_out, device="cuda") * 0.05).bfloat16().requires_grad_(True)
        go = torch.randn(n, c_out, device="cuda").bfloat16()
        yo = ConvolutionFunction.apply(xi, wi, km, False, None)
        yo.backward(go)
        want_y = T.conv_forward(npy(xi.detach().float()), npy(wi.detach().float()), nbmaps, nbsizes, (n, n), False)
        want_gx, want_gw = T.conv_backward(npy(xi.detach().float()), npy(go.float()), npy(wi.detach().float()), nbmaps, nbsizes, False)
        for got_t, want_t in ((yo, want_y), (xi.grad, want_gx), (wi.grad, want_gw)):
            dd = npy(got_t.detach().float()).astype(np.float64) - want_t
            assert np.linalg.norm(dd) / np.linalg.norm(want_t) < 1e-2, (c_in, c_out)


def test_conv_fp32_split_precision(ts, monkeypatch):
    """fp32 tensors through the module path run on the tensor cores as three bf16 products (x = xh + xl, w = wh + wl;
    nn/functional/conv.py FP32_SPLIT): forward, data gradient and weight gradient against the oracle (fp32 arithmetic) at
    <= 1e-4 — a tenth of the fp32 parity bar — and the exact CUDA-core kernels stay available (TSG_FP32_SPLIT=0)."""
    from taseg_b200 import ops
    from taseg_b200.nn.functional import conv as C
    rng = np.random.default_rng(17)
    c = np.unique(rng.integers(0, 22, (5000, 3)).astype(np.int32), axis=0)
    c = np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1)
    n = len(c)
    km = ops.build_kmap(ops.Table.from_coords(cu(c)), n, cu(c), T.get_kernel_offsets(3, 1))
    nbmaps, nbsizes = npy(km.nbmaps), npy(km.nbsizes)
    assert C.FP32_SPLIT
    for c_in, c_out in [(5, 32), (96, 96), (320, 272)]:
        x = torch.randn(n, c_in, device="cuda").requires_grad_(True)
        w = (torch.randn(27, c_in, c_out, device="cuda") * 0.05).requires_grad_(True)
        gy = torch.randn(n, c_out, device="cuda")
        y = C.ConvolutionFunction.apply(x, w, km, False, None)
        y.backward(gy)
        want_y = T.conv_forward(npy(x.detach()), npy(w.detach()), nbmaps, nbsizes, (n, n), False)
        want_gx, want_gw = T.conv_backward(npy(x.detach()), npy(gy), npy(w.detach()), nbmaps, nbsizes, False)
        assert y.dtype == torch.float32
        assert rel_err(npy(y.detach()), want_y) < 1e-4, (c_in, c_out)
        assert rel_err(npy(x.grad), want_gx) < 1e-4 and rel_err(npy(w.grad), want_gw) < 1e-4, (c_in, c_out)
    monkeypatch.setattr(C, "FP32_SPLIT", False)
    x = torch.randn(n, 32, device="cuda")
    w = torch.randn(27, 32, 32, device="cuda") * 0.05
    y = C.ConvolutionFunction.apply(x, w, km, False, None)
    assert rel_err(npy(y), T.conv_forward(npy(x), npy(w), nbmaps, nbsizes, (n, n), False)) < 1e-5


@pytest.mark.parametrize("c0,span", [(16, 160), (32, 160), (32, 60), (96, 160)])
def test_conv_tensor_core_many_light_tiles(ts, c0, span):
    """Hundreds of super tiles per SM-wave with only one or two pipeline stages each (two sub-tiles per CTA, packed K
    slices, plans turning over faster than stages): the shape on which conv_tc v9 dead-locked, because a producer
    waited for the next plan while it still owed an arrival on a full barrier."""
    for rep in range(2):
        err, n = _tc_case(900 + c0 + rep, 100000, c0, 0, 32, 3, True, rep == 1, torch.bfloat16, span)
        assert n > 2 * 148 * 256 and err < 2e-2, (err, n)


@pytest.mark.parametrize("c_in,c_out,ks,n", [(96, 96, 3, 30000), (16, 32, 3, 30000), (256, 256, 3, 6000), (128, 96, 2, 20000),
                                             (40, 24, 3, 500)])
def test_conv_wgrad_bf16(ts, c_in, c_out, ks, n):
    """Tensor-core weight gradient (bf16 rows, in-kernel pair compaction) against the exact fp32 kernel on the same
    bf16-representable inputs: both accumulate the same fp32 products, only the summation order differs."""
    from taseg_b200 import ops
    rng = np.random.default_rng(c_in + c_out)
    c = np.unique(rng.integers(0, 36, (n, 3)).astype(np.int32), axis=0)
    c = np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1)
    offs = T.get_kernel_offsets(ks, 1)
    km = ops.build_kmap(ops.Table.from_coords(cu(c)), len(c), cu(c), offs)
    x = torch.randn(len(c), c_in, device="cuda").bfloat16()
    gy = torch.randn(len(c), c_out, device="cuda").bfloat16()
    got = ops.conv_wgrad_bf16(x, gy, km.nbr, len(offs))
    want = ops.conv_wgrad(x.float(), gy.float(), km.nbr, len(offs))
    torch.cuda.synchronize()
    assert got.shape == want.shape == (len(offs), c_in, c_out)
    assert rel_err(npy(got), npy(want)) < 1e-4
    got_tc = ops.conv_wgrad_tc(x, gy, km.nbr, len(offs))     # the tcgen05 kernel: same products, fp32 accumulation in TMEM
    torch.cuda.synchronize()
    assert got_tc.shape == want.shape
    assert rel_err(npy(got_tc), npy(want)) < 1e-4


def test_conv_wgrad_tc_wide_and_empty(ts):
    """tcgen05 weight gradient: more than 128 input channels (several TMEM-lane tiles), a transposed (stride-2) map, an
    offset without any pair, fewer pairs than one 64-pair chunk, and run-to-run determinism with one split."""
    from taseg_b200 import ops
    rng = np.random.default_rng(5)
    c = np.unique(rng.integers(0, 30, (12000, 3)).astype(np.int32), axis=0)
    c = np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1)
    n = len(c)
    tab = ops.Table.from_coords(cu(c))
    km = ops.build_kmap(tab, n, cu(c), T.get_kernel_offsets(3, 1))
    for c_in, c_out in [(384, 256), (256, 128), (200, 72)]:
        x = torch.randn(n, c_in, device="cuda").bfloat16()
        gy = torch.randn(n, c_out, device="cuda").bfloat16()
        want = ops.conv_wgrad(x.float(), gy.float(), km.nbr, 27)
        got = ops.conv_wgrad_tc(x, gy, km.nbr, 27)
        torch.cuda.synchronize()
        assert rel_err(npy(got), npy(want)) < 1e-4, (c_in, c_out)
    # a map with empty offsets and very few pairs: 40 isolated voxels (only the centre offset has pairs)
    iso = np.stack([np.arange(40) * 5, np.zeros(40), np.zeros(40), np.zeros(40)], 1).astype(np.int32)
    km2 = ops.build_kmap(ops.Table.from_coords(cu(iso)), 40, cu(iso), T.get_kernel_offsets(3, 1))
    x = torch.randn(40, 32, device="cuda").bfloat16()
    gy = torch.randn(40, 64, device="cuda").bfloat16()
    got = ops.conv_wgrad_tc(x, gy, km2.nbr, 27)
    want = ops.conv_wgrad(x.float(), gy.float(), km2.nbr, 27)
    torch.cuda.synchronize()
    assert rel_err(npy(got), npy(want)) < 1e-4
    assert float(got[:13].abs().max()) == 0.0 and float(got[14:].abs().max()) == 0.0
    again = ops.conv_wgrad_tc(x, gy, km2.nbr, 27)
    assert torch.equal(got, again)


def test_backend_mirror(ts, golden):
    from taseg_b200 import backend as B
    g = golden("ops_kat")
    x, w = cu(g["conv_in"]), cu(g["conv_w3"])
    out = torch.zeros(len(g["coords"]), 8, device="cuda")
    B.convolution_forward_cuda(x, out, w, cu(g["kmap3_nbmaps"]).int(), torch.from_numpy(g["kmap3_nbsizes"]).int(), False)
    assert rel_err(npy(out), g["conv_out3"]) < 1e-3
    gx, gw = torch.zeros_like(x), torch.zeros_like(w)
    B.convolution_backward_cuda(x, gx, cu(g["bwd_gy"]), w, gw, cu(g["kmap3_nbmaps"]).int(), torch.from_numpy(g["kmap3_nbsizes"]).int(), False)
    assert rel_err(npy(gx), g["bwd_gx"]) < 1e-3 and rel_err(npy(gw), g["bwd_gw"]) < 1e-3
    q = B.hash_query_cuda(cu(g["khash27"]).reshape(-1), cu(g["hash"]), torch.arange(len(g["hash"]), device="cuda"))
    assert np.array_equal(npy(q).reshape(27, -1) - 1, g["query27"])
    assert np.array_equal(npy(B.hash_cuda(cu(g["coords"]))), g["hash"])
