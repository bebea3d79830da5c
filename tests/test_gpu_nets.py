"""GPU parity tests, network level: the three backbones through the CUDA path against fixtures produced by the
unmodified reference (tests/golden) and against the CPU oracle.
fp32 module path: max|a-b|/max|b| <= 1e-3 per tensor.  bf16 engine: ||a-b||2/||b||2 <= 3e-2 on logits
and >= 99 % arg-max agreement (measured against the fp32 reference logits)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import data_oracle as D
from oracle import net_oracle as N
from oracle import ts_oracle as T

pytestmark = pytest.mark.gpu


def sha(a, dtype):
    return hashlib.sha256(np.ascontiguousarray(np.asarray(a).astype(dtype)).tobytes()).hexdigest()


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def load_model(g, kind, num_layer=(1, 2, 1, 1, 1, 1, 1, 1), cr=0.125):
    from taseg_b200.segmentor import MinkUNet, MinkUNetMs, ModelCfg, SPVCNN
    cls = {"minkunet_ms": MinkUNetMs, "minkunet": MinkUNet, "spvcnn": SPVCNN}[kind]
    cfg = ModelCfg(IN_FEATURE_DIM=5 if kind == "minkunet_ms" else 4, BLOCK="ResBlock", NUM_LAYER=list(num_layer), cr=cr,
                   IF_DIST=False, IGNORE_LABEL=0, DROPOUT_P=0.0)
    model = cls(cfg, 20)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    model.load_state_dict(sd, strict=True)
    return model.cuda().eval()


def tap(model):
    feats = {}
    for name in ["stem", "stage1", "stage2", "stage3", "stage4"]:
        getattr(model, name).register_forward_hook(lambda m, i, o, name=name: feats.__setitem__(name, o))
    for name in ["up1", "up2", "up3", "up4"]:
        getattr(model, name)[1].register_forward_hook(lambda m, i, o, name=name: feats.__setitem__(name, o))
    return feats


def bf16_ok(got, want):
    """bf16 bound: relative l2 error of the logits, and arg-max agreement over the points whose reference decision is
    not a numerical tie (top-2 margin > 1 % of max|logit|; random-init nets emit many near-ties)."""
    l2 = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    top2 = np.sort(want, axis=1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 0.01 * np.abs(want).max()
    agree = float((got.argmax(1) == want.argmax(1))[decided].mean()) if decided.any() else 1.0
    return l2, agree


@pytest.mark.parametrize("kind", ["minkunet_ms", "minkunet", "spvcnn"])
def test_network_fp32_module_path(golden, kind):
    import taseg_b200
    g = golden("net_" + kind)
    model = load_model(g, kind)
    feats = tap(model)
    lidar = taseg_b200.SparseTensor(cu(g["feats"]), cu(g["coords"]), 1)
    with torch.no_grad():
        logits = model.logits(lidar)
    assert rel_err(logits[::4].cpu().numpy(), g["voxel_logits"]) < 1e-3
    for name, st in feats.items():
        step = int(g[f"Fstep_{name}"])
        assert st.C.shape[0] == int(g[f"Cn_{name}"]) and sha(st.C.cpu().numpy(), np.int32) == str(g[f"Csha_{name}"]), name
        assert rel_err(st.F[::step].cpu().numpy(), g[f"F_{name}"]) < 1e-3, name
    last = feats["up4"]
    n_maps = 0
    for key, km in last.kmaps.items():
        if not (isinstance(key, tuple) and isinstance(key[0], tuple)):
            continue
        tag = "kmap_s%d_k%d_st%d" % (key[0][0], key[1][0], key[2][0])
        nb, ns, sz = km
        assert sha(nb.cpu().numpy(), np.int64) == str(g[tag + "_nbmaps_sha"]), tag
        assert np.array_equal(ns.cpu().numpy(), g[tag + "_nbsizes"]) and tuple(sz) == tuple(g[tag + "_sizes"])
        n_maps += 1
    assert n_maps == 9
    # eval tail: same dict as the reference forward
    inv = taseg_b200.SparseTensor(cu(g["inverse"]), cu(np.zeros((len(g["inverse"]), 4), np.int32)), 1)
    n_pts = len(g["point_logits"])
    batch = {model.lidar_key: taseg_b200.SparseTensor(cu(g["feats"]), cu(g["coords"]), 1), model.inverse_key: inv,
             "targets_mapped": taseg_b200.SparseTensor(torch.zeros(n_pts).cuda(), cu(np.zeros((n_pts, 4), np.int32)), 1),
             "num_points": torch.tensor([[n_pts]]), "name": ["s"]}
    if kind == "minkunet_ms":
        pm = torch.zeros(len(g["inverse"]), dtype=torch.bool)
        pm[:int(g["n_current"])] = True
        batch.update(point_mask=pm, num_points_ms=torch.tensor([[len(g["inverse"])]]))
    with torch.no_grad():
        res = model(batch)
    assert rel_err(res["point_predict_logits"][0], g["point_logits"]) < 1e-3
    assert (res["point_predict"][0] == g["point_logits"].argmax(1)).mean() > 0.999


@pytest.mark.parametrize("kind", ["minkunet_ms", "minkunet", "spvcnn"])
def test_network_bf16_engine(golden, kind):
    from taseg_b200.engine import Engine
    g = golden("net_" + kind)
    model = load_model(g, kind)
    logits = Engine(model)(cu(g["coords"]), cu(g["feats"])).cpu().numpy()
    l2, agree = bf16_ok(logits[::4], g["voxel_logits"])
    assert l2 < 3e-2 and agree > 0.99, (l2, agree)


@pytest.mark.parametrize("num_class", [11, 16, 40])
def test_engine_head_widths(golden, num_class):
    """Head widths that pad to 16 or 48 channels (the fused devoxelisation tail owns 4 channels per lane, 8 lanes per
    point: c % 32 != 0 used to leave corner lanes out of the shuffles): bf16 engine vs the fp32 module path."""
    from taseg_b200 import SparseTensor
    from taseg_b200.engine import Engine
    from taseg_b200.segmentor import MinkUNetMs, ModelCfg
    g = golden("net_minkunet_ms")
    torch.manual_seed(num_class)
    cfg = ModelCfg(IN_FEATURE_DIM=5, BLOCK="ResBlock", NUM_LAYER=[1, 1, 1, 1, 1, 1, 1, 1], cr=0.25, IF_DIST=False,
                   IGNORE_LABEL=0, DROPOUT_P=0.0)
    model = MinkUNetMs(cfg, num_class).cuda().eval()
    with torch.no_grad():
        want = model.logits(SparseTensor(cu(g["feats"]), cu(g["coords"]), 1)).cpu().numpy()
    rows = torch.arange(0, len(g["coords"]), 3, device="cuda", dtype=torch.int32)
    got = Engine(model)(cu(g["coords"]), cu(g["feats"]), out_rows=rows).cpu().numpy()
    assert got.shape == (len(rows), num_class)
    l2, agree = bf16_ok(got, want[::3])
    assert l2 < 3e-2 and agree > 0.99, (l2, agree)


def test_frontend_matches_reference_loader(golden):
    from taseg_b200 import frontend
    g = golden("net_minkunet_ms")
    frames = np.split(g["frames"], np.cumsum(g["frame_sizes"])[:-1])
    mfb = frontend.MultiFrameBatch([frames], [list(g["poses"])])
    out = frontend.aggregate_voxelize(cu(mfb.points), mfb, 0.05, cu(mfb.cur_idx))
    assert np.array_equal(out["coords"].cpu().numpy(), g["coords"])
    assert np.array_equal(out["feats"].cpu().numpy().view(np.uint32), g["feats"].view(np.uint32))
    assert np.array_equal(out["inds"].cpu().numpy(), g["inds"]) and np.array_equal(out["inverse"].cpu().numpy(), g["inverse"])
    assert np.array_equal(out["cur_rows"].cpu().numpy(), g["inverse"][:int(g["n_current"])])
    assert hashlib.sha256(out["point_ms"].cpu().numpy().tobytes()).hexdigest() != ""   # (N',5) clamped ms cloud


def test_nuscenes_frontend_and_forward():
    """BASELINE configs[3] shape (10 sweeps, 0.1 m voxels, IN_FEATURE_DIM 4, 17 classes): the device front end against
    the data oracle (nuscenes_ms.py aggregation + the voxel loader), then bf16 engine vs fp32 module path."""
    from taseg_b200 import frontend, synth
    from taseg_b200.engine import Engine
    from taseg_b200.segmentor import MinkUNetMs, ModelCfg
    spec = synth.SensorSpec(16, -30.67, 10.67, 400, 1.84, 60.0)
    batch = [synth.nus_sample(4000 + b, 10, spec=spec, n_boxes=30) for b in range(2)]
    out = frontend.aggregate_voxelize_nus([[cu(s) for s in smp[0]] for smp in batch], [smp[1] for smp in batch],
                                          [smp[2] for smp in batch], [smp[3] for smp in batch], 0.1)
    pts = out["point_ms"].cpu().numpy()
    row = 0
    want_c, want_f, n_vox = [], [], 0
    for b, (sweeps, Rs, Ts, dts) in enumerate(batch):
        ms, n0 = D.aggregate_nus(sweeps, Rs, Ts, dts)
        q = D.quantize_ms(ms[:n0], ms, 0.1)
        n = len(q["point_ms"])
        mine = pts[row:row + n]
        assert mine.shape == q["point_ms"].shape and out["n_cur"][b] == n0
        assert (mine.view(np.uint32) != q["point_ms"].view(np.uint32)).mean() < 1e-5          # fp64 warp, last-bit rounding at most
        # the integer pipeline is bit-exact on identical points: re-run the oracle's quantisation on OUR warped points
        q2 = D.quantize_ms(mine[:n0], mine, 0.1)
        want_c.append(np.concatenate([q2["pc_ms"], np.full((len(q2["pc_ms"]), 1), b, np.int32)], 1))
        want_f.append(q2["feat_ms"])
        assert np.array_equal(out["inverse"].cpu().numpy()[row:row + n], q2["inverse_map_ms"] + n_vox)
        n_vox += len(q2["pc_ms"])
        row += n
    assert np.array_equal(out["coords"].cpu().numpy(), np.concatenate(want_c))
    assert np.array_equal(out["feats"].cpu().numpy().view(np.uint32), np.concatenate(want_f).view(np.uint32))
    torch.manual_seed(0)
    cfg = ModelCfg(IN_FEATURE_DIM=4, BLOCK="ResBlock", NUM_LAYER=[1, 1, 1, 1, 1, 1, 1, 1], cr=0.25, pres=0.1, vres=0.1,
                   IF_DIST=False, IGNORE_LABEL=0, DROPOUT_P=0.0)
    model = MinkUNetMs(cfg, 17).cuda().eval()
    with torch.no_grad():
        l32 = model.logits(frontend.as_lidar_ms(out)).cpu().numpy()
    l16 = Engine(model)(out["coords"], out["feats"], out_rows=out["cur_rows"]).cpu().numpy()
    want = l32[out["cur_rows"].cpu().numpy()]
    assert l16.shape == (sum(out["n_cur"]), 17)
    l2, agree = bf16_ok(l16, want)
    assert l2 < 3e-2 and agree > 0.99, (l2, agree)


def test_kd_twin_backbone(golden):
    """MinkUNetMsKd in training mode against the unmodified reference (tests/golden/kd.npz): teacher /
    student point features on the matched voxels, the sphashquery matching (bit-exact) and the distillation loss."""
    from taseg_b200 import SparseTensor
    from taseg_b200.segmentor import MinkUNetMsKd, ModelCfg
    g = golden("kd")
    cfg = ModelCfg(IN_FEATURE_DIM=5, BLOCK="ResBlock", NUM_LAYER=[1, 1, 1, 1, 1, 1, 1, 1], cr=0.125, pres=0.05, vres=0.05,
                   IF_DIST=False, IGNORE_LABEL=0, DROPOUT_P=0.0, SAMPLING_TYPE="random", MAX_VOXEL=10 ** 9, FEAT_KD_WEIGHT=2.0)
    model = MinkUNetMsKd(cfg, 20)
    model.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")})
    model = model.cuda().train()
    batch = dict(lidar_ms_gt=SparseTensor(cu(g["feats_t"]), cu(g["coords_t"]), 1), lidar_ms=SparseTensor(cu(g["feats_s"]), cu(g["coords_s"]), 1))
    x_gt, feat_gt = model.teacher_features(batch)
    x = batch["lidar_ms"]
    x.F = x.F[:, :5]
    from taseg_b200 import PointTensor
    feat = model.features(x, PointTensor(x.F, x.C.float()))
    loss, s2d = model.distillation_loss(x, feat, x_gt, feat_gt)
    assert sha(s2d.cpu().numpy(), np.int64) == str(g["s2d_sha"])
    mask = s2d >= 0
    assert int(mask.sum()) == int(g["n_matched"])
    assert rel_err(feat[mask].detach().cpu().numpy()[::8], g["feat_s"]) < 2e-3
    assert rel_err(feat_gt[s2d[mask]].cpu().numpy()[::8], g["feat_t"]) < 2e-3
    assert abs(float(loss.detach()) - float(g["loss_feat_kd"])) < 2e-3 * float(g["loss_feat_kd"])
    # the whole training forward (cross-entropy in place of pcseg.loss.Losses) runs and back-propagates into the student only
    labels = torch.randint(1, 20, (x.C.shape[0],), device="cuda")
    batch = dict(lidar_ms_gt=SparseTensor(cu(g["feats_t"]), cu(g["coords_t"]), 1), lidar_ms=SparseTensor(cu(g["feats_s"]), cu(g["coords_s"]), 1),
                 targets_ms=SparseTensor(labels, cu(g["coords_s"]), 1))
    ret, tb, _ = model(batch)
    ret["loss"].backward()
    assert tb["loss_feat_kd"] > 0 and model.stem[0].kernel.grad is not None and model.stem_gt[0].kernel.grad is None


def test_batched_forward_equals_per_sample(golden):
    """Batch index is a coordinate: a collated batch must give each sample the logits it gets alone (eval BN)."""
    from taseg_b200 import frontend, synth
    from taseg_b200.engine import Engine
    g = golden("net_minkunet_ms")
    model = load_model(g, "minkunet_ms")
    spec = synth.SensorSpec(16, -24.8, 2.0, 300, 1.73, 60.0)
    samples = [synth.kitti_sample(300 + b, 3, spec=spec, n_boxes=30) for b in range(3)]
    mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
    out = frontend.aggregate_voxelize(cu(mfb.points), mfb, 0.05, cu(mfb.cur_idx))
    eng = Engine(model)
    with torch.no_grad():
        full32 = model.logits(frontend.as_lidar_ms(out)).cpu().numpy()
    full16 = eng(out["coords"], out["feats"]).cpu().numpy()
    bcol = out["coords"][:, 3].cpu().numpy()
    for b, (frs, pss) in enumerate(samples):
        ms, n0 = D.aggregate_kitti(frs, pss)
        q = D.quantize_ms(ms[:n0], ms, 0.05)
        assert np.array_equal(out["coords"].cpu().numpy()[bcol == b][:, :3], q["pc_ms"])
        one = frontend.MultiFrameBatch([frs], [pss])
        o1 = frontend.aggregate_voxelize(cu(one.points), one, 0.05, cu(one.cur_idx))
        with torch.no_grad():
            alone = model.logits(frontend.as_lidar_ms(o1)).cpu().numpy()
        assert rel_err(full32[bcol == b], alone) < 1e-4
        l2, agree = bf16_ok(full16[bcol == b], alone)
        assert l2 < 3e-2 and agree > 0.99, (b, l2, agree)


def test_full_width_network_against_oracle():
    """mk34 cr1.0 channel widths (the benchmark model) on a 1/8-azimuth sector, CUDA vs the numpy oracle."""
    from taseg_b200 import frontend, synth
    from taseg_b200.engine import Engine
    from taseg_b200.segmentor import MinkUNetMs, ModelCfg
    frames, poses = synth.kitti_sample(2000, 3)
    frames = [synth.sector(f, 1 / 8) for f in frames]
    torch.manual_seed(0)
    cfg = ModelCfg(IN_FEATURE_DIM=5, BLOCK="ResBlock", NUM_LAYER=[2, 3, 4, 6, 2, 2, 2, 2], cr=1.0, IF_DIST=False,
                   IGNORE_LABEL=0, DROPOUT_P=0.0)
    model = MinkUNetMs(cfg, 20).cuda().eval()
    gen = torch.Generator().manual_seed(1)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=gen) + 0.5)
    mfb = frontend.MultiFrameBatch([frames], [poses])
    out = frontend.aggregate_voxelize(cu(mfb.points), mfb, 0.05, cu(mfb.cur_idx))
    with torch.no_grad():
        l32 = model.logits(frontend.as_lidar_ms(out)).cpu().numpy()
    l16 = Engine(model)(out["coords"], out["feats"]).cpu().numpy()
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    want = N.Net(sd).minkunet_ms(out["coords"].cpu().numpy(), out["feats"].cpu().numpy())
    assert rel_err(l32, want) < 1e-3
    l2, agree = bf16_ok(l16, want)
    assert l2 < 3e-2 and agree > 0.99, (l2, agree)


def test_full_size_batch4_against_reference(golden):
    """BASELINE configs[1] at FULL size — the exact batch bench.py times (seeds 2000..2003, 3 frames, 702 k voxels,
    MinkUNetMs mk34 cr1.0, bench.make_model weights) — against the reference run in the build container
    (tests/golden/make_golden_full.py: compiled reference CPU backend, one scan at a time): voxel counts and coordinate
    hashes bit-exact, bf16 engine logits within 3e-2 / 99 % arg-max, fp32 module path within 1e-3."""
    import bench
    from taseg_b200 import frontend
    from taseg_b200.engine import Engine
    g = golden("full_cfg1")
    step = int(g["step"])
    model = bench.make_model()
    samples = bench.make_samples(2000, bench.BATCH)
    mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
    pts, cur_idx = cu(mfb.points), cu(mfb.cur_idx)
    out = frontend.aggregate_voxelize(pts, mfb, bench.VOXEL, cur_idx)
    coords = out["coords"].cpu().numpy()
    for b in range(bench.BATCH):
        cb = coords[coords[:, 3] == b][:, :3]
        assert len(cb) == int(g[f"n_vox_{b}"]) and mfb.n_cur[b] == int(g[f"n_cur_{b}"])
        assert sha(cb, np.int32) == str(g[f"coords_sha_{b}"]), b
    logits = Engine(model)(out["coords"], out["feats"], field_bits=out["field_bits"], out_rows=out["cur_rows"]).cpu().numpy()
    assert logits.shape == (sum(mfb.n_cur), 20)
    with torch.no_grad():
        l32 = model.logits(frontend.as_lidar_ms(out))[out["cur_rows"].long()].cpu().numpy()
    row = 0
    for b in range(bench.BATCH):
        want = g[f"logits_{b}"]
        l2, agree = bf16_ok(logits[row:row + mfb.n_cur[b]][::step], want)
        assert l2 < 3e-2 and agree > 0.99, (b, l2, agree)
        assert rel_err(l32[row:row + mfb.n_cur[b]][::step], want) < 1e-3, b
        row += mfb.n_cur[b]


def test_full_size_properties():
    """BASELINE config-2 size (3 frames, full scan): size-independent invariants of the integer path and the convs."""
    from taseg_b200 import frontend, ops, synth
    from taseg_b200.nn.utils.kernel import kernel_offsets_np
    frames, poses = synth.kitti_sample(2001, 3)
    mfb = frontend.MultiFrameBatch([frames, frames], [poses, poses])
    out = frontend.aggregate_voxelize(cu(mfb.points), mfb, 0.05, cu(mfb.cur_idx))
    c = out["coords"]
    n = c.shape[0]
    # sortedness + uniqueness of the voxel list (lexicographic b,x,y,z) and inverse/inds consistency
    key = ((c[:, 3].long() << 57) | (c[:, 0].long() << 38) | (c[:, 1].long() << 19) | c[:, 2].long())
    assert bool((key[1:] > key[:-1]).all())
    assert torch.equal(out["pc_ms"][out["inds"].long()], c) and torch.equal(c[out["inverse"].long()], out["pc_ms"])
    assert int(c[:, :3].min()) == 0
    half = n // 2                                          # identical samples -> identical halves
    assert torch.equal(c[:half, :3], c[half:, :3])
    # kernel map: centre offset is the identity, offset k and 26-k are mutually transposed
    tab = ops.Table.from_coords(c)
    km = ops.build_kmap(tab, n, c, kernel_offsets_np(3, 1))
    ar = torch.arange(n, device="cuda", dtype=torch.int32)
    assert torch.equal(km.nbr[13], ar)
    nt = km.nbr_t
    for k in (0, 5, 12):
        assert torch.equal(nt[k], km.nbr[26 - k])
    assert int(km.nbsizes32.sum()) == int((km.nbr >= 0).sum())
    # idempotence of the downsample and containment of parents
    c2 = ops.unique_coords(c, trunc_stride=2)
    assert torch.equal(ops.unique_coords(c2, trunc_stride=2), c2)
    par = ops.Table.from_coords(c2)
    km2 = ops.build_kmap(par, c2.shape[0], c2, kernel_offsets_np(1, 1))
    assert int((km2.nbr >= 0).sum()) == c2.shape[0]
    km_down = ops.build_kmap(tab, n, c2, kernel_offsets_np(2, 1))
    assert int(km_down.nbsizes32.sum()) == n               # every fine voxel has exactly one parent slot
    # convolution linearity on the tensor-core path: conv(a*x1 + x2) == a*conv(x1) + conv(x2) up to bf16
    torch.manual_seed(0)
    w = torch.randn(27, 32, 32, device="cuda") * 0.1
    packed = ops.pack_weights(w, 32)
    x1 = torch.randn(n, 32, device="cuda").bfloat16()
    x2 = torch.randn(n, 32, device="cuda").bfloat16()
    f = lambda x: ops.conv_forward_tc(x, None, packed, 27, 32, km.nbr, km.tile_mask(), n, out_dtype=torch.float32)
    lhs = f((2 * x1.float() + x2.float()).bfloat16())
    rhs = 2 * f(x1) + f(x2)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 2e-2
    assert torch.equal(f(x1)[:half], f(x1)[:half]) and float((f(torch.zeros_like(x1))).abs().max()) == 0.0
