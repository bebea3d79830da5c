#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run once here (needs /root/reference and oracle/_ref); the fixtures are committed because the
reference cannot travel to the GPU box.

What is imported from the reference, unmodified:
  * torchsparse 1.4.0 Python (extracted from /root/reference/package/torchsparse.zip to a scratch dir)
    with `torchsparse.backend` = oracle/_ref (the reference CPU backend compiled by oracle/build_ref.py);
  * pcseg MinkUNet / MinkUNetMs / SPVCNN from /root/reference/pcseg/model/segmentor/** via empty package
    shells (SURVEY §8c recipe step 4);
  * SemantickittiMsDataset.fuse_multi_scan (unbound) with petrel_client stubbed.
All inputs are batch index 0 only: the reference's CPU kernel_hash mishandles batch>0
(TS/backend/hash/hash_cpu.cpp:29).
"""
import hashlib
import os
import sys
import tempfile
import types
import zipfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def import_reference():
    from oracle import ref_backend
    scratch = tempfile.mkdtemp(prefix="taseg_golden_")
    zipfile.ZipFile(os.path.join(REF, "package", "torchsparse.zip")).extractall(scratch)
    sys.path.insert(0, os.path.join(scratch, "torchsparse"))
    sys.modules["torchsparse.backend"] = ref_backend.backend()
    import torchsparse  # noqa: F401  reference python
    torchsparse.backend = sys.modules["torchsparse.backend"]
    sys.path.insert(0, REF)
    for name, sub in [("pcseg", "pcseg"), ("pcseg.model", "pcseg/model"),
                      ("pcseg.model.segmentor", "pcseg/model/segmentor"),
                      ("pcseg.model.segmentor.voxel", "pcseg/model/segmentor/voxel"),
                      ("pcseg.model.segmentor.voxel.minkunet", "pcseg/model/segmentor/voxel/minkunet"),
                      ("pcseg.model.segmentor.fusion", "pcseg/model/segmentor/fusion"),
                      ("pcseg.model.segmentor.fusion.spvcnn", "pcseg/model/segmentor/fusion/spvcnn"),
                      ("pcseg.data", "pcseg/data"), ("pcseg.data.dataset", "pcseg/data/dataset"),
                      ("pcseg.data.dataset.semantickitti", "pcseg/data/dataset/semantickitti")]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = m
    sys.modules["petrel_client"] = types.ModuleType("petrel_client")
    sys.modules["petrel_client.client"] = types.ModuleType("petrel_client.client")
    sys.modules["petrel_client.client"].Client = object
    return torchsparse


def small_sample(seed, n_frames):
    from taseg_b200 import synth
    spec = synth.SensorSpec(16, -24.8, 2.0, 300, 1.73, 60.0)
    return synth.kitti_sample(seed, n_frames, spec=spec, n_boxes=40)


def randomize_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def model_cfg(in_dim, cr, num_layer):
    return AttrDict(IN_FEATURE_DIM=in_dim, BLOCK="ResBlock", NUM_LAYER=num_layer, cr=cr,
                    PLANES=[32, 32, 64, 128, 256, 256, 128, 96, 96], pres=0.05, vres=0.05,
                    DROPOUT_P=0.0, LABEL_SMOOTHING=0.0, IF_DIST=False, IGNORE_LABEL=0)


def run_net(kind, ts):
    """Returns the fixture dict for one network."""
    from oracle import data_oracle as D
    from torchsparse import SparseTensor
    from torchsparse.utils.collate import sparse_collate
    if kind == "minkunet_ms":
        from pcseg.model.segmentor.voxel.minkunet.minkunet_ms import MinkUNetMs as Model
        in_dim, n_frames = 5, 3
    elif kind == "minkunet":
        from pcseg.model.segmentor.voxel.minkunet.minkunet import MinkUNet as Model
        in_dim, n_frames = 4, 1
    else:
        from pcseg.model.segmentor.fusion.spvcnn.spvcnn import SPVCNN as Model
        in_dim, n_frames = 4, 1
    frames, poses = small_sample(7000 + len(kind), n_frames)
    torch.manual_seed(0)
    model = Model(model_cfg(in_dim, 0.125, [1, 2, 1, 1, 1, 1, 1, 1]), 20)
    randomize_bn(model, 1)
    model.eval()

    out = {}
    feats_log = {}

    def hook(name):
        def fn(mod, inp, res):
            feats_log[name] = res
        return fn
    for name in ["stem", "stage1", "stage2", "stage3", "stage4"]:
        getattr(model, name).register_forward_hook(hook(name))
    for name in ["up1", "up2", "up3", "up4"]:
        getattr(model, name)[1].register_forward_hook(hook(name))
    model.classifier.register_forward_hook(hook("logits"))

    if kind == "minkunet_ms":
        ms, n0 = D.aggregate_kitti(frames, poses)
        q = D.quantize_ms(ms[:n0], ms, 0.05)
        lidar = sparse_collate([SparseTensor(torch.from_numpy(q["feat_ms"]), torch.from_numpy(q["pc_ms"]))])
        inv = sparse_collate([SparseTensor(torch.from_numpy(q["inverse_map_ms"]), torch.from_numpy(q["pc_ms_"]))])
        lab = sparse_collate([SparseTensor(torch.zeros(n0), torch.from_numpy(q["pc_ms_"][:n0]))])
        n_ms = len(q["pc_ms_"])
        pm = torch.zeros(n_ms, dtype=torch.bool)
        pm[:n0] = True
        batch = dict(lidar_ms=lidar, inverse_map_ms=inv, targets_mapped=lab, point_mask=pm,
                     num_points=torch.tensor([[n0]]), num_points_ms=torch.tensor([[n_ms]]), name=["s"])
        out.update(frames=np.concatenate(frames), frame_sizes=np.array([len(f) for f in frames]),
                   poses=np.stack(poses), xyzret_ms_sha=np.array(hashlib.sha256(np.ascontiguousarray(ms).tobytes()).hexdigest()), coords=lidar.C.numpy().copy(), feats=lidar.F.numpy().copy(),
                   inds=q["inds_ms"], inverse=q["inverse_map_ms"], n_current=n0)
    else:
        q = D.quantize_single(frames[0], 0.05)
        lidar = sparse_collate([SparseTensor(torch.from_numpy(q["feat"]), torch.from_numpy(q["pc"]))])
        inv = sparse_collate([SparseTensor(torch.from_numpy(q["inverse_map"]), torch.from_numpy(q["pc_"]))])
        lab = sparse_collate([SparseTensor(torch.zeros(len(q["pc_"])), torch.from_numpy(q["pc_"]))])
        batch = dict(lidar=lidar, inverse_map=inv, targets_mapped=lab, num_points=torch.tensor([[len(q["pc_"])]]), name=["s"])
        out.update(points=frames[0], coords=lidar.C.numpy().copy(), feats=lidar.F.numpy().copy(),
                   inds=q["inds"], inverse=q["inverse_map"])
    with torch.no_grad():
        res = model(batch)
    out["point_logits"] = res["point_predict_logits"][0]
    out["voxel_logits"] = feats_log["logits"].numpy()[::4]   # every 4th voxel row
    last = None
    for name in ["stem", "stage1", "stage2", "stage3", "stage4", "up1", "up2", "up3", "up4"]:
        st = feats_log[name]
        step = 8 if st.F.shape[0] > 2000 else 1       # row-subsampled to keep the fixture small
        out[f"F_{name}"] = st.F.numpy()[::step]
        out[f"Fstep_{name}"] = np.array(step)
        out[f"Csha_{name}"] = np.array(hashlib.sha256(np.ascontiguousarray(st.C.numpy()).tobytes()).hexdigest())
        out[f"Cn_{name}"] = np.array(st.C.shape[0])
        last = st
    for key, (nbmaps, nbsizes, sizes) in last.kmaps.items():
        tag = "kmap_s%d_k%d_st%d" % (key[0][0], key[1][0], key[2][0])
        out[tag + "_nbmaps_sha"] = np.array(hashlib.sha256(np.ascontiguousarray(nbmaps.numpy().astype(np.int64)).tobytes()).hexdigest())
        out[tag + "_nbsizes"] = nbsizes.numpy()
        out[tag + "_sizes"] = np.array(sizes)
    for k, v in model.state_dict().items():
        out["sd/" + k] = v.numpy()
    return out


def run_ops(ts):
    """Known-answer vectors for every native entry point, produced by the reference Python + backend."""
    import torchsparse.nn.functional as F
    from torchsparse import SparseTensor
    from torchsparse.nn.utils import get_kernel_offsets
    from torchsparse.utils.quantize import sparse_quantize
    rng = np.random.default_rng(42)
    out = {}
    pts = rng.integers(-6, 22, (1800, 3)).astype(np.int32)
    c3, inds, inv = sparse_quantize(pts, 1, return_index=True, return_inverse=True)
    out.update(q_in=pts, q_coords=c3, q_inds=inds, q_inv=inv)
    coords = torch.from_numpy(np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1))
    out["coords"] = coords.numpy()
    out["hash"] = F.sphash(coords).numpy()
    out["hash_kat_in"] = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [-1, -1, -1, 0], [3, 2, 1, 5], [2147483647, -2147483648, 7, 9]], np.int32)
    out["hash_kat"] = F.sphash(torch.from_numpy(out["hash_kat_in"])).numpy()
    for ks, st in [(3, 1), (2, 1), (2, 2), (3, 2)]:
        out[f"offsets_k{ks}_s{st}"] = get_kernel_offsets(ks, st).numpy()
    out["khash27"] = F.sphash(coords, get_kernel_offsets(3, 1)).numpy()
    out["query27"] = F.sphashquery(torch.from_numpy(out["khash27"]), torch.from_numpy(out["hash"])).numpy()
    torch.manual_seed(3)
    x = SparseTensor(torch.randn(len(coords), 5), coords, 1)
    w3 = torch.randn(27, 5, 8) * 0.2
    y = F.conv3d(x, w3, 3)
    out.update(conv_in=x.F.numpy(), conv_w3=w3.numpy(), conv_out3=y.F.numpy())
    nb, ns, sz = x.kmaps[((1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    out.update(kmap3_nbmaps=nb.numpy(), kmap3_nbsizes=ns.numpy())
    w2 = torch.randn(8, 8, 16) * 0.2
    y2 = F.conv3d(y, w2, 2, stride=2)
    out.update(conv_w2=w2.numpy(), conv_out2=y2.F.numpy(), coords_s2=y2.C.numpy())
    nb, ns, sz = x.kmaps[((1, 1, 1), (2, 2, 2), (2, 2, 2), (1, 1, 1))]
    out.update(kmap2_nbmaps=nb.numpy(), kmap2_nbsizes=ns.numpy())
    w3b = torch.randn(27, 16, 16) * 0.2
    y3 = F.conv3d(y2, w3b, 3)
    out.update(conv_w3b=w3b.numpy(), conv_out3b=y3.F.numpy())
    nb, ns, sz = x.kmaps[((2, 2, 2), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    out.update(kmap3s2_nbmaps=nb.numpy(), kmap3s2_nbsizes=ns.numpy())
    y4 = F.conv3d(y3, torch.randn(16, 12) * 0.2, 1)
    wt = torch.randn(8, 12, 8) * 0.2
    torch.manual_seed(4)
    w1 = torch.randn(16, 12) * 0.2
    y4 = F.conv3d(y3, w1, 1)
    y5 = F.conv3d(y4, wt, 2, stride=2, transposed=True)
    out.update(conv_w1=w1.numpy(), conv_out1=y4.F.numpy(), conv_wt=wt.numpy(), conv_outT=y5.F.numpy())
    # backward through the 3^3 conv and the strided conv
    xi = x.F.clone().requires_grad_(True)
    w3g = w3.clone().requires_grad_(True)
    xx = SparseTensor(xi, coords, 1)
    xx.kmaps, xx.cmaps = x.kmaps, x.cmaps
    g = torch.randn(len(coords), 8)
    F.conv3d(xx, w3g, 3).F.backward(g)
    out.update(bwd_gy=g.numpy(), bwd_gx=xi.grad.numpy(), bwd_gw=w3g.grad.numpy())
    yi = y.F.detach().clone().requires_grad_(True)
    w2g = w2.clone().requires_grad_(True)
    yy = SparseTensor(yi, coords, 1)
    yy.kmaps, yy.cmaps = x.kmaps, x.cmaps
    g2 = torch.randn(y2.F.shape)
    F.conv3d(yy, w2g, 2, stride=2).F.backward(g2)
    out.update(bwd2_gy=g2.numpy(), bwd2_gx=yi.grad.numpy(), bwd2_gw=w2g.grad.numpy())
    # point <-> voxel
    p = torch.from_numpy(np.concatenate([rng.uniform(0, 16, (1500, 3)), np.zeros((1500, 1))], 1).astype(np.float32))
    for s, cs in [(1, coords), (2, y2.C)]:
        fl = torch.cat([torch.floor(p[:, :3] / s).int() * s, p[:, -1].int().view(-1, 1)], 1)
        off = get_kernel_offsets(2, s, 1)
        idx = F.sphashquery(F.sphash(fl, off), F.sphash(cs))
        wts = F.calc_ti_weights(p, idx, scale=s).transpose(0, 1).contiguous()
        idx = idx.transpose(0, 1).contiguous()
        feat = torch.randn(len(cs), 6)
        out[f"dv{s}_idx"] = idx.numpy()
        out[f"dv{s}_w"] = wts.numpy()
        out[f"dv{s}_feat"] = feat.numpy()
        out[f"dv{s}_out"] = F.spdevoxelize(feat, idx, wts).numpy()
        iq = F.sphashquery(F.sphash(fl), F.sphash(cs))
        cnt = F.spcount(iq.int(), len(cs))
        pf = torch.randn(len(p), 6)
        ok = iq >= 0  # reference voxelize indexes counts[idx] without a bounds check; feed hits only
        out[f"vx{s}_idx"] = iq[ok].numpy()
        out[f"vx{s}_cnt"] = cnt.numpy()
        out[f"vx{s}_feat"] = pf[ok].numpy()
        out[f"vx{s}_out"] = F.spvoxelize(pf[ok], iq[ok], cnt).numpy()
    out["pv_points"] = p.numpy()
    return out


def run_fuse():
    from pcseg.data.dataset.semantickitti.semantickitti_ms import SemantickittiMsDataset as DS
    from oracle import data_oracle as D
    frames, poses = small_sample(11, 3)
    frames = [f[:2500] for f in frames]
    out = {"points": frames[1], "pose0": poses[0], "pose": poses[1]}
    ref = DS.fuse_multi_scan(None, frames[1], poses[0], poses[1])
    out["fused"] = ref
    mine = D.fuse_multi_scan(frames[1], poses[0], poses[1])
    assert ref.dtype == np.float32 and np.array_equal(ref.view(np.uint32), mine.view(np.uint32)), "oracle fuse_multi_scan != reference"
    # a harder pose (full 3D rotation + translation), still float32 like load_calib_poses (:346)
    rng = np.random.default_rng(5)
    A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    P = np.eye(4)
    P[:3, :3] = A
    P[:3, 3] = rng.normal(size=3) * 20
    P0 = np.eye(4)
    P0[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    P0[:3, 3] = rng.normal(size=3) * 20
    ref2 = DS.fuse_multi_scan(None, frames[2], P0.astype(np.float32), P.astype(np.float32))
    mine2 = D.fuse_multi_scan(frames[2], P0.astype(np.float32), P.astype(np.float32))
    assert np.array_equal(ref2.view(np.uint32), mine2.view(np.uint32)), "oracle fuse_multi_scan != reference (3D pose)"
    out.update(points2=frames[2], pose0_2=P0.astype(np.float32), pose_2=P.astype(np.float32), fused2=ref2)
    ms = DS.append_time_flag(None, frames[0], np.concatenate([frames[0], ref], 0))
    out["time_flag_ms"] = ms
    return out


def main():
    ts = import_reference()
    np.savez_compressed(os.path.join(HERE, "ops_kat.npz"), **run_ops(ts))
    np.savez_compressed(os.path.join(HERE, "fuse_kat.npz"), **run_fuse())
    for kind in ["minkunet_ms", "minkunet", "spvcnn"]:
        d = run_net(kind, ts)
        np.savez_compressed(os.path.join(HERE, f"net_{kind}.npz"), **d)
        print(kind, d["coords"].shape, d["point_logits"].shape, float(np.abs(d["point_logits"]).max()))


if __name__ == "__main__":
    main()
