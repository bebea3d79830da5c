"""One small pass of the whole hot path on cuda:0, checked against the CPU oracle (used by __graft_entry__.smoke)."""
import numpy as np
import torch


def run(seed: int = 5, cr: float = 0.25) -> float:
    from oracle import data_oracle as D
    from oracle import net_oracle as N
    from oracle import ts_oracle as T
    from . import frontend, ops, synth
    from .engine import Engine
    from .segmentor import MinkUNetMs, ModelCfg

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    spec = synth.SensorSpec(16, -24.8, 2.0, 256, 1.73, 60.0)
    frames, poses = synth.kitti_sample(seed, 3, spec=spec, n_boxes=30)
    torch.manual_seed(0)
    cfg = ModelCfg(IN_FEATURE_DIM=5, BLOCK="ResBlock", NUM_LAYER=[1, 1, 1, 1, 1, 1, 1, 1], cr=cr, IF_DIST=False,
                   IGNORE_LABEL=0, DROPOUT_P=0.0)
    model = MinkUNetMs(cfg, 20).cuda().eval()
    g = torch.Generator().manual_seed(1)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    # CUDA path: device front end -> fp32 module graph (exact parity mode) and the fused bf16 engine
    mfb = frontend.MultiFrameBatch([frames], [poses])
    out = frontend.aggregate_voxelize(torch.from_numpy(mfb.points).cuda(), mfb, 0.05, torch.from_numpy(mfb.cur_idx).cuda())
    with torch.no_grad():
        logits = model.logits(frontend.as_lidar_ms(out))
        logits_bf16 = Engine(model)(out["coords"], out["feats"])
    got = ops.gather_rows(logits, out["cur_rows"]).cpu().numpy()
    got16 = ops.gather_rows(logits_bf16.contiguous(), out["cur_rows"]).cpu().numpy()
    # oracle
    ms, n0 = D.aggregate_kitti(frames, poses)
    q = D.quantize_ms(ms[:n0], ms, 0.05)
    coords, feats = T.sparse_collate([q["pc_ms"]], [q["feat_ms"]])
    assert np.array_equal(out["coords"].cpu().numpy(), coords), "voxel coordinates differ from the oracle"
    assert np.array_equal(out["feats"].cpu().numpy(), feats), "voxel features differ from the oracle"
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    want = N.Net(sd).minkunet_ms(coords, feats)[q["inverse_map_ms"]][:n0]
    err = float(np.abs(got - want).max() / np.abs(want).max())
    err16 = float(np.linalg.norm(got16 - want) / np.linalg.norm(want))
    print("smoke: fp32 module path rel err %.2e ; bf16 engine rel-l2 %.2e, argmax agreement %.4f" %
          (err, err16, float((got16.argmax(1) == want.argmax(1)).mean())))
    assert err16 < 3e-2
    return err
