"""GPU tests of the sync-free, graph-captured pipeline (taseg_b200/pipeline.py): bit-identical to the eager engine
(which the other GPU tests pin against the reference), robust to a change of the data under a captured graph, and loud
on capacity overflow."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make(seed0, n_samples, spec=None, num_layer=(1, 1, 1, 1, 1, 1, 1, 1), cr=0.25):
    from taseg_b200 import frontend, synth
    from taseg_b200.engine import Engine
    from taseg_b200.segmentor import MinkUNetMs, ModelCfg
    spec = spec or synth.SensorSpec(16, -24.8, 2.0, 300, 1.73, 60.0)
    samples = [synth.kitti_sample(seed0 + b, 3, spec=spec, n_boxes=30) for b in range(n_samples)]
    mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
    torch.manual_seed(0)
    cfg = ModelCfg(IN_FEATURE_DIM=5, BLOCK="ResBlock", NUM_LAYER=list(num_layer), cr=cr, IF_DIST=False, IGNORE_LABEL=0,
                   DROPOUT_P=0.0)
    model = MinkUNetMs(cfg, 20).cuda().eval()
    return mfb, Engine(model)


def eager(engine, mfb, pts):
    from taseg_b200 import frontend
    out = frontend.aggregate_voxelize(pts, mfb, 0.05, cu(mfb.cur_idx))
    sizes = [out["point_ms"].shape[0], out["coords"].shape[0]]
    return engine(out["coords"], out["feats"], field_bits=out["field_bits"], out_rows=out["cur_rows"]), sizes


def test_pipeline_matches_eager_and_replays():
    from taseg_b200.pipeline import Pipeline
    mfb, engine = make(700, 2)
    pts = cu(mfb.points)
    want, sizes = eager(engine, mfb, pts)
    pipe = Pipeline(engine, mfb, 0.05)
    learnt = pipe.calibrate(pts)
    assert learnt["kept"] == sizes[0] and learnt["levels"][0] == sizes[1]
    got = pipe(pts).clone()                       # sync-free kernels, launched one by one
    assert pipe.check() == 0 and pipe.level_counts()[:2] == sizes
    assert torch.equal(got, want)
    pipe.capture()
    got_g = pipe(pts).clone()                     # one graph launch
    assert pipe.check() == 0 and torch.equal(got_g, want)
    # new data under the SAME graph: jitter the points (voxel counts change, within the capacity margin)
    g = torch.Generator(device="cuda").manual_seed(1)
    pts2 = pts.clone()
    pts2[:, :3] += torch.randn(pts.shape[0], 3, device="cuda", generator=g) * 0.03
    want2, sizes2 = eager(engine, mfb, pts2)
    assert sizes2[1] != sizes[1]
    got2 = pipe(pts2).clone()
    assert pipe.check() == 0 and pipe.level_counts()[:2] == sizes2
    assert torch.equal(got2, want2)
    # and back: no state leaks from one replay into the next
    assert torch.equal(pipe(pts), want)


def test_two_pipelines_replay_concurrently():
    """Two captured graphs in flight on two streams (the benchmark's configuration) must not share any state — each owns
    its counters, status word and dynamic-tile-scheduler tickets."""
    from taseg_b200.pipeline import Pipeline
    mfb, engine = make(720, 2)
    pts = cu(mfb.points)
    want, _ = eager(engine, mfb, pts)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    pipes = []
    for st in streams:
        with torch.cuda.stream(st):
            p = Pipeline(engine, mfb, 0.05)
            p.calibrate(pts)
            p.points.copy_(pts)
            p.capture()
            pipes.append(p)
    torch.cuda.synchronize()
    for rep in range(6):
        for p, st in zip(pipes, streams):
            with torch.cuda.stream(st):
                p()
    torch.cuda.synchronize()
    for p in pipes:
        assert p.check() == 0 and torch.equal(p.logits, want)


def test_pipeline_flags_capacity_overflow():
    from taseg_b200.pipeline import OVERFLOW, Pipeline
    mfb, engine = make(710, 1)
    pts = cu(mfb.points)
    pipe = Pipeline(engine, mfb, 0.05, margin=1.0)
    pipe.calibrate(pts)
    pipe.caps[2] = max(256, pipe.caps[2] // 2 // 256 * 256)      # level 2 cannot hold its voxels any more
    pipe.caps[3] = min(pipe.caps[3], pipe.caps[2])
    pipe.caps[4] = min(pipe.caps[4], pipe.caps[3])
    pipe(pts)
    assert pipe.check() & OVERFLOW
    assert pipe.level_counts()[3] == pipe.caps[2]                  # clamped, never past the buffers


def test_pipeline_full_size_bench_batch():
    """The benchmark batch (4 x 3-frame full scans) through the captured graph equals the eager engine bit for bit."""
    import bench
    from taseg_b200 import frontend
    from taseg_b200.engine import Engine
    from taseg_b200.pipeline import Pipeline
    engine = Engine(bench.make_model())
    samples = bench.make_samples(2000, bench.BATCH)
    mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
    pts = cu(mfb.points)
    want, _ = eager(engine, mfb, pts)
    pipe = Pipeline(engine, mfb, bench.VOXEL)
    pipe.calibrate(pts)
    pipe.capture()
    got = pipe(pts)
    assert pipe.check() == 0
    assert torch.equal(got, want)
