"""Training step of the hot path: autograd through the CUDA kernels, bf16 autocast, gradient reducer."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(seed=0, n=6000, span=48, n_cls=5):
    import taseg_b200 as ts
    rng = np.random.default_rng(seed)
    xyz = rng.normal(0, span / 6, (n, 3))
    xyz[:, 2] *= 0.2
    c = np.unique(np.round(xyz).astype(np.int32), axis=0)
    c -= c.min(0)
    coords = np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1)
    feats = rng.normal(size=(len(c), 5)).astype(np.float32)
    labels = 1 + (c[:, 0] > c[:, 0].mean()).astype(np.int64) * 2 + (c[:, 1] > c[:, 1].mean())   # learnable from position
    dev = "cuda"
    return {"lidar_ms": ts.SparseTensor(torch.from_numpy(feats).to(dev), torch.from_numpy(coords).to(dev)),
            "targets_ms": ts.SparseTensor(torch.from_numpy(labels).to(dev), torch.from_numpy(coords).to(dev))}, len(c)


def _model(n_cls=5, planes=16):
    from taseg_b200.segmentor import MinkUNetMs, ModelCfg
    torch.manual_seed(0)
    cfg = ModelCfg(IN_FEATURE_DIM=5, BLOCK="ResBlock", NUM_LAYER=[1, 1, 1, 1, 1, 1, 1, 1], cr=1.0,
                   PLANES=[planes, planes, 2 * planes, 2 * planes, 4 * planes, 4 * planes, 2 * planes, planes, planes],
                   pres=1.0, vres=1.0, IF_DIST=False, IGNORE_LABEL=0, DROPOUT_P=0.0)
    return MinkUNetMs(cfg, n_cls).cuda()


def _fresh(batch):
    import taseg_b200 as ts
    x, t = batch["lidar_ms"], batch["targets_ms"]
    return {"lidar_ms": ts.SparseTensor(x.F.clone(), x.C), "targets_ms": ts.SparseTensor(t.F, t.C)}


def test_training_steps_reduce_the_loss():
    from taseg_b200 import parallel
    batch, _ = _batch()
    model = _model()
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    reducer = parallel.GradientReducer(model.parameters())      # world size 1: no exchange
    losses = [parallel.train_step(model, _fresh(batch), opt, reducer, amp_dtype=torch.bfloat16) for _ in range(10)]
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < 0.9 * losses[0], losses
    assert reducer.bytes_per_step() == 4 * sum(p.numel() for p in model.parameters())


def test_bf16_gradients_agree_with_fp32():
    batch, _ = _batch(seed=1)
    grads = {}
    for name, amp in (("fp32", None), ("bf16", torch.bfloat16)):
        model = _model()
        model.train()
        with torch.autocast("cuda", dtype=amp or torch.bfloat16, enabled=amp is not None):
            loss = model(_fresh(batch))[0]["loss"]
        loss.backward()
        grads[name] = {k: p.grad.float().flatten() for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads["fp32"]) == set(grads["bf16"])
    cos = {k: float(torch.nn.functional.cosine_similarity(grads["fp32"][k], grads["bf16"][k], dim=0)) for k in grads["fp32"]
           if grads["fp32"][k].norm() > 0}
    worst = min(cos, key=cos.get)
    # stated bf16 bound: bf16 activations and bf16 dgrad (fp32 accumulate) through ~25 layers on a 3 k-voxel batch
    assert cos[worst] > 0.8 and float(np.mean(list(cos.values()))) > 0.97, (worst, cos[worst], np.mean(list(cos.values())))


def test_fp32_gradient_against_finite_differences():
    batch, _ = _batch(seed=2, n=1500, span=24)
    model = _model(planes=8)
    model.train()
    for m in model.modules():                      # freeze BN statistics so that the loss is a smooth function of one weight
        if isinstance(m, torch.nn.BatchNorm1d):
            m.eval()
    loss = model(_fresh(batch))[0]["loss"]
    loss.backward()
    w = model.stem[3].kernel                        # (27, 8, 8)
    picks = [(13, 0, 0), (4, 3, 5), (22, 7, 1)]
    eps = 2e-2
    for k, i, j in picks:
        with torch.no_grad():
            orig = float(w[k, i, j])
            w[k, i, j] = orig + eps
            lp = float(model(_fresh(batch))[0]["loss"])
            w[k, i, j] = orig - eps
            lm = float(model(_fresh(batch))[0]["loss"])
            w[k, i, j] = orig
        fd = (lp - lm) / (2 * eps)
        an = float(w.grad[k, i, j])
        assert abs(fd - an) <= 5e-2 * max(abs(fd), abs(an)) + 2e-4, ((k, i, j), fd, an)


@pytest.mark.parametrize("c,dtype", [(32, torch.float32), (96, torch.bfloat16), (256, torch.bfloat16), (8, torch.float32)])
def test_batch_norm_rows_vs_aten(c, dtype):
    """spnn.BatchNorm on CUDA rows (csrc/bn.cu) against nn.BatchNorm1d (ATen) with the same parameters: training forward,
    running statistics, input / weight / bias gradients, then evaluation mode; and run-to-run determinism."""
    from taseg_b200 import nn as spnn
    from taseg_b200 import ops
    torch.manual_seed(c)
    n = 70001
    x = (torch.randn(n, c, device="cuda") * 1.7 + 0.3).to(dtype)
    ours, ref = spnn.BatchNorm(c).cuda(), torch.nn.BatchNorm1d(c).cuda()
    with torch.no_grad():
        ours.weight.copy_(torch.rand(c) + 0.5)
        ours.bias.copy_(torch.randn(c) * 0.2)
    ref.load_state_dict(ours.state_dict())
    gy = torch.randn(n, c, device="cuda").to(dtype)
    outs = []
    for m in (ours, ref):
        xi = x.clone().requires_grad_(True)
        y = m._rows(xi) if m is ours else m(xi)
        y.backward(gy)
        outs.append((y.detach().float(), xi.grad.float(), m.weight.grad.clone(), m.bias.grad.clone()))
    assert ops.bn_supported(x)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    for a, b in zip(outs[0], outs[1]):
        assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max())), (c, dtype)
    assert torch.allclose(ours.running_mean, ref.running_mean, atol=1e-5) and torch.allclose(ours.running_var, ref.running_var, rtol=1e-4)
    assert int(ours.num_batches_tracked) == 1
    xi = x.clone().requires_grad_(True)
    y2 = ours._rows(xi)
    y2.backward(gy)
    assert torch.equal(y2.detach().float(), outs[0][0]) and torch.equal(xi.grad.float(), outs[0][1]), "not deterministic"
    ours.eval(), ref.eval()
    ref.load_state_dict(ours.state_dict())
    xe = x.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ye, yr = ours._rows(xe), ref(xr)
    ye.backward(gy), yr.backward(gy)
    assert float((ye.detach().float() - yr.detach().float()).abs().max()) <= tol * max(1.0, float(yr.detach().float().abs().max()))
    assert float((xe.grad.float() - xr.grad.float()).abs().max()) <= tol * max(1.0, float(xr.grad.float().abs().max()))


@pytest.mark.parametrize("c,dtype", [(32, torch.float32), (96, torch.bfloat16)])
def test_batch_norm_relu_fused(c, dtype):
    """BatchNorm with the ReLU of its convolution block applied in the same pass (norm.fuse_bn_relu) against
    relu(nn.BatchNorm1d(x)): outputs and all gradients in training mode, outputs in evaluation mode; the ReLU module behind a
    fused BatchNorm is the identity; module / parameter names are unchanged."""
    from taseg_b200 import nn as spnn
    from taseg_b200.nn.modules.norm import fuse_bn_relu
    torch.manual_seed(c + 1)
    n = 50003
    x = (torch.randn(n, c, device="cuda") * 1.3 + 0.2).to(dtype)
    seq = fuse_bn_relu(torch.nn.Sequential(spnn.BatchNorm(c), spnn.ReLU(True))).cuda()
    ours, ref = seq[0], torch.nn.BatchNorm1d(c).cuda()
    assert ours.fuse_relu and seq[1].fused_upstream and list(seq.state_dict()) == ["0." + k for k in ref.state_dict()]
    with torch.no_grad():
        ours.weight.copy_(torch.rand(c) - 0.3)          # some negative scales: the mask must come from bn(x), not from x
        ours.bias.copy_(torch.randn(c) * 0.3)
    ref.load_state_dict(ours.state_dict())
    gy = torch.randn(n, c, device="cuda").to(dtype)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = ours._rows(xa)
    yb = torch.relu(ref(xb))
    ya.backward(gy), yb.backward(gy)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    for a, b in ((ya, yb), (xa.grad, xb.grad), (ours.weight.grad, ref.weight.grad), (ours.bias.grad, ref.bias.grad)):
        a, b = a.detach().float(), b.detach().float()
        assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max())), (c, dtype)
    assert float(ya.min()) >= 0.0
    ours.eval(), ref.eval()
    ref.load_state_dict(ours.state_dict())
    with torch.no_grad():
        ye, yr = ours._rows(x), torch.relu(ref(x))
    assert float((ye.float() - yr.float()).abs().max()) <= tol * max(1.0, float(yr.float().abs().max()))


def test_batch_norm_statistics_are_stable():
    """Mean / variance of rows whose mean is 10^3 standard deviations away from zero (fp32 rows, fp64 reference): the one-pass
    combination of the per-CTA records is pivoted, so nothing cancels there (a plain E[x^2] - E[x]^2 in fp32 would be off by
    ~10 % of the variance here)."""
    from taseg_b200 import ops
    torch.manual_seed(3)
    n, c = 200003, 64
    x = (torch.randn(n, c, device="cuda", dtype=torch.float64) * 0.01 + 10.0).float()
    mean, invstd = ops.bn_stats(x, 0.0, 0.1, None, None)
    xd = x.double()
    mu, var = xd.mean(0), xd.var(0, unbiased=False)
    err_mean = float(((mean.double() - mu).abs() / 0.01).max())             # in units of the standard deviation
    err_std = float((invstd.double() * var.sqrt() - 1).abs().max())
    assert err_mean < 5e-3 and err_std < 5e-3, (err_mean, err_std)
