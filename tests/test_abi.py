"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the header
declares; the host mirror exposes the reference's import surface; host-only entry points behave."""
import ctypes
import os

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def lib():
    from taseg_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    from taseg_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 40
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.tsg_version() >= 100
    assert lib.tsg_last_error() is not None


def test_host_only_entry_points(lib):
    for n in (0, 1, 511, 512, 513, 100000):
        s = lib.tsg_table_slots(n)
        assert s >= 2 * n and s & (s - 1) == 0
    assert lib.tsg_kmap_blocks(1) == 1 and lib.tsg_kmap_blocks(1025) == 2
    assert lib.tsg_sort_ws_bytes(1000) > 1000 * 12
    assert lib.tsg_unique_ws_bytes(1000) > lib.tsg_sort_ws_bytes(1000)
    assert lib.tsg_conv_pack_bytes(27, 64, 0, 64) == 27 * 64 * 64 * 2
    assert lib.tsg_conv_pack_bytes(27, 96, 32, 96) == 27 * 2 * 96 * 64 * 2      # 128 flat channels: 2 slices per offset
    assert lib.tsg_conv_pack_bytes(27, 96, 0, 96) == 14 * 3 * 96 * 64 * 2       # 96 channels: 3 slices per 2 offsets
    assert lib.tsg_conv_pack_bytes(27, 32, 0, 32) == 14 * 1 * 32 * 64 * 2       # 32 channels: 2 offsets per slice


def test_no_cpu_fallback():
    import taseg_b200
    from taseg_b200.nn import functional as F
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.sphash(torch.zeros((4, 4), dtype=torch.int32))
    x = taseg_b200.SparseTensor(torch.zeros(4, 4), torch.zeros((4, 4), dtype=torch.int32))
    with pytest.raises(RuntimeError):
        F.conv3d(x, torch.zeros(27, 4, 8), 3)


def test_import_surface_matches_reference():
    """Names pcseg imports (SURVEY §8b 'import surface')."""
    import sys
    import taseg_b200
    taseg_b200.install_as_torchsparse()
    import torchsparse
    import torchsparse.nn as spnn
    import torchsparse.nn.functional as F
    from torchsparse import PointTensor, SparseTensor  # noqa: F401
    from torchsparse.nn.utils import fapply, get_kernel_offsets  # noqa: F401
    from torchsparse.utils import make_ntuple
    from torchsparse.utils.collate import sparse_collate, sparse_collate_fn  # noqa: F401
    from torchsparse.utils.quantize import sparse_quantize  # noqa: F401
    for name in ["conv3d", "sphash", "sphashquery", "spcount", "spvoxelize", "spdevoxelize", "calc_ti_weights",
                 "spdownsample", "relu"]:
        assert hasattr(F, name), name
    for name in ["Conv3d", "BatchNorm", "ReLU", "LeakyReLU"]:
        assert hasattr(spnn, name), name
    assert torchsparse.cat is taseg_b200.cat and make_ntuple(2, 3) == (2, 2, 2)
    conv = spnn.Conv3d(4, 8, 3)
    assert tuple(conv.kernel.shape) == (27, 4, 8) and conv.bias is None
    assert tuple(spnn.Conv3d(4, 8, 1).kernel.shape) == (4, 8)
    assert set(spnn.Conv3d(4, 8, 2, stride=2, bias=True).state_dict()) == {"kernel", "bias"}
    off = get_kernel_offsets(3)
    assert off.dtype == torch.int32 and off[0].tolist() == [-1, -1, -1] and off[1].tolist() == [0, -1, -1] and off[13].tolist() == [0, 0, 0]
    assert get_kernel_offsets(2, 4)[1].tolist() == [0, 0, 4]


def test_collate_and_containers():
    import taseg_b200
    from taseg_b200.utils.collate import sparse_collate_fn
    a = taseg_b200.SparseTensor(np.ones((3, 2), np.float32), np.zeros((3, 3), np.int32))
    b = taseg_b200.SparseTensor(np.ones((2, 2), np.float32), np.ones((2, 3), np.int32))
    out = sparse_collate_fn([{"lidar": a, "n": np.array([3]), "name": "a"}, {"lidar": b, "n": np.array([2]), "name": "b"}])
    assert out["lidar"].C.dtype == torch.int32 and out["lidar"].C[:, 3].tolist() == [0, 0, 0, 1, 1]
    assert out["n"].shape == (2, 1) and out["name"] == ["a", "b"] and out["lidar"].s == (1, 1, 1)
    s = out["lidar"] + out["lidar"]
    assert s.kmaps is out["lidar"].kmaps and s.cmaps is out["lidar"].cmaps and float(s.F.sum()) == 20.0
    z = taseg_b200.PointTensor(torch.zeros(2, 2), torch.zeros(2, 4))
    assert z.additional_features == {"idx_query": {}, "counts": {}} and (z + z).idx_query is z.idx_query


def test_slice_plan_host_logic():
    """K-slice packing of the tensor-core convolution (host arithmetic only, no device call): the flat chunk stream of a
    layer with C = c0 + c1 input channels repeats every P = 8 / gcd(C/8, 8) offsets with Q = (C/8) / gcd 64-channel slices,
    so the packed weights hold ceil(K / P) * Q blocks of c_out x 64 bf16."""
    from math import gcd
    from taseg_b200 import _lib
    L = _lib.lib()
    for k, c0, c1, c_out in [(27, 96, 0, 96), (27, 96, 32, 96), (27, 32, 0, 32), (27, 16, 0, 32), (8, 64, 0, 64), (27, 256, 128, 256),
                             (1, 128, 64, 128), (27, 48, 0, 80), (9, 128, 0, 16)]:
        cpo = (c0 + c1) // 8
        g = gcd(cpo, 8)
        p, q = 8 // g, cpo // g
        want = -(-k // p) * q * c_out * 64 * 2
        assert L.tsg_conv_pack_bytes(k, c0, c1, c_out) == want, (k, c0, c1, c_out)
        # every slice is full: the blocks cover exactly ceil(K/P)*P offsets x C channels
        assert -(-k // p) * q * 64 == -(-k // p) * p * (c0 + c1)


def test_kd_model_state_dict_matches_reference():
    """MinkUNetMsKd mirror: parameter / buffer names and shapes are those of the reference model (the fixture holds the
    reference's own state_dict, tests/golden/make_golden_kd.py), so reference checkpoints load."""
    import os
    import numpy as np
    from taseg_b200.segmentor import MinkUNetMsKd, ModelCfg
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kd.npz"))
    cfg = ModelCfg(IN_FEATURE_DIM=5, BLOCK="ResBlock", NUM_LAYER=[1, 1, 1, 1, 1, 1, 1, 1], cr=0.125, pres=0.05, vres=0.05,
                   IF_DIST=False, IGNORE_LABEL=0, DROPOUT_P=0.0, SAMPLING_TYPE="random", MAX_VOXEL=10 ** 9, FEAT_KD_WEIGHT=2.0)
    mine = {k: tuple(v.shape) for k, v in MinkUNetMsKd(cfg, 20).state_dict().items()}
    ref = {k[3:]: tuple(g[k].shape) for k in g.files if k.startswith("sd/")}
    assert mine == ref
    assert sum(k.startswith("stem_gt.") for k in mine) == sum(k.startswith("stem.") for k in mine) > 0


def test_training_path_host_logic_cpu():
    """Host-side pieces of the round-2 training path that need no GPU: column blocks of the tensor-core data gradient, the
    slice-plan guard of the split-precision fp32 path, BatchNorm + ReLU fusion marks (names unchanged, CPU rows fall back to
    ATen + relu with identical results), K-split policy switches."""
    import torch
    from taseg_b200 import nn as spnn
    from taseg_b200 import ops
    from taseg_b200.nn.functional import conv as C
    from taseg_b200.nn.modules.norm import fuse_bn_relu
    assert C._col_blocks(96) == [(0, 96)] and C._col_blocks(256) == [(0, 256)]
    assert C._col_blocks(384) == [(0, 192), (192, 384)] and C._col_blocks(272) == [(0, 144), (144, 272)]
    assert all((b - a) % 16 == 0 and b - a <= 256 for a, b in C._col_blocks(1040))
    seq = fuse_bn_relu(torch.nn.Sequential(spnn.BatchNorm(8), spnn.ReLU(True), spnn.BatchNorm(8)))
    assert seq[0].fuse_relu and seq[1].fused_upstream and not seq[2].fuse_relu
    assert list(seq.state_dict()) == list(torch.nn.Sequential(torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.BatchNorm1d(8)).state_dict())
    x = torch.randn(64, 8)
    ref = torch.nn.BatchNorm1d(8)
    ref.load_state_dict(seq[0].state_dict())
    got = seq[1](spnn.modules.norm.BatchNorm._rows(seq[0], x)) if False else seq[0]._rows(x)     # CPU rows: ATen + relu
    assert torch.allclose(got, torch.relu(ref(x)), atol=1e-6)
    assert not ops.SplitItems.wanted(15307, 27) or ops.SPLIT_K          # off by default
    assert not ops.bn_supported(x)                                      # CPU tensors never reach the kernels
