"""Loader for oracle/_ref — the reference's OWN compiled CPU backend (torchsparse 1.4.0,
TS/backend/pybind_cpu.cpp:18-29) — and an `ops` adapter that runs the oracle graphs on it.

TEST INFRASTRUCTURE ONLY (see oracle/ts_oracle.py header).  Used (a) to pin the numpy restatement
and (b) as the `"kind": "reference"` CPU baseline that bench.py times.  The Python glue between the
native calls follows TS/nn/functional/*.py; it is restated (not imported) because the reference's
Python sources may not be copied into this repository and /root/reference is absent on the GPU box.
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np

from . import ts_oracle as T
from .build_ref import OUT, so_name

_backend = None


def available() -> bool:
    return os.path.exists(os.path.join(OUT, so_name()))


def backend():
    """The pybind11 module: hash_cpu, kernel_hash_cpu, hash_query_cpu, count_cpu, voxelize_forward_cpu,
    voxelize_backward_cpu, devoxelize_forward_cpu, devoxelize_backward_cpu, convolution_forward_cpu,
    convolution_backward_cpu."""
    global _backend
    if _backend is None:
        import torch  # noqa: F401  (the .so links libtorch)
        path = os.path.join(OUT, so_name())
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref is not built; run `python oracle/build_ref.py` in the build container")
        spec = importlib.util.spec_from_file_location("backend", path)
        _backend = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_backend)
    return _backend


def _t(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t if dtype is None else t.to(dtype)


class RefOps:
    """Same interface as net_oracle.NumpyOps, native calls executed by the reference backend.
    kernel_hash_cpu reuses row 0's batch index (hash_cpu.cpp:29), so callers must feed ONE scan
    (batch index 0) at a time — exactly how bench.py's reference arm and the tests use it."""

    @staticmethod
    def sphash(coords, offsets=None):
        import torch
        be = backend()
        c = _t(coords, torch.int32)
        if offsets is None:
            return be.hash_cpu(c).numpy()
        return be.kernel_hash_cpu(c, _t(offsets, torch.int32)).numpy()

    @staticmethod
    def sphashquery(queries, references):
        import torch
        be = backend()
        q = _t(queries, torch.int64)
        r = _t(references, torch.int64)
        out = be.hash_query_cpu(q.reshape(-1), r, torch.arange(len(r), dtype=torch.long))
        return (out - 1).reshape(q.shape).numpy()

    @staticmethod
    def spcount(idx, num):
        import torch
        return backend().count_cpu(_t(idx, torch.int32), int(num)).numpy()

    @staticmethod
    def spvoxelize(feats, idx, counts):
        import torch
        return backend().voxelize_forward_cpu(_t(feats, torch.float32), _t(idx, torch.int32),
                                              _t(counts, torch.int32)).numpy()

    @staticmethod
    def spdevoxelize(feats, idx, weights):
        import torch
        return backend().devoxelize_forward_cpu(_t(feats, torch.float32), _t(idx, torch.int32),
                                                _t(weights, torch.float32)).numpy()

    @staticmethod
    def conv_forward(in_feat, kernel, nbmaps, nbsizes, sizes, transposed=False):
        import torch
        x = _t(in_feat, torch.float32)
        w = _t(kernel, torch.float32)
        out = torch.zeros(sizes[0] if transposed else sizes[1], w.shape[-1])
        backend().convolution_forward_cpu(x, out, w, _t(nbmaps, torch.int32), _t(nbsizes, torch.int32),
                                          bool(transposed))
        return out.numpy()

    @staticmethod
    def conv_backward(in_feat, grad_out, kernel, nbmaps, nbsizes, transposed=False):
        import torch
        x = _t(in_feat, torch.float32)
        w = _t(kernel, torch.float32)
        gx, gw = torch.zeros_like(x), torch.zeros_like(w)
        backend().convolution_backward_cpu(x, gx, _t(grad_out, torch.float32), w, gw, _t(nbmaps, torch.int32),
                                           _t(nbsizes, torch.int32), bool(transposed))
        return gx.numpy(), gw.numpy()

    @staticmethod
    def matmul(a, b):
        import torch
        return (_t(a, torch.float32) @ _t(b, torch.float32)).numpy()

    # ---- glue restated from TS/nn/functional/conv.py:122-205 on top of the native calls
    @classmethod
    def conv3d(cls, x, weight, kernel_size, stride=1, dilation=1, transposed=False, bias=None):
        ks, st, dl = T.make_ntuple(kernel_size), T.make_ntuple(stride), T.make_ntuple(dilation)
        if ks == (1, 1, 1) and st == (1, 1, 1):
            return x.like(cls.matmul(x.F, weight))
        if not transposed:
            os_ = tuple(x.s[k] * st[k] for k in range(3))
            oc = x.cmaps.get(os_)
            if oc is None:
                oc = x.C if all(s == 1 for s in st) else T.spdownsample(x.C, st, ks, x.s)
            key = (x.s, ks, st, dl)
            if key not in x.kmaps:
                offsets = T.get_kernel_offsets(ks, stride=x.s, dilation=dl)
                results = cls.sphashquery(cls.sphash(oc, offsets), cls.sphash(x.C))
                nbsizes = (results != -1).sum(axis=1)
                kk, jj = np.nonzero(results != -1)
                x.kmaps[key] = [np.stack([results[kk, jj], jj], 1), nbsizes, (x.C.shape[0], oc.shape[0])]
            out = x.like(cls.conv_forward(x.F, weight, *x.kmaps[key], transposed=False), oc, os_)
        else:
            os_ = tuple(x.s[k] // st[k] for k in range(3))
            oc = x.cmaps[os_]
            out = x.like(cls.conv_forward(x.F, weight, *x.kmaps[(os_, ks, st, dl)], transposed=True), oc, os_)
        out.cmaps.setdefault(out.s, out.C)
        return out

    # ---- glue restated from R/pcseg/model/segmentor/voxel/minkunet/utils.py:11-105
    @classmethod
    def initial_voxelize(cls, z, init_res, after_res):
        f = np.float32
        nfc = np.concatenate([(z.C[:, :3].astype(f) * f(init_res)) / f(after_res), z.C[:, -1:].astype(f)], 1)
        fl = np.floor(nfc)
        pc_hash = cls.sphash(fl.astype(np.int32))
        sparse_hash = np.unique(pc_hash)
        idx_query = cls.sphashquery(pc_hash, sparse_hash)
        counts = cls.spcount(idx_query.astype(np.int32), len(sparse_hash))
        ic = np.round(cls.spvoxelize(fl, idx_query, counts)).astype(np.int32)
        x = T.SparseTensor(cls.spvoxelize(z.F, idx_query, counts), ic, 1)
        x.cmaps.setdefault(x.s, x.C)
        z.additional_features["idx_query"][1] = idx_query
        z.additional_features["counts"][1] = counts
        z.C = nfc
        return x

    @classmethod
    def point_to_voxel(cls, x, z):
        if z.additional_features["idx_query"].get(x.s) is None:
            idx_query = cls.sphashquery(cls.sphash(T._floor_to_stride(z.C, x.s[0])), cls.sphash(x.C))
            z.additional_features["idx_query"][x.s] = idx_query
            z.additional_features["counts"][x.s] = cls.spcount(idx_query.astype(np.int32), x.C.shape[0])
        return x.like(cls.spvoxelize(z.F, z.additional_features["idx_query"][x.s],
                                     z.additional_features["counts"][x.s]))

    @classmethod
    def voxel_to_point(cls, x, z, nearest=False):
        if z.idx_query.get(x.s) is None or z.weights.get(x.s) is None:
            off = T.get_kernel_offsets(2, x.s, 1)
            idx_query = cls.sphashquery(cls.sphash(T._floor_to_stride(z.C, x.s[0]), off), cls.sphash(x.C))
            weights = np.ascontiguousarray(T.calc_ti_weights(z.C, idx_query, scale=x.s[0]).T)
            idx_query = np.ascontiguousarray(idx_query.T)
            if nearest:
                weights[:, 1:] = 0.0
                idx_query[:, 1:] = -1
            z.idx_query[x.s] = idx_query
            z.weights[x.s] = weights
        new = T.PointTensor(cls.spdevoxelize(x.F, z.idx_query[x.s], z.weights[x.s]), z.C, z.idx_query, z.weights)
        new.additional_features = z.additional_features
        return new
