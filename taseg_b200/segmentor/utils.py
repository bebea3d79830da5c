"""Point <-> voxel helpers of the backbones, fused (R/pcseg/model/segmentor/voxel/minkunet/utils.py:11-105).

Same inputs, outputs and cache slots as the reference helpers (which also run unmodified on the drop-in F.* ops);
here each helper is one or two device passes over an exact-coordinate table instead of hash / hash / table-build /
query / ~30 elementwise launches:
  initial_voxelize : rescale+floor -> unique by ascending FNV hash (the reference's torch.unique(pc_hash) order)
                     -> counts -> scatter-mean
  point_to_voxel   : floor-to-stride + table probe -> counts -> scatter-mean
  voxel_to_point   : 8-corner probe + trilinear weights in one kernel -> weighted gather
idx_query tensors are kept as int32 (the reference holds int64 and casts to int32 at every use).
"""
import torch

from .. import ops
from ..nn import functional as F
from ..nn.functional.conv import coord_table
from ..tensor import PointTensor, SparseTensor

__all__ = ['initial_voxelize', 'point_to_voxel', 'voxel_to_point']


def initial_voxelize(z: PointTensor, init_res: float, after_res: float) -> SparseTensor:
    new_float_coord, floor_coord = ops.rescale_coords(z.C, init_res, after_res)
    voxels, _, idx_query = ops.unique_coords(floor_coord, want_index=True, want_inverse=True, by_hash=True)
    counts = ops.spcount(idx_query, voxels.shape[0])
    feats = F.spvoxelize(z.F, idx_query, counts)
    x = SparseTensor(feats, voxels.contiguous(), 1)
    x.cmaps.setdefault(x.stride, x.coords)
    z.additional_features['idx_query'][1] = idx_query
    z.additional_features['counts'][1] = counts
    z.C = new_float_coord
    return x


def point_to_voxel(x: SparseTensor, z: PointTensor) -> SparseTensor:
    cache_i, cache_c = z.additional_features['idx_query'], z.additional_features['counts']
    if cache_i.get(x.s) is None:
        idx_query = ops.point_query(coord_table(x, x.C, x.s), z.C, x.s[0])
        cache_i[x.s] = idx_query
        cache_c[x.s] = ops.spcount(idx_query, x.C.shape[0])
    return x.derive(F.spvoxelize(z.F, cache_i[x.s], cache_c[x.s]))


def voxel_to_point(x: SparseTensor, z: PointTensor, nearest: bool = False) -> PointTensor:
    if z.idx_query.get(x.s) is None or z.weights.get(x.s) is None:
        idx8, w8 = ops.trilinear_query(coord_table(x, x.C, x.s), z.C, x.s[0], nearest)
        z.idx_query[x.s], z.weights[x.s] = idx8, w8
    out = PointTensor(F.spdevoxelize(x.F, z.idx_query[x.s], z.weights[x.s]), z.C, idx_query=z.idx_query, weights=z.weights)
    out.additional_features = z.additional_features
    return out
