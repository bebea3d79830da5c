"""Mirror of the reference's pybind11 module `torchsparse.backend` (TS/torchsparse/backend/pybind_cuda.cpp:18-39)
on top of the C ABI, for code that reaches below torchsparse.nn.functional.  Same names, argument order and
return conventions as the reference's *_cuda entry points; the *_cpu names are deliberately absent (no CPU path).
"""
import torch

from . import ops

__all__ = ['hash_cuda', 'kernel_hash_cuda', 'hash_query_cuda', 'count_cuda', 'voxelize_forward_cuda',
           'voxelize_backward_cuda', 'devoxelize_forward_cuda', 'devoxelize_backward_cuda',
           'convolution_forward_cuda', 'convolution_backward_cuda']


def hash_cuda(idx):
    return ops.sphash(idx)


def kernel_hash_cuda(idx, kernel_offset):
    return ops.sphash(idx, kernel_offset)


def hash_query_cuda(hash_query, hash_target, idx_target):
    """Returns idx_target[pos]+1 for hits and 0 for misses, like the reference (query_cuda.cu:9-56)."""
    pos = ops.sphashquery(hash_query, hash_target)
    hit = pos >= 0
    out = torch.zeros_like(pos)
    out[hit] = idx_target[pos[hit]] + 1
    return out


def count_cuda(idx, s):
    return ops.spcount(idx, s)


def voxelize_forward_cuda(inputs, idx, counts):
    return ops.voxelize_forward(inputs, idx, counts)


def voxelize_backward_cuda(top_grad, idx, counts, N):
    return ops.voxelize_backward(top_grad, idx, counts, N)


def devoxelize_forward_cuda(feat, indices, weight):
    return ops.devoxelize_forward(feat, indices, weight)


def devoxelize_backward_cuda(top_grad, indices, weight, n):
    return ops.devoxelize_backward(top_grad, indices, weight, n)


def convolution_forward_cuda(in_feat, out_feat, kernel, neighbor_map, neighbor_offset, transpose):
    """In-place on out_feat like the reference (convolution_cuda.cu:53-165); neighbor_offset is the CPU int32 nbsizes."""
    if in_feat.shape[1] != kernel.shape[1]:
        raise ValueError('Input feature size and kernel size mismatch')
    nbr = ops.kmap_from_pairs(neighbor_map, neighbor_offset, kernel.shape[0], bool(transpose), out_feat.shape[0])
    out_feat.copy_(ops.conv_forward(in_feat, kernel, nbr, out_feat.shape[0]))


def convolution_backward_cuda(in_feat, grad_in_feat, grad_out_feat, kernel, grad_kernel, neighbor_map, neighbor_offset,
                              transpose):
    k = kernel.shape[0]
    rows_out_of_in = ops.kmap_from_pairs(neighbor_map, neighbor_offset, k, not bool(transpose), in_feat.shape[0])
    rows_in_of_out = ops.kmap_from_pairs(neighbor_map, neighbor_offset, k, bool(transpose), grad_out_feat.shape[0])
    grad_in_feat.copy_(ops.conv_dgrad(grad_out_feat, kernel, rows_out_of_in, in_feat.shape[0]))
    grad_kernel.copy_(ops.conv_wgrad(in_feat, grad_out_feat, rows_in_of_out, k))
