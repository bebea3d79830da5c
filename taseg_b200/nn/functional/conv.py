"""conv3d (TS/torchsparse/nn/functional/conv.py:122-205) on the output-stationary implicit-GEMM kernels.

Differences from the reference that do not change results:
  * the kernel map is ONE fused device pass (hash of shifted coordinates, table probe, nonzero, sum) over an
    exact-coordinate table that is built once per tensor stride and cached next to the kernel maps;
  * the convolution is ONE launch (no per-offset gather/GEMM/scatter, no nbsizes.cpu() sync);
  * `kmaps[key]` holds a KernelMap that still unpacks as `[nbmaps, nbsizes, (n_in, n_out)]`.
"""
from typing import List, Optional, Tuple, Union

import torch
from torch.autograd import Function

from ... import ops
from ...tensor import SparseTensor
from ...utils import make_ntuple
from ..utils.kernel import kernel_offsets_np
from .downsample import spdownsample

__all__ = ['conv3d']


def coord_table(x: SparseTensor, coords: torch.Tensor, stride) -> ops.Table:
    key = ('coord_table', tuple(stride))
    tab = x.kmaps.get(key)
    if tab is None or tab.n != coords.shape[0]:
        tab = ops.Table.from_coords(coords)
        x.kmaps[key] = tab
    return tab


def as_kernel_map(entry, k: int) -> ops.KernelMap:
    if isinstance(entry, ops.KernelMap):
        return entry
    nbmaps, nbsizes, (n_in, n_out) = entry     # reference-format list supplied by user code
    nbr = ops.kmap_from_pairs(nbmaps, nbsizes, k, False, n_out)
    return ops.KernelMap(nbr, nbsizes.to(torch.int32).to(nbr.device), None, n_in, n_out, k)


def _tc_ok(feats: torch.Tensor, c_in: int, c_out: int) -> bool:
    """The tensor-core kernel takes it, possibly after zero-padding the input channels to a multiple of 16 (the 5-channel
    stem) and / or in column blocks of <= 256 output channels (the data gradient of the 384-channel decoder concat)."""
    return feats.dtype == torch.bfloat16 and c_out % 16 == 0


def _col_blocks(c_out: int):
    """Output-channel blocks of at most 256 (multiples of 16) the tensor-core kernel is launched over."""
    nb = (c_out + 255) // 256
    step = (c_out // 16 + nb - 1) // nb * 16
    return [(a, min(c_out, a + step)) for a in range(0, c_out, step)]


# Packed tensor-core images of a layer's weights, W[k] for the forward and W[k]^T for the data gradient, cached per
# PARAMETER VERSION: the optimizer's in-place update bumps `_version`, so a step packs each layer once instead of on every
# forward and backward call (and an evaluation loop never re-packs).
_PACKS = {}


def _pack(w: torch.Tensor, cols):
    """(K, c_in, c_out) -> packed image of W[:, :, cols[0]:cols[1]] with c_in zero-padded to a multiple of 16."""
    k, c_in, c_out = w.shape
    w = w[:, :, cols[0]:cols[1]]
    if c_in % 16:
        w = torch.nn.functional.pad(w, (0, 0, 0, 16 - c_in % 16))
    return ops.pack_weights(w.contiguous(), w.shape[1])


def packed_weights(param: Optional[torch.Tensor], weight: torch.Tensor, transposed_w: bool, cols) -> torch.Tensor:
    """param: the fp32 parameter `weight` was cast from (the cache key), or None (no caching)."""
    if param is None:
        w = weight.detach().float()
        return _pack(w.transpose(1, 2) if transposed_w else w, cols)
    key = (id(param), transposed_w, cols)
    hit = _PACKS.get(key)
    if hit is not None and hit[0] == param._version and hit[1] == param.data_ptr():
        return hit[2]
    w = param.detach().float()
    packed = _pack(w.transpose(1, 2) if transposed_w else w, cols)
    _PACKS[key] = (param._version, param.data_ptr(), packed)
    return packed


def _conv_tc(feats: torch.Tensor, param, weight: torch.Tensor, transposed_w: bool, kmap: ops.KernelMap, map_transposed: bool,
             n_out: int) -> torch.Tensor:
    """out = conv(feats, W or W^T) on the tensor cores over the (mask-sorted) map; bf16 in, bf16 out."""
    k = weight.shape[0]
    c_in, c_out = (weight.shape[2], weight.shape[1]) if transposed_w else (weight.shape[1], weight.shape[2])
    if c_in % 16:
        feats = torch.nn.functional.pad(feats, (0, 16 - c_in % 16))
    feats = feats.contiguous()
    nbr_s, mask_s, perm = kmap.sorted(map_transposed)      # mask-sorted tile rows: empty (tile, offset) pairs are skipped
    outs = [ops.conv_forward_tc(feats, None, packed_weights(param, weight, transposed_w, cols), k, cols[1] - cols[0], nbr_s,
                                mask_s, n_out, perm=perm) for cols in _col_blocks(c_out)]
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


class ConvolutionFunction(Function):

    @staticmethod
    def forward(ctx, input: torch.Tensor, weight: torch.Tensor, kmap: ops.KernelMap, transposed: bool = False, param=None):
        """param: the fp32 parameter `weight` was cast from (autocast), the key of the packed-weight cache."""
        input = input.contiguous()
        weight = weight.contiguous()
        if input.shape[1] != weight.shape[1]:
            raise ValueError('Input feature size and kernel size mismatch')
        nbr = kmap.nbr_t if transposed else kmap.nbr
        n_out = kmap.n_in if transposed else kmap.n_out
        k, c_in, c_out = weight.shape
        if _tc_ok(input, c_in, c_out):
            out = _conv_tc(input, param, weight, False, kmap, transposed, n_out)
        else:
            out = ops.conv_forward(input, weight, nbr, n_out)
        ctx.for_backwards = (input, weight, kmap, transposed, param)
        return out

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        input, weight, kmap, transposed, param = ctx.for_backwards
        k = weight.shape[0]
        # pairs (in i, out o, k): dX[i] += dY[o] W[k]^T ; dW[k] += X[i]^T dY[o]
        tab_in_of_out = kmap.nbr_t if transposed else kmap.nbr       # rows = outputs of the forward, values = inputs
        tab_out_of_in = kmap.nbr if transposed else kmap.nbr_t       # rows = inputs of the forward, values = outputs
        c_in, c_out = weight.shape[1], weight.shape[2]
        if ctx.needs_input_grad[0] and _tc_ok(input, c_out, c_in):
            # dgrad = the same tensor-core kernel over the transposed table with W[k]^T (bf16 operands, fp32 accumulate)
            grad_input = _conv_tc(grad_output.contiguous().to(torch.bfloat16), param, weight, True, kmap, not transposed,
                                  input.shape[0])
        elif ctx.needs_input_grad[0]:
            grad_input = ops.conv_dgrad(grad_output, weight, tab_out_of_in, input.shape[0]).to(input.dtype)
        else:
            grad_input = None
        if input.dtype == torch.bfloat16 and c_out % 8 == 0 and c_out <= 256 and ops.WGRAD_TC:
            x = input if c_in % 8 == 0 else torch.nn.functional.pad(input, (0, 8 - c_in % 8))     # tcgen05; 5-channel stem padded
            grad_weight = ops.conv_wgrad_tc(x, grad_output, tab_in_of_out, k)[:, :c_in].to(weight.dtype)
        elif input.dtype == torch.bfloat16 and c_in % 8 == 0 and c_out % 8 == 0:
            grad_weight = ops.conv_wgrad_bf16(input, grad_output, tab_in_of_out, k).to(weight.dtype)
        else:
            grad_weight = ops.conv_wgrad(input, grad_output, tab_in_of_out, k).to(weight.dtype)
        return grad_input, grad_weight, None, None, None


def conv3d(input: SparseTensor, weight: torch.Tensor, kernel_size: Union[int, List[int], Tuple[int, ...]],
           bias: Optional[torch.Tensor] = None, stride: Union[int, List[int], Tuple[int, ...]] = 1,
           dilation: Union[int, Tuple[int, ...]] = 1, transposed: bool = False) -> SparseTensor:
    kernel_size, stride, dilation = make_ntuple(kernel_size, 3), make_ntuple(stride, 3), make_ntuple(dilation, 3)
    feats = input.feats
    param = weight if (weight.dim() == 3 and weight.is_leaf) else None
    if torch.is_autocast_enabled():        # reference: custom_fwd(cast_inputs=torch.half) (conv.py:19)
        dt = torch.get_autocast_dtype('cuda')
        feats, weight = feats.to(dt), weight.to(dt)

    if kernel_size == (1, 1, 1) and stride == (1, 1, 1) and dilation == (1, 1, 1):
        out_stride, out_coords = input.stride, input.coords
        out_feats = feats.matmul(weight)
    elif not transposed:
        out_stride = tuple(input.stride[k] * stride[k] for k in range(3))
        if out_stride in input.cmaps:
            out_coords = input.cmaps[out_stride]
        elif all(s == 1 for s in stride):
            out_coords = input.coords
        else:
            out_coords = spdownsample(input.coords, stride, kernel_size, input.stride)
        key = (input.stride, kernel_size, stride, dilation)
        if key not in input.kmaps:
            offsets = kernel_offsets_np(kernel_size, stride=input.stride, dilation=dilation)
            table = coord_table(input, input.coords, input.stride)
            input.kmaps[key] = ops.build_kmap(table, input.coords.shape[0], out_coords, offsets)
        kmap = as_kernel_map(input.kmaps[key], weight.shape[0])
        out_feats = ConvolutionFunction.apply(feats, weight, kmap, False, param)
    else:
        out_stride = tuple(input.stride[k] // stride[k] for k in range(3))
        out_coords = input.cmaps[out_stride]                                   # KeyError like the reference (conv.py:186)
        kmap = as_kernel_map(input.kmaps[(out_stride, kernel_size, stride, dilation)], weight.shape[0])
        out_feats = ConvolutionFunction.apply(feats, weight, kmap, True, param)

    if bias is not None:
        out_feats = out_feats + bias.to(out_feats.dtype)

    output = input.derive(out_feats, out_coords, out_stride)
    output.cmaps.setdefault(out_stride, out_coords)
    return output
