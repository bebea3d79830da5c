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


# fp32 tensors (train.py without --amp, fp32 evaluation through the module path) on the TENSOR cores: every operand is split
# into two bf16 terms, x = xh + xl and w = wh + wl (16 mantissa bits together), and
#   y = xh wh + xl wh + xh wl        (fp32 accumulation in TMEM; the dropped xl wl term is ~2^-16 relative)
# is two launches of the bf16 kernel with fp32 output: (xh | xl) against (wh ; wh) — the kernel's two-tensor K split — and
# xh against wl.  Error ~2e-5 relative, inside the fp32 parity bar (1e-3) and 10x faster than the CUDA-core FMA kernel at
# benchmark sizes.  TSG_FP32_SPLIT=0 restores the exact fp32 kernels.
import os as _os

FP32_SPLIT = _os.environ.get("TSG_FP32_SPLIT", "1") != "0"


def _split_bf16(t: torch.Tensor):
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


def _fp32_split_ok(feats: torch.Tensor, c_out: int) -> bool:
    """fp32 rows the split path takes: the two-tensor K split (xh | xl) must fit the kernel's slice plan (<= 16 slices per
    group of offsets: every channel count of the TASeg networks does; e.g. 272 channels do not)."""
    if not (FP32_SPLIT and feats.is_cuda and feats.dtype == torch.float32 and c_out % 16 == 0 and feats.shape[0] > 0):
        return False
    import math
    cpo = 2 * ((feats.shape[1] + 15) // 16 * 16) // 8
    return cpo // math.gcd(cpo, 8) <= 16


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
    return ops.pack_weights(w[:, :, cols[0]:cols[1]], (c_in + 15) // 16 * 16)      # a strided view: packed without a copy, rows past c_in are zero


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


def _pack_split(w: torch.Tensor, cols):
    """(K, c_in, c_out) fp32 -> packed images of (wh ; wh) for the (xh | xl) launch and of wl, columns cols, c_in padded to 16."""
    k, c_in, c_out = w.shape
    w = w[:, :, cols[0]:cols[1]]
    if c_in % 16:
        w = torch.nn.functional.pad(w, (0, 0, 0, 16 - c_in % 16))
    wh, wl = _split_bf16(w)
    c = w.shape[1]
    return (ops.pack_weights(torch.cat([wh, wh], dim=1).float().contiguous(), c, c), ops.pack_weights(wl.float().contiguous(), c))


def _conv_tc_fp32(feats: torch.Tensor, param, weight: torch.Tensor, transposed_w: bool, kmap: ops.KernelMap, map_transposed: bool,
                  n_out: int) -> torch.Tensor:
    """fp32 in, fp32 out through three bf16 products on the tensor cores (see FP32_SPLIT above)."""
    k = weight.shape[0]
    c_in, c_out = (weight.shape[2], weight.shape[1]) if transposed_w else (weight.shape[1], weight.shape[2])
    if c_in % 16:
        feats = torch.nn.functional.pad(feats, (0, 16 - c_in % 16))
    xh, xl = _split_bf16(feats.contiguous())
    nbr_s, mask_s, perm = kmap.sorted(map_transposed)
    outs = []
    for cols in _col_blocks(c_out):
        key = (id(param), transposed_w, cols, "split") if param is not None else None
        hit = _PACKS.get(key) if key is not None else None
        if hit is not None and hit[0] == param._version and hit[1] == param.data_ptr():
            p_hh, p_l = hit[2]
        else:
            w = (param if param is not None else weight).detach().float()
            p_hh, p_l = _pack_split(w.transpose(1, 2) if transposed_w else w, cols)
            if key is not None:
                _PACKS[key] = (param._version, param.data_ptr(), (p_hh, p_l))
        n_c = cols[1] - cols[0]
        y = ops.conv_forward_tc(xh, xl, p_hh, k, n_c, nbr_s, mask_s, n_out, perm=perm, out_dtype=torch.float32)
        y += ops.conv_forward_tc(xh, None, p_l, k, n_c, nbr_s, mask_s, n_out, perm=perm, out_dtype=torch.float32)
        outs.append(y)
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


def _wgrad_tc_fp32(x: torch.Tensor, gy: torch.Tensor, nbr: torch.Tensor, k: int) -> torch.Tensor:
    """fp32 weight gradient as three bf16 pair-list GEMMs: xh gh + xl gh + xh gl."""
    c_in = x.shape[1]
    if c_in % 8:
        x = torch.nn.functional.pad(x, (0, 8 - c_in % 8))
    xh, xl = _split_bf16(x.contiguous())
    gh, gl = _split_bf16(gy.contiguous())
    gw = ops.conv_wgrad_tc(xh, gh, nbr, k)
    gw += ops.conv_wgrad_tc(xl, gh, nbr, k)
    gw += ops.conv_wgrad_tc(xh, gl, nbr, k)
    return gw[:, :c_in]


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
        elif _fp32_split_ok(input, c_out):
            out = _conv_tc_fp32(input, param, weight, False, kmap, transposed, n_out)
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
        elif ctx.needs_input_grad[0] and _fp32_split_ok(grad_output, c_in):
            grad_input = _conv_tc_fp32(grad_output, param, weight, True, kmap, not transposed, input.shape[0])
        elif ctx.needs_input_grad[0]:
            grad_input = ops.conv_dgrad(grad_output, weight, tab_out_of_in, input.shape[0]).to(input.dtype)
        else:
            grad_input = None
        if input.dtype == torch.bfloat16 and c_out % 8 == 0 and c_out <= 256 and ops.WGRAD_TC:
            x = input if c_in % 8 == 0 else torch.nn.functional.pad(input, (0, 8 - c_in % 8))     # tcgen05; 5-channel stem padded
            grad_weight = ops.conv_wgrad_tc(x, grad_output, tab_in_of_out, k)[:, :c_in].to(weight.dtype)
        elif input.dtype == torch.bfloat16 and c_in % 8 == 0 and c_out % 8 == 0:
            grad_weight = ops.conv_wgrad_bf16(input, grad_output, tab_in_of_out, k).to(weight.dtype)
        elif (FP32_SPLIT and ops.WGRAD_TC and input.is_cuda and input.dtype == torch.float32 and grad_output.dtype == torch.float32
              and c_out % 8 == 0 and c_out <= 256 and input.shape[0] > 0):
            grad_weight = _wgrad_tc_fp32(input, grad_output, tab_in_of_out, k).to(weight.dtype)
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
        feats = feats.to(dt)
        # A leaf fp32 parameter is NOT cast: the tensor-core paths read its packed bf16 images (cached per parameter
        # version) and the weight gradient comes out of the kernels in fp32, so the cast, the fp32 -> bf16 rounding of the
        # gradient and autograd's cast back (three launches per convolution and step) buy nothing.
        if param is None:
            weight = weight.to(dt)

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
