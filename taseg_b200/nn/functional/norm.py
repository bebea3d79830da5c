"""BatchNorm over the (N, C) feature matrix on the C-ABI kernels (taseg_b200/csrc/bn.cu): the function spnn.BatchNorm and the
segmentor's BatchNorm call on CUDA feature rows.  Semantics of nn.BatchNorm1d (TS/torchsparse/nn/modules/norm.py:10-13 applies
it to SparseTensor.F): batch statistics + running-statistics update in training, running statistics in evaluation, fp32
affine parameters, output in the input's dtype (so it composes with autocast like ATen's batch_norm does)."""
from typing import Optional

import torch
from torch.autograd import Function

from ... import ops

__all__ = ['batch_norm']


class BatchNormFunction(Function):

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training: bool, momentum: float, eps: float, relu: bool = False,
                batches_tracked=None):
        x = x.contiguous()
        w = weight.float().contiguous() if weight is not None else None
        b = bias.float().contiguous() if bias is not None else None
        if training:
            mean, invstd = ops.bn_stats(x, eps, momentum, running_mean, running_var, batches_tracked)
        else:
            mean, invstd = running_mean.float(), torch.rsqrt(running_var.float() + eps)
        y = ops.bn_apply(x, mean, invstd, w, b, relu)
        ctx.save_for_backward(x, mean, invstd, w, b)
        ctx.training = training
        ctx.relu = relu
        ctx.has_affine = (weight is not None, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, invstd, w, b = ctx.saved_tensors
        dx, dgamma, dbeta = ops.bn_backward(x, dy, mean, invstd, w, ctx.training, want_dx=ctx.needs_input_grad[0], beta=b,
                                            relu=ctx.relu)
        return (dx, dgamma if ctx.has_affine[0] else None, dbeta if ctx.has_affine[1] else None, None, None, None, None, None, None,
                None)


def batch_norm(module: torch.nn.modules.batchnorm._BatchNorm, x: torch.Tensor, relu: bool = False) -> Optional[torch.Tensor]:
    """nn.BatchNorm1d.forward(x) (followed by ReLU when `relu`) for CUDA feature rows the kernels take; None when the caller
    should use ATen.  The fused ReLU is differentiable in training mode only (evaluation-mode backward: unfused)."""
    if not ops.bn_supported(x):
        return None
    if relu and not module.training and torch.is_grad_enabled() and x.requires_grad:
        return None
    training = module.training or (module.running_mean is None and module.running_var is None)
    momentum = 0.0 if module.momentum is None else module.momentum
    tracked = None
    if module.training and module.track_running_stats and module.num_batches_tracked is not None:
        if module.momentum is None:      # cumulative moving average: the factor depends on the count (a host read, as in torch)
            module.num_batches_tracked.add_(1)
            momentum = 1.0 / float(module.num_batches_tracked)
        else:                            # incremented by the statistics kernel (one launch less per BatchNorm and step)
            tracked = module.num_batches_tracked
            if not tracked.is_cuda or tracked.dtype != torch.int64:
                tracked.add_(1)
                tracked = None
    rm = module.running_mean if (not training or module.track_running_stats) else None
    rv = module.running_var if (not training or module.track_running_stats) else None
    if rm is not None and (rm.dtype != torch.float32 or not rm.is_contiguous()):
        return None
    return BatchNormFunction.apply(x, module.weight, module.bias, rm, rv, training, momentum, module.eps, relu, tracked)
