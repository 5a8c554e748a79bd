"""Frozen batch-norm + residual + ReLU of the ResNet trunk as one kernel per convolution output.

torchvision's ``Bottleneck.forward`` (the block the reference backbone is made of, backbone.py:98-100) runs
``relu(bn(conv(x)))`` twice and ``relu(bn3(conv3(.)) + identity)`` once; with the frozen batch-norm folded to a
per-channel (scale, shift) each of those is ``act(x * scale + shift [+ identity])``: one pass over the activation
instead of two or three (csrc/frozen_bn_act.cu), and one pass in the backward.  Values: fused multiply-add instead of
separate multiply and add (<= 1 ulp apart), otherwise identical.  Applies to channels-last fp32 CUDA activations with
C % 4 == 0 (every ResNet stage); anything else takes the module chain.
"""
from __future__ import annotations

import types

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ext


class _FrozenBNAct(Function):
    @staticmethod
    def forward(ctx, x, scale, shift, residual, relu):
        y = ext.load().frozen_bn_act_forward(x, residual, scale, shift, relu)
        ctx.relu = relu
        ctx.has_res = residual is not None
        ctx.save_for_backward(y if relu else None, scale)      # the ReLU mask is read off the output
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        y, scale = ctx.saved_tensors
        need_dx = ctx.needs_input_grad[0]
        need_dres = ctx.has_res and ctx.needs_input_grad[3]
        if not (need_dx or need_dres):
            return None, None, None, None, None
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx, dres = ext.load().frozen_bn_act_backward(dy, y if y is not None else dy, scale, ctx.relu, need_dx, need_dres)
        return dx, None, None, dres, None


def fusable(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0
            and x.is_contiguous(memory_format=torch.channels_last))


def bn_act(x: torch.Tensor, bn, relu: bool, residual: torch.Tensor = None) -> torch.Tensor:
    """``act(bn(x) [+ residual])`` for a FrozenBatchNorm2d ``bn``."""
    if fusable(x) and (residual is None or (fusable(residual) and residual.shape == x.shape)):
        scale, shift = bn.affine()
        return _FrozenBNAct.apply(x, scale, shift, residual, relu)
    out = bn(x)
    if residual is not None:
        out = out + residual
    return F.relu(out) if relu else out


def _bottleneck_forward(self, x):
    out = bn_act(self.conv1(x), self.bn1, True)
    out = bn_act(self.conv2(out), self.bn2, True)
    identity = x if self.downsample is None else bn_act(self.downsample[0](x), self.downsample[1], False)
    return bn_act(self.conv3(out), self.bn3, True, identity)


def patch_trunk(trunk: torch.nn.Module) -> int:
    """Give every torchvision Bottleneck of ``trunk`` the fused forward; returns how many were patched."""
    from torchvision.models.resnet import Bottleneck
    from .backbone import FrozenBatchNorm2d
    n = 0
    for m in trunk.modules():
        if isinstance(m, Bottleneck) and isinstance(m.bn1, FrozenBatchNorm2d) and \
                (m.downsample is None or (len(m.downsample) == 2 and isinstance(m.downsample[1], FrozenBatchNorm2d))):
            m.forward = types.MethodType(_bottleneck_forward, m)
            n += 1
    return n
