"""ResNet backbone + frozen batch-norm + position encoding joiner.

Mirror of src/trackformer/models/backbone.py:19-134.  The convolutions are dense contractions and stay on
cuDNN's tensor-core path (torchvision ``resnet50`` body, ``IntermediateLayerGetter`` for layer1..4); the
frozen batch-norm is folded into one per-channel scale/shift.  ``state_dict`` keys are identical to the
reference's (``backbone.0.body.layerX...``, FrozenBN buffers ``weight/bias/running_mean/running_var``) so its
checkpoints load unchanged.  Weights are never downloaded here (the reference passes
``pretrained=is_main_process()``, backbone.py:100 -- there is no network on the box).
"""
from __future__ import annotations

import os
from typing import Dict, List

import torch
import torch.nn.functional as F
import torchvision
from torch import nn
from torchvision.models._utils import IntermediateLayerGetter

from .position_encoding import build_position_encoding
from .util import NestedTensor


class FrozenBatchNorm2d(nn.Module):
    """BatchNorm2d with fixed statistics and affine terms (eps = 1e-5 inside the rsqrt)."""

    def __init__(self, n: int):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)

    def forward(self, x):
        scale = self.weight * (self.running_var + 1e-5).rsqrt()
        shift = self.bias - self.running_mean * scale
        return torch.addcmul(shift.view(1, -1, 1, 1), x, scale.view(1, -1, 1, 1))     # one pass: x * scale + shift


def _resize_mask(mask: torch.Tensor, size) -> torch.Tensor:
    out = F.interpolate(mask[None].float(), size=size).to(torch.bool)[0]
    if getattr(mask, "_no_padding", False):
        out._no_padding = True
    return out


class BackboneBase(nn.Module):
    def __init__(self, backbone: nn.Module, train_backbone: bool, return_interm_layers: bool):
        super().__init__()
        for name, p in backbone.named_parameters():
            if not train_backbone or not any(k in name for k in ("layer2", "layer3", "layer4")):
                p.requires_grad_(False)
        if return_interm_layers:
            layers = {"layer1": "0", "layer2": "1", "layer3": "2", "layer4": "3"}
            self.strides = [4, 8, 16, 32]
            self.num_channels = [256, 512, 1024, 2048]
        else:
            layers = {"layer4": "0"}
            self.strides = [32]
            self.num_channels = [2048]
        self.body = IntermediateLayerGetter(backbone, return_layers=layers)
        # NHWC ("channels_last") activations and filters: cuDNN's tensor-core kernels are NHWC-native, so this removes
        # the nchw<->nhwc transposes cuDNN otherwise inserts around every convolution (~1.1 ms per C2 step), and the
        # [N, C, H, W] -> [N, H*W, C] flatten the transformer needs becomes a view.  Values are unchanged.
        # (applied on first CUDA use; CPU runs -- tests, the CPU baseline -- keep the reference's NCHW execution.)
        self.channels_last = os.environ.get("TFB200_CHANNELS_LAST", "1") != "0"
        self._filters_nhwc = False

    def prepare(self) -> None:
        """Convert the filters to NHWC once the module lives on a CUDA device (idempotent).  Callers that build
        views of the parameters / gradients (TrainStep's flat gradient buffer) call this first."""
        if self.channels_last and not self._filters_nhwc and next(self.body.parameters()).is_cuda:
            self.body.to(memory_format=torch.channels_last)
            self._filters_nhwc = True

    def forward(self, tensor_list: NestedTensor) -> Dict[str, NestedTensor]:
        x = tensor_list.tensors
        if self.channels_last and x.is_cuda:
            self.prepare()
            x = x.contiguous(memory_format=torch.channels_last)
        feats = self.body(x)
        mask = tensor_list.mask
        assert mask is not None
        return {name: NestedTensor(x, _resize_mask(mask, x.shape[-2:])) for name, x in feats.items()}


class Backbone(BackboneBase):
    def __init__(self, name: str, train_backbone: bool, return_interm_layers: bool, dilation: bool):
        net = getattr(torchvision.models, name)(
            replace_stride_with_dilation=[False, False, dilation], weights=None, norm_layer=FrozenBatchNorm2d)
        super().__init__(net, train_backbone, return_interm_layers)
        if dilation:
            self.strides[-1] = self.strides[-1] // 2


class Joiner(nn.Sequential):
    """backbone -> (features, position encodings), one entry per returned layer."""

    def __init__(self, backbone, position_embedding):
        super().__init__(backbone, position_embedding)
        self.strides = backbone.strides
        self.num_channels = backbone.num_channels

    def forward(self, tensor_list: NestedTensor):
        if tensor_list.mask is not None and not hasattr(tensor_list.mask, "_no_padding"):
            # one host sync per frame; lets every level reuse cached position encodings
            tensor_list.mask._no_padding = not bool(tensor_list.mask.any())
        feats = self[0](tensor_list)
        out: List[NestedTensor] = []
        pos = []
        for x in feats.values():
            out.append(x)
            pos.append(self[1](x).to(x.tensors.dtype))
        return out, pos


def build_backbone(args):
    return Joiner(Backbone(args.backbone, args.lr_backbone > 0,
                           args.masks or (args.num_feature_levels > 1), args.dilation),
                  build_position_encoding(args))
