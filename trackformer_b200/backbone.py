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

    def affine(self):
        """Per-channel (scale, shift) with  scale = weight / sqrt(running_var + eps),  shift = bias - running_mean * scale,
        shaped [1, C, 1, 1].  All four inputs are frozen buffers, so the pair is computed once and reused until a
        buffer is written to or moved (the reference recomputes it in every forward, backbone.py:46-55: five sub-3-us
        launches x 53 layers per ResNet-50 pass).  Same values either way."""
        bufs = (self.weight, self.bias, self.running_mean, self.running_var)
        # inference tensors carry no version counter and must not leak into later autograd graphs: no caching there
        cacheable = not torch.is_inference_mode_enabled() and not any(b.is_inference() for b in bufs)
        if cacheable:
            key = tuple(b._version for b in bufs) + tuple(b.data_ptr() for b in bufs) + (self.weight.dtype,)
            cached = getattr(self, "_affine_cache", None)
            if cached is not None and cached[0] == key:
                return cached[1], cached[2]
        scale = self.weight * (self.running_var + 1e-5).rsqrt()
        shift = self.bias - self.running_mean * scale
        scale, shift = scale.view(1, -1, 1, 1), shift.view(1, -1, 1, 1)
        if cacheable:
            self._affine_cache = (key, scale, shift)
        return scale, shift

    def forward(self, x):
        scale, shift = self.affine()
        return torch.addcmul(shift, x, scale)             # one pass: x * scale + shift


_DENSE_MASKS: Dict[tuple, torch.Tensor] = {}


def _resize_mask(mask: torch.Tensor, size) -> torch.Tensor:
    """Nearest-neighbour resize of the padding mask to a feature map's size (backbone.py:80-88).  For a batch known to
    be dense (`_no_padding`) the result is all False whatever the size: one shared tensor per (batch, size, device)
    instead of three small launches per level and frame."""
    size = tuple(int(v) for v in size)
    if getattr(mask, "_no_padding", False) and not torch.is_inference_mode_enabled():
        key = (int(mask.shape[0]), size, mask.device)
        hit = _DENSE_MASKS.get(key)
        if hit is None:
            # an expanded single element: read-only by construction (an in-place write into it raises), so a caller
            # cannot corrupt the mask every other dense batch of this shape shares
            hit = torch.zeros((1, 1, 1), dtype=torch.bool, device=mask.device).expand((mask.shape[0],) + size)
            hit._no_padding = True
            _DENSE_MASKS[key] = hit
        return hit
    out = F.interpolate(mask[None].float(), size=size).to(torch.bool)[0]
    if getattr(mask, "_no_padding", False):
        out._no_padding = True
    return out


class _Trunk(nn.ModuleDict):
    """The convolutional trunk of a torchvision ResNet as an ordered dict of its stages (``conv1 .. layer4``; the
    classifier head is dropped).  ``forward`` runs the stages in order and collects the outputs listed in ``taps``
    (stage name -> output key).  Registered under the attribute name ``body`` so that parameter names stay
    ``backbone.0.body.<stage>...`` like the reference's IntermediateLayerGetter-based trunk (backbone.py:70-78)."""

    def __init__(self, resnet: nn.Module, taps: Dict[str, str]):
        last = max(i for i, (name, _) in enumerate(resnet.named_children()) if name in taps)
        super().__init__({name: mod for i, (name, mod) in enumerate(resnet.named_children()) if i <= last})
        self.taps = dict(taps)

    #: optional ``cut(tag, tensor) -> tensor`` installed by TrainStep's phased backward: called on every tapped stage
    #: output and on the tensor handed from layer3 to layer4, it may return a detached leaf so that the backward can be
    #: run (and its gradients all-reduced) in pieces.  None = plain forward.
    cut = None

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        collected = {}
        fuse_stem = getattr(self, "fused_bn", False)
        cut = self.cut
        skip_pool = False
        for name, stage in self.items():
            if fuse_stem and name == "bn1":
                from .fused_bn import bn_act, fusable
                pool = self["maxpool"] if "maxpool" in self else None
                if (pool is not None and not x.requires_grad and fusable(x) and pool.kernel_size == 3 and pool.stride == 2
                        and pool.padding == 1 and pool.dilation == 1 and not pool.ceil_mode):
                    # frozen stem (reference backbone.py:64-68): bn1 + relu + maxpool as one forward-only pass
                    from . import ext
                    scale, shift = stage.affine()
                    x = ext.load().frozen_bn_relu_maxpool(x, scale, shift)
                    skip_pool = True
                else:
                    x = bn_act(x, stage, True)         # stem: bn1 + relu in one pass
                continue
            if fuse_stem and name == "relu":
                continue
            if skip_pool and name == "maxpool":
                continue
            x = stage(x)
            key = self.taps.get(name)
            if key is not None:
                collected[key] = x if cut is None else cut(name + "/tap", x)
            if cut is not None and name == "layer3":
                x = cut("layer3/next", x)
        return collected


class BackboneBase(nn.Module):
    #: (strides, channels) of the four residual stages / of the last stage only
    _PYRAMID = ([4, 8, 16, 32], [256, 512, 1024, 2048])

    def __init__(self, backbone: nn.Module, train_backbone: bool, return_interm_layers: bool):
        super().__init__()
        trainable_stages = ("layer2", "layer3", "layer4") if train_backbone else ()
        for pname, param in backbone.named_parameters():         # stem and layer1 are always frozen (backbone.py:63-68)
            if not any(stage in pname for stage in trainable_stages):
                param.requires_grad_(False)
        strides, channels = self._PYRAMID
        if return_interm_layers:
            taps = {f"layer{i + 1}": str(i) for i in range(4)}
            self.strides, self.num_channels = list(strides), list(channels)
        else:
            taps = {"layer4": "0"}
            self.strides, self.num_channels = strides[-1:], channels[-1:]
        self.body = _Trunk(backbone, taps)
        # NHWC ("channels_last") activations and filters: cuDNN's tensor-core kernels are NHWC-native, so this removes
        # the nchw<->nhwc transposes cuDNN otherwise inserts around every convolution (~1.1 ms per C2 step), and the
        # [N, C, H, W] -> [N, H*W, C] flatten the transformer needs becomes a view.  Values are unchanged.
        # (applied on first CUDA use; CPU runs -- tests, the CPU baseline -- keep the reference's NCHW execution.)
        self.channels_last = os.environ.get("TFB200_CHANNELS_LAST", "1") != "0"
        self.fused_bn = os.environ.get("TFB200_FUSED_BN", "1") != "0"
        self._filters_nhwc = False

    def prepare(self) -> None:
        """Convert the filters to NHWC once the module lives on a CUDA device (idempotent).  Callers that build
        views of the parameters / gradients (TrainStep's flat gradient buffer) call this first."""
        if self.channels_last and not self._filters_nhwc and next(self.body.parameters()).is_cuda:
            self.body.to(memory_format=torch.channels_last)
            self._filters_nhwc = True
            if self.fused_bn:                           # NHWC activations from here on: fused FrozenBN + ReLU kernels
                from .fused_bn import patch_trunk
                self.body.fused_bn = patch_trunk(self.body) > 0

    def forward(self, tensor_list: NestedTensor) -> Dict[str, NestedTensor]:
        frames, pad = tensor_list.tensors, tensor_list.mask
        assert pad is not None
        if self.channels_last and frames.is_cuda:
            self.prepare()
            frames = frames.contiguous(memory_format=torch.channels_last)
        return {key: NestedTensor(fmap, _resize_mask(pad, fmap.shape[-2:])) for key, fmap in self.body(frames).items()}


class Backbone(BackboneBase):
    """torchvision ResNet trunk with frozen batch-norm; weights are never downloaded here."""

    def __init__(self, name: str, train_backbone: bool, return_interm_layers: bool, dilation: bool):
        factory = getattr(torchvision.models, name)
        net = factory(weights=None, norm_layer=FrozenBatchNorm2d, replace_stride_with_dilation=[False, False, dilation])
        super().__init__(net, train_backbone, return_interm_layers)
        if dilation:
            self.strides[-1] //= 2


class Joiner(nn.Sequential):
    """backbone -> (features, position encodings), one entry per returned layer."""

    def __init__(self, backbone, position_embedding):
        super().__init__(backbone, position_embedding)
        self.strides = backbone.strides
        self.num_channels = backbone.num_channels

    def forward(self, tensor_list: NestedTensor):
        mask = tensor_list.mask
        if mask is not None and (not hasattr(mask, "_no_padding")
                                 or getattr(mask, "_no_padding_version", mask._version) != mask._version):
            # one host sync per frame; lets every level reuse cached position encodings.  The answer is tied to the
            # tensor's version counter: a caller that writes padding into a reused mask buffer gets it re-evaluated.
            mask._no_padding = not bool(mask.any())
            mask._no_padding_version = mask._version
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
