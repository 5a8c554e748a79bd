"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by ``trackformer_b200``).

Pure-PyTorch (``F.grid_sample``) restatement of the reference's CPU/debug path
``ms_deform_attn_core_pytorch``
(``src/trackformer/models/ops/functions/ms_deform_attn_func.py:34-54``): per
level, view the value slab as ``[N*M, D, H, W]``, bilinear/zeros/
``align_corners=False`` sample it at ``2*loc-1``, gather all ``L*P`` samples,
weight by the attention weights and reduce.  It keeps the reference's cost
structure (one grid_sample per level, one stacked ``[N*M, D, Lq, L*P]``
intermediate, one multiply+sum) so that timing it on the host cores is a fair
"reference pure-PyTorch CPU path" baseline; autograd through it is the gradient
oracle for the backward kernels.

Third-party arithmetic: ``torch.nn.functional.grid_sample`` (ATen
grid_sampler_2d), torch 2.11.0 in this image -- the reference pins "PyTorch
1.5" in prose only (docs/INSTALL.md:12).

Pinned by tests/test_oracle.py against fixtures produced by the reference
function itself (tests/golden/make_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def msda_core_torch(value: torch.Tensor, spatial_shapes, sampling_locations: torch.Tensor,
                    attention_weights: torch.Tensor) -> torch.Tensor:
    n, _, heads, ch = value.shape
    _, lq, _, levels, points, _ = sampling_locations.shape
    hw = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes)
                                        else spatial_shapes)]
    grid = sampling_locations * 2 - 1                      # grid_sample's [-1, 1] convention
    per_level = []
    first = 0
    for lvl, (h, w) in enumerate(hw):
        slab = value[:, first:first + h * w]               # [N, H*W, M, D]
        first += h * w
        img = slab.reshape(n, h * w, heads * ch).transpose(1, 2).reshape(n * heads, ch, h, w)
        g = grid[:, :, :, lvl].transpose(1, 2).reshape(n * heads, lq, points, 2)
        per_level.append(F.grid_sample(img, g, mode="bilinear", padding_mode="zeros",
                                       align_corners=False))   # [N*M, D, Lq, P]
    sampled = torch.stack(per_level, dim=-2).flatten(-2)   # [N*M, D, Lq, L*P]
    w_ = attention_weights.transpose(1, 2).reshape(n * heads, 1, lq, levels * points)
    out = (sampled * w_).sum(-1)                           # [N*M, D, Lq]
    return out.view(n, heads * ch, lq).transpose(1, 2).contiguous()
