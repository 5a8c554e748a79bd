"""Query self-attention of the decoder on the hand-written small-attention kernels (csrc/small_attn.cu).

``mha_forward(module, query, key, value, key_padding_mask)`` evaluates an ``nn.MultiheadAttention`` module (sequence
first, ``query is key`` -- the decoder's call, models/deformable_transformer.py:366-368 in the reference) with the same
parameters (``in_proj_weight`` / ``in_proj_bias`` / ``out_proj``: the state_dict contract is untouched): the packed
in-projection runs as two library GEMMs ([q | k] from ``query``, v from ``value``), the attention core
``dropout(softmax(q k^T / sqrt(d))) v`` as one launch forward and two backward, the output projection as a GEMM.
Head width must be 32 (256 / 8 and 288 / 9); anything else, CPU tensors and attention masks other than a key-padding
mask use the stock module.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ext, seeds
from .fused_linear import linear as fused_linear

_ENABLED = os.environ.get("TFB200_SMALL_ATTN", "1") != "0"


class _SmallAttention(Function):
    @staticmethod
    def forward(ctx, q, k, v, key_pad, seed, scale, keep_prob):
        out, lse = ext.load().small_attention_forward(q, k, v, key_pad, seed, scale, keep_prob)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.key_pad, ctx.seed, ctx.scale, ctx.keep_prob = key_pad, seed, scale, keep_prob
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        q, k, v, out, lse = ctx.saved_tensors
        dq, dk, dv = ext.load().small_attention_backward(q, k, v, ctx.key_pad, ctx.seed, out, lse, grad_out, ctx.scale,
                                                         ctx.keep_prob)
        return dq, dk, dv, None, None, None, None


def supported(module: torch.nn.MultiheadAttention, query: torch.Tensor) -> bool:
    e, h = module.embed_dim, module.num_heads
    return bool(_ENABLED and query.is_cuda and query.dtype == torch.float32 and e // h == 32 and e % 4 == 0
                and module._qkv_same_embed_dim and module.in_proj_bias is not None and not module.batch_first
                and module.bias_k is None and not module.add_zero_attn)


def mha_forward(module: torch.nn.MultiheadAttention, query: torch.Tensor, value: torch.Tensor,
                key_padding_mask=None) -> torch.Tensor:
    """``module(query, query, value, key_padding_mask=..., need_weights=False)[0]`` for [L, B, E] inputs."""
    length, batch, e = query.shape
    h = module.num_heads
    w, b = module.in_proj_weight, module.in_proj_bias
    qk = fused_linear(query, w[:2 * e], b[:2 * e])                  # [L, B, 2E]: q | k
    vv = fused_linear(value, w[2 * e:], b[2 * e:])                  # [L, B, E]
    q = qk[..., :e].unflatten(-1, (h, 32))
    k = qk[..., e:].unflatten(-1, (h, 32))
    v = vv.unflatten(-1, (h, 32))
    p = float(module.dropout) if module.training else 0.0
    seed = seeds.next_seed(query.device) if p > 0.0 else None
    key_pad = None
    if key_padding_mask is not None:
        key_pad = key_padding_mask if key_padding_mask.dtype == torch.bool else key_padding_mask != 0
    out = _SmallAttention.apply(q, k, v, key_pad, seed, 32 ** -0.5, 1.0 - p)
    return fused_linear(out.flatten(2), module.out_proj.weight, module.out_proj.bias)
