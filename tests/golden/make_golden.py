#!/usr/bin/env python
"""Generate op-level golden fixtures FROM THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``

It imports the reference's own pure-PyTorch definition
``ms_deform_attn_core_pytorch``
(/root/reference/src/trackformer/models/ops/functions/ms_deform_attn_func.py:34-54)
-- with the compiled extension it imports at top level (:11) stubbed out -- and
records, for seeded inputs, its forward output and the autograd gradients of a
scalar loss w.r.t. value / sampling locations / attention weights.

Cases
  ref_test_f32   the reference's ops/test.py problem (seed 3; N=2 M=2 D=4 Lq=3 L=3 P=2;
                 shapes (8,8),(4,4),(2,2); value=rand*0.01, loc=rand, attn=rand+1e-5
                 normalised over (L,P); loss = out.abs().sum())          test.py:14-49
  ref_test_f64   ops/test_double_precision.py: shapes (12,8),(6,4),(3,2), double
  model_d32      M=8 D=32 L=4 P=4 (the shipped Deformable-DETR head geometry), non-square
                 levels, locations in [-0.15, 1.15] (exercises the zero-padding rule)
  model_d36_l8   M=8 D=36 L=8 P=4 (multi-frame TrackFormer decoder geometry)
  border(_fwd)   sample points on / next to pixel centres, borders and the validity limits
  batch_ragged   N=3, Lq=1 and tiny 1xW / Hx1 levels
Each fixture stores inputs AND outputs (npz), so the GPU box needs neither the
reference nor a bit-identical RNG.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"


def load_reference_fn():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    sys.path.insert(0, os.path.join(REF_SRC, "trackformer", "models", "ops"))
    from functions.ms_deform_attn_func import ms_deform_attn_core_pytorch  # noqa
    return ms_deform_attn_core_pytorch


def run_case(fn, value, shapes, loc, attn, grad_out=None):
    value = value.clone().requires_grad_(True)
    loc = loc.clone().requires_grad_(True)
    attn = attn.clone().requires_grad_(True)
    out = fn(value, shapes, loc, attn)
    if grad_out is None:                       # the reference test's loss: out.abs().sum()
        loss = out.abs().sum()
        grad_out = torch.sign(out.detach())
    else:
        loss = (out * grad_out).sum()
    gv, gl, ga = torch.autograd.grad(loss, (value, loc, attn))
    return dict(value=value.detach().numpy(), shapes=shapes.numpy(), loc=loc.detach().numpy(),
                attn=attn.detach().numpy(), grad_out=grad_out.numpy(), out=out.detach().numpy(),
                grad_value=gv.numpy(), grad_loc=gl.numpy(), grad_attn=ga.numpy())


def norm_attn(a):
    return a / a.sum(-1, keepdim=True).sum(-2, keepdim=True)


def main():
    fn = load_reference_fn()
    cases = {}

    # --- the reference's own test problems (same seed / construction order) ----------
    for name, dt, hw in (("ref_test_f32", torch.float32, [(8, 8), (4, 4), (2, 2)]),
                         ("ref_test_f64", torch.float64, [(12, 8), (6, 4), (3, 2)])):
        N, M, D, Lq, L, P = 2, 2, 4, 3, 3, 2
        shapes = torch.as_tensor(hw, dtype=torch.long)
        S = int((shapes[:, 0] * shapes[:, 1]).sum())
        torch.manual_seed(3)
        value = (torch.rand(N, S, M, D) * 0.01).to(dt)
        loc = torch.rand(N, Lq, M, L, P, 2).to(dt)
        attn = norm_attn(torch.rand(N, Lq, M, L, P).to(dt) + 1e-5)
        cases[name] = run_case(fn, value, shapes, loc, attn)

    # --- model geometries ----------------------------------------------------------------
    def rand_case(seed, N, M, D, Lq, hw, P, lo, hi, dt=torch.float32):
        g = torch.Generator().manual_seed(seed)
        shapes = torch.as_tensor(hw, dtype=torch.long)
        L = len(hw)
        S = int((shapes[:, 0] * shapes[:, 1]).sum())
        value = torch.randn(N, S, M, D, generator=g).to(dt)
        loc = (torch.rand(N, Lq, M, L, P, 2, generator=g) * (hi - lo) + lo).to(dt)
        attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P).to(dt)
        grad_out = torch.randn(N, Lq, M * D, generator=g).to(dt)
        return run_case(fn, value, shapes, loc, attn, grad_out)

    cases["model_d32"] = rand_case(11, 1, 8, 32, 40, [(12, 16), (6, 8), (3, 4), (2, 2)], 4, -0.15, 1.15)
    cases["model_d36_l8"] = rand_case(12, 1, 8, 36, 17,
                                      [(9, 12), (5, 6), (3, 3), (2, 2), (9, 12), (5, 6), (3, 3), (2, 2)],
                                      4, -0.05, 1.05)
    cases["batch_ragged"] = rand_case(13, 3, 2, 8, 1, [(1, 7), (5, 1), (1, 1)], 3, -0.3, 1.3)
    cases["model_d32_f64"] = rand_case(14, 2, 8, 32, 9, [(6, 9), (3, 5)], 4, -0.1, 1.1, torch.float64)

    # --- border / validity-limit placements ------------------------------------------------
    # "border_fwd_f64": exact pixel centres, exact borders and the exact validity limits
    #   (x = -1, x = W).  The bilinear surface is continuous there, so the FORWARD value is
    #   well defined, but its derivative is not (one-sided) -- tests compare `out` only.
    # "border_f64": the same neighbourhood with non-integral pixel coordinates, where the
    #   gradients are well defined -- tests compare everything.
    H, W = 4, 6
    shapes = torch.as_tensor([(H, W)], dtype=torch.long)

    def border_case(px, py, seed):
        gx, gy = torch.meshgrid(torch.tensor(px), torch.tensor(py), indexing="ij")
        lx = (gx.reshape(-1).double() + 0.5) / W          # invert x = loc*W - 0.5
        ly = (gy.reshape(-1).double() + 0.5) / H
        Lq = lx.numel()
        g = torch.Generator().manual_seed(seed)
        value = torch.randn(1, H * W, 2, 4, generator=g, dtype=torch.float64)
        loc = torch.stack([lx, ly], -1).view(1, Lq, 1, 1, 1, 2).expand(1, Lq, 2, 1, 1, 2).contiguous()
        attn = torch.rand(1, Lq, 2, 1, 1, generator=g, dtype=torch.float64) + 0.5
        grad_out = torch.randn(1, Lq, 8, generator=g, dtype=torch.float64)
        return run_case(fn, value, shapes, loc, attn, grad_out)

    cases["border_fwd_f64"] = border_case(
        [-1.5, -1.0, -0.5, 0.0, 1.0, 2.0, W - 1.0, W - 0.5, float(W), W + 0.5],
        [-1.5, -1.0, -0.5, 0.0, 1.0, 2.0, H - 1.0, H - 0.5, float(H), H + 0.5], 15)
    cases["border_f64"] = border_case(
        [-1.25, -0.999, -0.5, -0.125, 0.25, 2.375, W - 1.25, W - 0.75, W - 0.5, W - 1e-3, W + 0.25],
        [-1.25, -0.999, -0.5, -0.125, 0.25, 2.375, H - 1.25, H - 0.75, H - 0.5, H - 1e-3, H + 0.25], 16)

    for name, c in cases.items():
        path = os.path.join(HERE, f"msda_{name}.npz")
        np.savez(path, **c)
        print(f"{name:16s} out{tuple(c['out'].shape)} -> {os.path.relpath(path)} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
