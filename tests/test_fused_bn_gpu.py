"""Fused frozen-BN + residual + ReLU kernels (csrc/frozen_bn_act.cu) against the PyTorch op chain, and the patched
ResNet trunk against the unpatched one."""
import copy

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 64, 100, 167), (2, 256, 13, 21), (1, 2048, 25, 42), (3, 4, 1, 1)])
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False), (False, True)])
def test_kernels_match_torch_chain(cuda_device, shape, relu, res):
    from trackformer_b200.fused_bn import _FrozenBNAct
    dev = cuda_device
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = torch.randn(*shape, generator=g).to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True) if res else None
    scale = (torch.rand(1, shape[1], 1, 1, generator=g) + 0.5).to(dev)
    shift = torch.randn(1, shape[1], 1, 1, generator=g).to(dev)
    dy = torch.randn(*shape, generator=g).to(dev)
    y = _FrozenBNAct.apply(x, scale, shift, r, relu)
    assert y.is_contiguous(memory_format=torch.channels_last)
    ref = x * scale + shift
    if res:
        ref = ref + r
    if relu:
        ref = torch.relu(ref)
    torch.testing.assert_close(y, ref, rtol=1e-6, atol=1e-6)
    inputs = [x] + ([r] if res else [])
    got = torch.autograd.grad(y, inputs, dy)
    want = torch.autograd.grad(ref, inputs, dy)
    # elements within rounding of the ReLU kink may fall on either side: compare where the reference output is clear
    clear = (ref.detach().abs() > 1e-5) if relu else torch.ones_like(ref, dtype=torch.bool)
    for a, b in zip(got, want):
        torch.testing.assert_close(a[clear], b[clear], rtol=1e-6, atol=1e-6)
    # only the residual needs a gradient (identity path of a frozen stage)
    if res:
        x2 = x.detach()
        y2 = _FrozenBNAct.apply(x2, scale, shift, r, relu)
        (gr,) = torch.autograd.grad(y2, [r], dy)
        torch.testing.assert_close(gr[clear], want[1][clear], rtol=1e-6, atol=1e-6)


def test_patched_trunk_equals_unpatched_on_gpu(cuda_device, monkeypatch):
    from trackformer_b200.backbone import Backbone
    dev = cuda_device
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    fused = Backbone("resnet50", train_backbone=True, return_interm_layers=True, dilation=False)
    with torch.no_grad():
        for m in fused.modules():
            if hasattr(m, "running_var"):
                m.running_var.uniform_(0.5, 1.5)
                m.running_mean.normal_(0, 0.1)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.1)
    plain = copy.deepcopy(fused)
    plain.fused_bn = False
    fused.to(dev), plain.to(dev)
    x = torch.randn(1, 3, 160, 224, device=dev).contiguous(memory_format=torch.channels_last)
    fused.prepare(), plain.prepare()
    assert fused.body.fused_bn and not getattr(plain.body, "fused_bn", False)
    a, b = fused.body(x), plain.body(x)
    for k in a:
        torch.testing.assert_close(a[k], b[k], rtol=1e-4, atol=1e-4 * float(b[k].abs().max()))
    pa = [p for p in fused.parameters() if p.requires_grad]
    pb = [p for p in plain.parameters() if p.requires_grad]
    ga = torch.autograd.grad(sum(v.square().mean() for v in a.values()), pa)
    gb = torch.autograd.grad(sum(v.square().mean() for v in b.values()), pb)
    for u, v in zip(ga, gb):
        torch.testing.assert_close(u, v, rtol=1e-3, atol=1e-3 * float(v.abs().max()))


@pytest.mark.parametrize("shape", [(1, 64, 45, 67), (2, 64, 80, 112), (1, 8, 1, 1), (1, 64, 400, 667)])
def test_stem_bn_relu_maxpool_kernel(cuda_device, shape):
    """bn1 + relu + maxpool(3, 2, 1) of the frozen stem in one pass against the PyTorch chain (odd sizes: the last
    window row / column hangs over the border; negative scales: max must be taken AFTER the affine map)."""
    from trackformer_b200 import ext
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).to(cuda_device).contiguous(memory_format=torch.channels_last)
    scale = (torch.randn(shape[1], generator=g) * 1.5).to(cuda_device)
    shift = torch.randn(shape[1], generator=g).to(cuda_device)
    y = ext.load().frozen_bn_relu_maxpool(x, scale, shift)
    ref = F.max_pool2d(F.relu(x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)), 3, 2, 1)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(y, ref, rtol=1e-6, atol=1e-6)
