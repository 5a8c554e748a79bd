"""The re-written Bottleneck forward (fused_bn.py) must equal torchvision's block on the module-chain route (CPU);
the kernels behind the fused route are checked on the GPU (test_fused_bn_gpu.py)."""
import torch


def test_patched_trunk_equals_torchvision_blocks_on_cpu():
    import copy
    from trackformer_b200.backbone import Backbone
    from trackformer_b200.fused_bn import patch_trunk
    torch.manual_seed(0)
    bb = Backbone("resnet50", train_backbone=True, return_interm_layers=True, dilation=False)
    with torch.no_grad():
        for m in bb.modules():
            if hasattr(m, "running_var"):
                m.running_var.uniform_(0.5, 1.5)
                m.running_mean.normal_(0, 0.1)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.1)
    ref = copy.deepcopy(bb)
    assert patch_trunk(bb.body) == 16                       # ResNet-50: 3 + 4 + 6 + 3 bottlenecks
    bb.body.fused_bn = True
    x = torch.randn(1, 3, 64, 96)
    a, b = bb.body(x), ref.body(x)
    assert list(a) == list(b)
    for k in a:
        torch.testing.assert_close(a[k], b[k], rtol=1e-5, atol=1e-5)
    ga = torch.autograd.grad(sum(v.sum() for v in a.values()), [p for p in bb.parameters() if p.requires_grad])
    gb = torch.autograd.grad(sum(v.sum() for v in b.values()), [p for p in ref.parameters() if p.requires_grad])
    for u, v in zip(ga, gb):
        torch.testing.assert_close(u, v, rtol=1e-4, atol=1e-4 * float(v.abs().max()))
