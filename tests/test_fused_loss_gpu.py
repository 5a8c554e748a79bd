"""Fused SetCriterion / matching-cost kernels (csrc/set_loss.cu) against the PyTorch op chains they replace
(the mirror of models/detr.py:213-328 / matcher.py:60-75 in trackformer_b200/criterion.py, matcher.py) and the fused
box refinement (csrc/box_refine.cu) against util.inverse_sigmoid."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _problem(dev, k=6, bs=2, nq=300, nc=91, sizes=(20, 7), seed=0):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(k, bs, nq, nc, generator=g) * 2 - 2).to(dev).requires_grad_(True)
    boxes = torch.cat([torch.rand(k, bs, nq, 2, generator=g) * 0.6 + 0.2, torch.rand(k, bs, nq, 2, generator=g) * 0.3 + 0.02],
                      -1).to(dev).requires_grad_(True)
    targets = []
    for n in sizes:
        b = torch.cat([torch.rand(n, 2, generator=g) * 0.6 + 0.2, torch.rand(n, 2, generator=g) * 0.25 + 0.05], 1)
        targets.append({"boxes": b.to(dev), "labels": torch.randint(0, nc, (n,), generator=g).to(dev)})
    return logits, boxes, targets


def _criterion(dev):
    from trackformer_b200.model_factory import build_model, default_args
    _, criterion, _ = build_model(default_args(device=str(dev), enc_layers=1, dec_layers=1, num_queries=20))
    return criterion.to(dev).train()


def test_match_cost_kernel_equals_op_chain(cuda_device):
    from trackformer_b200 import ext
    dev = cuda_device
    crit = _criterion(dev)
    m = crit.matcher
    logits, boxes, targets = _problem(dev)
    tgt_ids = torch.cat([t["labels"] for t in targets])
    tgt_boxes = torch.cat([t["boxes"] for t in targets])
    with torch.no_grad():
        ref = m._cost(logits.flatten(0, 2), boxes.flatten(0, 2), tgt_ids, tgt_boxes)
        got = ext.load().match_cost(logits.detach(), boxes.detach(), tgt_ids, tgt_boxes, m.cost_class, m.cost_bbox,
                                    m.cost_giou, m.focal_alpha, m.focal_gamma)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("sizes", [(20, 7), (1, 1), (33,)])
def test_fused_set_loss_equals_op_chain(cuda_device, sizes, monkeypatch):
    import trackformer_b200.criterion as C
    import trackformer_b200.matcher as M
    dev = cuda_device
    crit = _criterion(dev)
    assert crit.device_matcher
    res = {}
    for fused in (False, True):
        monkeypatch.setattr(C, "_FUSED_LOSS", fused)
        monkeypatch.setattr(M, "_FUSED_LOSS", fused)
        logits, boxes, targets = _problem(dev, bs=len(sizes), sizes=sizes, seed=3)
        num_boxes = torch.tensor(float(sum(sizes)) / 1.5, device=dev)
        losses = crit.forward_stacked(logits, boxes, targets, num_boxes)
        wd = crit.weight_dict
        total = sum(losses[k] * wd[k] for k in losses if k in wd)
        total.backward()
        res[fused] = ({k: v.detach().clone() for k, v in losses.items()}, logits.grad.clone(), boxes.grad.clone())
    ref, got = res[False], res[True]
    assert set(ref[0]) == set(got[0])
    for k in ref[0]:
        torch.testing.assert_close(got[0][k], ref[0][k], rtol=2e-4, atol=1e-5, msg=lambda m, k=k: f"{k}: {m}")
    torch.testing.assert_close(got[1], ref[1], rtol=2e-4, atol=1e-7)
    torch.testing.assert_close(got[2], ref[2], rtol=2e-4, atol=1e-6)


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_refine_boxes_equals_op_chain(cuda_device, ref_dim):
    from trackformer_b200.util import inverse_sigmoid, refine_boxes
    dev = cuda_device
    g = torch.Generator().manual_seed(ref_dim)
    delta = torch.randn(2, 300, 4, generator=g).to(dev).requires_grad_(True)
    ref = torch.rand(2, 300, ref_dim, generator=g)
    ref[0, :5] = torch.tensor([0.0, 1.0, 1e-6, 1 - 1e-6][:ref_dim])          # the clamp gates
    ref = ref.to(dev).requires_grad_(True)
    out = refine_boxes(delta, ref)
    gout = torch.randn_like(out)
    out.backward(gout)
    d2 = delta.detach().clone().requires_grad_(True)
    r2 = ref.detach().clone().requires_grad_(True)
    inv = inverse_sigmoid(r2)
    exp = (d2 + inv).sigmoid() if ref_dim == 4 else torch.cat([d2[..., :2] + inv, d2[..., 2:]], -1).sigmoid()
    exp.backward(gout)
    torch.testing.assert_close(out, exp, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(delta.grad, d2.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(ref.grad, r2.grad, rtol=1e-4, atol=1e-4)
