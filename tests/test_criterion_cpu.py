"""The stacked (all-layers-at-once) criterion must reproduce the per-layer loop of the reference
(models/detr.py:382-443) key by key."""
import torch

from trackformer_b200.model_factory import build_model, default_args


def test_forward_stacked_equals_layer_loop():
    torch.manual_seed(0)
    _, criterion, _ = build_model(default_args(device="cpu", enc_layers=1, dec_layers=6, num_queries=40))
    g = torch.Generator().manual_seed(5)
    k, bs, nq, c = 6, 3, 40, 91
    logits = torch.randn(k, bs, nq, c, generator=g, requires_grad=True)
    boxes_raw = torch.randn(k, bs, nq, 4, generator=g, requires_grad=True)
    boxes = boxes_raw.sigmoid() * 0.5 + 0.2
    targets = []
    for b in range(bs):
        n = 4 + 3 * b
        targets.append({"labels": torch.randint(0, c, (n,), generator=g),
                        "boxes": torch.cat([torch.rand(n, 2, generator=g) * 0.5 + 0.25,
                                            torch.rand(n, 2, generator=g) * 0.3 + 0.05], 1)})
    out = {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
           "aux_outputs": [{"pred_logits": logits[i], "pred_boxes": boxes[i]} for i in range(k - 1)]}
    ref = criterion(out, targets)
    got = criterion.forward_stacked(logits, boxes, targets)
    assert set(ref) == set(got)
    for key in ref:
        torch.testing.assert_close(got[key], ref[key], rtol=1e-5, atol=1e-6, msg=key)
    wd = criterion.weight_dict
    l_ref = sum(ref[key] * wd[key] for key in ref if key in wd)
    l_got = sum(got[key] * wd[key] for key in got if key in wd)
    g_ref = torch.autograd.grad(l_ref, (logits, boxes_raw), retain_graph=True)
    g_got = torch.autograd.grad(l_got, (logits, boxes_raw))
    for a, b in zip(g_got, g_ref):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-7)
