"""Device-side Hungarian matching (csrc/lsa.cu) against scipy.optimize.linear_sum_assignment -- the solver the
reference calls on the host (matcher.py:127) -- and the stacked criterion with device matching against the host path."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,B,Q,sizes", [(6, 1, 300, [20]), (6, 2, 300, [20, 7]), (1, 3, 50, [50, 1, 13]),
                                         (2, 2, 900, [128, 200]), (3, 4, 17, [3, 0, 17, 5])])
def test_lsa_matches_scipy(cuda_device, K, B, Q, sizes):
    from trackformer_b200 import ext
    m = ext.load()
    g = torch.Generator().manual_seed(K * 1000 + Q)
    T = sum(sizes)
    cost = torch.randn(K, B, Q, T, generator=g) * 3
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
    src, tgt, status = m.lsa(cost.to(cuda_device), off.to(cuda_device), max(sizes))
    assert int(status) == 0
    src, tgt = src.cpu(), tgt.cpu()
    for k in range(K):
        for b in range(B):
            lo, hi = int(off[b]), int(off[b + 1])
            if hi == lo:
                continue
            rows, cols = linear_sum_assignment(cost[k, b, :, lo:hi].numpy())
            assert np.array_equal(src[k, lo:hi].numpy(), rows), (k, b)
            assert np.array_equal(tgt[k, lo:hi].numpy() - lo, cols), (k, b)


def test_lsa_on_real_matching_costs_and_criterion_equivalence(cuda_device, monkeypatch):
    """Costs as the matcher builds them (focal class cost + L1 + GIoU): device matching == scipy matching, and the
    stacked criterion gives the same losses / gradients with either."""
    from trackformer_b200.model_factory import build_model, default_args
    dev = cuda_device
    torch.manual_seed(0)
    _, criterion, _ = build_model(default_args(device=str(dev), enc_layers=1, dec_layers=6, num_queries=300))
    criterion = criterion.to(dev)
    g = torch.Generator().manual_seed(9)
    k, bs, nq, c = 6, 2, 300, 91
    logits = torch.randn(k, bs, nq, c, generator=g).to(dev).requires_grad_(True)
    raw = torch.randn(k, bs, nq, 4, generator=g).to(dev).requires_grad_(True)
    boxes = raw.sigmoid() * 0.5 + 0.2
    targets = []
    for b in range(bs):
        n = 20 - 9 * b
        targets.append({"labels": torch.randint(0, c, (n,), generator=g).to(dev),
                        "boxes": torch.cat([torch.rand(n, 2, generator=g) * 0.5 + 0.25,
                                            torch.rand(n, 2, generator=g) * 0.3 + 0.05], 1).to(dev)})
    assert criterion.device_matcher
    got = criterion.forward_stacked(logits, boxes, targets)
    criterion.device_matcher = False
    ref = criterion.forward_stacked(logits, boxes, targets)
    criterion.device_matcher = True
    assert set(got) == set(ref)
    for key in ref:
        torch.testing.assert_close(got[key], ref[key], rtol=1e-6, atol=1e-6, msg=key)
    wd = criterion.weight_dict
    lg = sum(got[x] * wd[x] for x in got if x in wd)
    lr = sum(ref[x] * wd[x] for x in ref if x in wd)
    for a, b_ in zip(torch.autograd.grad(lg, (logits, raw), retain_graph=True), torch.autograd.grad(lr, (logits, raw))):
        torch.testing.assert_close(a, b_, rtol=1e-5, atol=1e-7)
    # and the raw index tensors equal scipy's
    src, tgt, status = criterion.matcher.match_layers_device(logits.detach(), boxes.detach(), targets)
    host = criterion.matcher.match_layers([{"pred_logits": logits[i].detach(), "pred_boxes": boxes[i].detach()}
                                           for i in range(k)], targets)
    off = [0, 20, 31]
    for i in range(k):
        for b in range(bs):
            assert torch.equal(src[i, off[b]:off[b + 1]].cpu(), host[i][b][0])
            assert torch.equal(tgt[i, off[b]:off[b + 1]].cpu() - off[b], host[i][b][1])
