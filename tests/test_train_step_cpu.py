"""Host-side logic of the flat parameter layout (TrainStep(flat_adamw=...)) on CPU: parameters move into one buffer
without changing the model; the update itself is CUDA-only and must refuse CPU tensors loudly."""
import pytest
import torch

import model_fixtures as mf
from test_model_parity_cpu import build, oracle_op  # noqa: F401  (fixture)


def test_flat_layout_keeps_the_model_and_partitions_groups(oracle_op):
    from trackformer_b200.flat_adamw import reference_param_groups
    from trackformer_b200.train_step import TrainStep
    model, criterion = build(False, False, enc_layers=1, dec_layers=1, num_queries=20, dropout=0.0)
    mf.canonical_weights_(model, 0)
    model.eval()
    x = mf.make_images(3, [(96, 128)])[0][None]
    with torch.no_grad():
        before = model(x)[0]["pred_boxes"].clone()
    groups = reference_param_groups(model)
    n_trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert sum(p.numel() for g in groups for p in g["params"]) == n_trainable
    step = TrainStep(model, criterion, None, use_graphs=False, flat_adamw={"groups": groups})
    ranges = step.flat_optimizer.ranges
    # (TrainStep moves the backbone group to the end of the buffers: ranges follow the optimizer's own group order)
    placed = step.flat_optimizer.param_groups
    assert sorted(id(p) for g in placed for p in g["params"]) == sorted(id(p) for g in groups for p in g["params"])
    assert all(e - b >= sum(p.numel() for p in g["params"]) for (b, e), g in zip(ranges, placed))
    names = {id(p): n for n, p in model.named_parameters()}
    assert all(names[id(p)].startswith("backbone.") for p in placed[-1]["params"])
    assert all(b % 64 == 0 for b, _ in ranges) and all((p.data_ptr() - step.flat_param.data_ptr()) % 256 == 0 and (p.grad.data_ptr() - step.flat_grad.data_ptr()) % 256 == 0 for p in step.params)
    assert all(b % 4 == 0 for b, _ in ranges) and all(ranges[i][1] <= ranges[i + 1][0] for i in range(len(ranges) - 1))
    lo, hi = step.flat_param.data_ptr(), step.flat_param.data_ptr() + 4 * step.flat_param.numel()
    for p in step.params:
        assert lo <= p.data_ptr() < hi and lo <= p.grad.data_ptr() - step.flat_grad.data_ptr() + lo < hi
        assert p.grad.stride() == p.stride()
    with torch.no_grad():
        after = model(x)[0]["pred_boxes"]
    assert torch.equal(before, after)
    # writing through the flat buffer is writing the parameters
    step.flat_param.zero_()
    assert all(float(p.detach().abs().sum()) == 0.0 for p in step.params)
    with pytest.raises(RuntimeError):
        step.flat_optimizer.step()                       # CUDA-only update: no CPU fallback
    with pytest.raises(AssertionError):
        TrainStep(model, criterion, None, use_graphs=False, flat_adamw={"groups": groups[:2]})


def test_detector_core_returns_the_stacked_heads(oracle_op):
    """_DetectorCore hands out the per-layer heads the model stacked itself; values and gradients equal re-stacking the
    slices of the output dictionary"""
    from trackformer_b200.train_step import _DetectorCore
    model, _ = build(False, False, enc_layers=1, dec_layers=3, num_queries=20, dropout=0.0)
    mf.canonical_weights_(model, 0)
    model.train()
    core = _DetectorCore(model)
    x = mf.make_images(3, [(96, 128)])[0][None]
    logits, boxes = core(x)
    assert logits.shape[0] == boxes.shape[0] == 3 and model.stacked_heads[0] is logits
    import copy
    clone = copy.deepcopy(model)                         # still deep-copyable after a training forward
    assert clone.stacked_heads is None and len(clone.state_dict()) == len(model.state_dict())
    out = model(x, None, None)[0]
    re_logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
    re_boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
    torch.testing.assert_close(logits, re_logits, rtol=0, atol=0)
    torch.testing.assert_close(boxes, re_boxes, rtol=0, atol=0)
    params = [p for p in model.parameters() if p.requires_grad]
    w = torch.randn(logits.shape, generator=torch.Generator().manual_seed(0))
    ga = torch.autograd.grad((logits * w).sum() + boxes.sum(), params, allow_unused=True)
    gb = torch.autograd.grad((re_logits * w).sum() + re_boxes.sum(), params, allow_unused=True)
    for a, b in zip(ga, gb):
        assert (a is None) == (b is None)
        if a is not None:
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6 * max(1.0, float(b.abs().max())))


def test_gradient_handover_equals_accumulation(oracle_op):
    """_backward_into_flat: gathering the handed-over gradients gives the flat buffer the accumulating path gives,
    including a zeroed slice for a parameter that received no gradient"""
    import copy
    from trackformer_b200.train_step import TrainStep
    model, criterion = build(False, False, enc_layers=1, dec_layers=2, num_queries=20, dropout=0.0)
    mf.canonical_weights_(model, 0)
    model.train()
    model_b = copy.deepcopy(model)
    x = mf.make_images(3, [(96, 128)])[0][None]
    targets = mf.make_targets(4, 1, 3)
    gather = TrainStep(model, criterion, None, use_graphs=False)
    accumulate = TrainStep(model_b, criterion, None, use_graphs=False)
    gather.gather_grads, accumulate.gather_grads = True, False
    gather.flat_grad.fill_(7.0)                          # stale content must not survive either path
    accumulate.flat_grad.fill_(7.0)
    la, lb = gather(x, targets), accumulate(x, targets)
    torch.testing.assert_close(la, lb, rtol=0, atol=0)
    pads = torch.ones_like(gather.flat_grad, dtype=torch.bool)
    for p in gather.params:
        o = (p.grad.data_ptr() - gather.flat_grad.data_ptr()) // 4
        pads[o:o + p.numel()] = False
        assert p.grad.data_ptr() >= gather.flat_grad.data_ptr()          # the flat views are back in place
    torch.testing.assert_close(gather.flat_grad[~pads], accumulate.flat_grad[~pads], rtol=0, atol=0)
    assert float(gather.flat_grad[~pads].abs().sum()) > 0


def test_seed_pool_hands_out_distinct_views_and_falls_back():
    """seeds.py: one draw per step, consecutive one-element views; outside a step (or past the pool) a per-call draw."""
    from trackformer_b200 import seeds
    seeds.end_step()
    a = seeds.next_seed("cpu")
    assert a.shape == (1,) and a.dtype == torch.int64
    seeds.begin_step("cpu")
    views = [seeds.next_seed("cpu") for _ in range(seeds._POOL_SIZE + 3)]
    pool = seeds._pool
    assert all(v.shape == (1,) for v in views)
    assert views[0].data_ptr() == pool.data_ptr() and views[1].data_ptr() == pool.data_ptr() + 8
    assert views[-1].data_ptr() not in range(pool.data_ptr(), pool.data_ptr() + 8 * seeds._POOL_SIZE)   # fell back
    old = views[0].clone()
    seeds.begin_step("cpu")                                   # a new pool is a new tensor: earlier views stay valid
    assert torch.equal(views[0], old)
    seeds.end_step()


def test_prefetch_hands_frames_to_the_next_step(oracle_op):
    """host logic of TrainStep.prefetch on CPU: the staged frames are consumed exactly once"""
    import copy
    from trackformer_b200.train_step import TrainStep
    model, criterion = build(False, False, enc_layers=1, dec_layers=2, num_queries=20, dropout=0.0)
    mf.canonical_weights_(model, 0)
    model.train()
    model_b = copy.deepcopy(model)
    x = mf.make_images(3, [(96, 128)])[0][None]
    targets = mf.make_targets(4, 1, 3)
    a = TrainStep(model, criterion, None, use_graphs=False)
    b = TrainStep(model_b, criterion, None, use_graphs=False)
    with pytest.raises(ValueError, match="prefetch"):
        a(None, targets)
    a.prefetch(x)
    la, lb = a(None, targets), b(x, targets)
    torch.testing.assert_close(la, lb, rtol=0, atol=0)
    torch.testing.assert_close(a.flat_grad, b.flat_grad, rtol=0, atol=0)
    with pytest.raises(ValueError, match="prefetch"):
        a(None, targets)


def test_weighted_loss_total_from_stacked_vectors_equals_the_entry_sum(oracle_op):
    """TrainStep._loss: (stack(ce, l1, giou) * weights).sum() == sum(loss_dict[k] * weight_dict[k]) (engine.py:139-140),
    values and gradients"""
    from trackformer_b200.criterion import LossDict
    from trackformer_b200.train_step import TrainStep
    model, criterion = build(False, False, enc_layers=1, dec_layers=3, num_queries=20, dropout=0.0)
    mf.canonical_weights_(model, 0)
    model.train()
    step = TrainStep(model, criterion, None, use_graphs=False)
    x = mf.make_images(3, [(96, 128)])[0][None]
    targets = mf.make_targets(4, 1, 3)
    wd = criterion.weight_dict
    assert sorted(k for k in wd if k.endswith("_1")) == ["loss_bbox_1", "loss_ce_1", "loss_giou_1"]
    grads = []
    for naive in (False, True):
        logits, boxes = step.core(x)
        logits, boxes = logits.detach().requires_grad_(True), boxes.detach().requires_grad_(True)
        if naive:
            d = criterion.forward_stacked(logits, boxes, targets)
            assert isinstance(d, LossDict) and d.stacked is not None
            total = sum(d[k] * wd[k] for k in d if k in wd)
        else:
            total = step._loss(logits, boxes, targets)
        grads.append((total.detach(), *torch.autograd.grad(total, (logits, boxes))))
    for a, b in zip(*grads):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)
    # layers without an entry in weight_dict weigh nothing
    wd_backup = dict(wd)
    try:
        for k in list(wd):
            if k.endswith("_0"):
                del wd[k]
        step._loss_w_key = None
        logits, boxes = step.core(x)
        d = criterion.forward_stacked(logits, boxes, targets)
        torch.testing.assert_close(step._loss(logits, boxes, targets), sum(d[k] * wd[k] for k in d if k in wd),
                                   rtol=1e-5, atol=1e-7)
    finally:
        wd.clear()
        wd.update(wd_backup)
