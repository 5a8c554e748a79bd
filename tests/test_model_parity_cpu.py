"""CPU tests of the HOST-SIDE mirror (module / transformer / detector / tracking / criterion) against golden
outputs recorded from the reference classes (tests/golden/make_golden_model.py).

The product has no CPU path for the MSDeformAttn core, so these tests -- and only the tests -- substitute
the oracle's torch restatement for the CUDA function; everything else is product code.  The GPU twin of this
file (tests/test_model_parity_gpu.py) runs the same cases on the real kernels.
"""
import numpy as np
import pytest
import torch

import model_fixtures as mf
from conftest import load_golden


@pytest.fixture()
def oracle_op(monkeypatch):
    from oracle.torch_ref import msda_core_torch
    import trackformer_b200.msda_module as mm

    class _OracleFn:
        @staticmethod
        def apply(value, shapes, loc, attn, step):
            return msda_core_torch(value, shapes, loc, attn)
    monkeypatch.setattr(mm, "MSDeformAttnFunction", _OracleFn)


def build(tracking, multi_frame, **overrides):
    from trackformer_b200.model_factory import build_model, default_args
    torch.manual_seed(0)
    model, criterion, _ = build_model(default_args(tracking, multi_frame, device="cpu", **overrides))
    return model, criterion


def check(res, gold, keys, rtol=1e-3, atol=2e-5):
    for k in keys:
        np.testing.assert_allclose(np.asarray(res[k]), gold[k], rtol=rtol, atol=atol, err_msg=k)


def test_state_dict_contract():
    """597 keys / 40.85 M parameters with the reference's names; the
    name-substring optimiser groups of src/train.py:101-110 find their parameters."""
    model, _ = build(False, False)
    sd = model.state_dict()
    assert len(sd) == 597
    assert sum(p.numel() for p in model.parameters()) == 40_849_660
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 40_627_260
    names = [n for n, _ in model.named_parameters()]
    assert sum("sampling_offsets" in n for n in names) == 24 and sum("reference_points" in n for n in names) == 2
    for k in ("transformer.encoder.layers.0.self_attn.value_proj.weight", "transformer.level_embed",
              "transformer.decoder.layers.5.cross_attn.output_proj.bias", "input_proj.3.0.weight",
              "class_embed.5.weight", "bbox_embed.0.layers.2.bias", "query_embed.weight",
              "backbone.0.body.layer4.2.bn3.running_var", "transformer.decoder.layers.0.self_attn.in_proj_weight"):
        assert k in sd, k
    tm, _ = build(True, True)
    assert sum(p.numel() for p in tm.parameters()) == 44_233_874          # multi-frame model, appendix A.7


def test_detection_forward_matches_reference(oracle_op):
    gold = load_golden("det_mini", "model_")
    res = mf.run_detection(build, [(160, 224)])
    assert res["image_digest"] == str(gold["image_digest"])
    check(res, gold, ["pred_logits", "pred_boxes", "hs_last_mean", "aux4_boxes", "memory0_mean"])


def test_padded_batch_matches_reference(oracle_op):
    gold = load_golden("det_padded_batch", "model_")
    res = mf.run_detection(build, [(160, 224), (128, 192)])
    check(res, gold, ["pred_logits", "pred_boxes", "hs_last_mean", "aux4_boxes", "memory0_mean"])


def test_two_frame_tracking_matches_reference(oracle_op):
    gold = load_golden("track_two_frames", "model_")
    res = mf.run_two_frame_tracking(build, (160, 224), 12)
    assert res["f2_logits"].shape == (1, 312, 20)
    check(res, gold, ["f1_logits", "f1_boxes", "f2_logits", "f2_boxes", "f2_hs_mean"])


def test_multi_frame_tracking_matches_reference(oracle_op):
    gold = load_golden("track_multi_frame", "model_")
    res = mf.run_two_frame_tracking(build, (128, 160), 9, multi_frame=True)
    assert int(res["n_levels_memory"]) == int(gold["n_levels_memory"]) == 8
    check(res, gold, ["f1_logits", "f1_boxes", "f2_logits", "f2_boxes", "f2_hs_mean"])


def test_train_step_losses_and_grads_match_reference(oracle_op):
    gold = load_golden("train_step_det", "model_")
    res = mf.run_train_step(build, [(160, 224), (160, 224)], 6)
    keys = [k for k in gold if k.startswith("loss/")] + ["loss_total", "pred_logits", "pred_boxes"]
    check(res, gold, keys)
    gkeys = [k for k in gold if k.startswith("grad/")]
    assert len(gkeys) >= 10
    for k in gkeys:
        scale = float(np.abs(gold[k]).max())
        np.testing.assert_allclose(res[k], gold[k], rtol=2e-3, atol=2e-4 * max(scale, 1e-3), err_msg=k)
    np.testing.assert_allclose(res["grad_global_norm"], gold["grad_global_norm"], rtol=1e-3)


def test_tracking_train_step_bookkeeping_is_bit_exact(oracle_op):
    gold = load_golden("train_step_tracking", "model_")
    res = mf.run_train_step(build, [(128, 160)], 7, tracking=True)
    idx_keys = [k for k in gold if k.startswith("idx/")]
    assert idx_keys
    for k in idx_keys:                                   # integer bookkeeping: exact
        assert np.array_equal(res[k], gold[k]), k
    check(res, gold, [k for k in gold if k.startswith("loss/")] + ["loss_total"])


def test_add_track_queries_bit_exact_against_reference():
    gold = load_golden("bookkeeping", "model_")
    model, _ = build(True, False)
    res = mf.run_bookkeeping(model, 40)
    assert set(res) == set(gold.keys())
    for k in gold:
        if k.endswith(("match_ids", "track_mask", "fal_pos_mask")):
            assert np.array_equal(res[k], gold[k]), k      # indices and masks: bit-exact
        else:
            assert np.array_equal(res[k], gold[k]), k      # gathered rows of identical inputs: also exact


def test_msdeformattn_init_constants_match_reference_recipe():
    """ms_deform_attn.py:33-47: zero offset/attention weights, compass-grid offset bias x (point index + 1),
    zero attention/value/output biases."""
    from trackformer_b200.msda_module import MSDeformAttn
    m = MSDeformAttn(256, 4, 8, 4)
    grid = torch.tensor([-1, -1, -1, 0, -1, 1, 0, -1, 0, 1, 1, -1, 1, 0, 1, 1], dtype=torch.float32)
    grid = grid.view(8, 1, 1, 2).repeat(1, 4, 4, 1)
    for i in range(4):
        grid[:, :, i, :] *= i + 1
    assert torch.equal(m.sampling_offsets.bias.detach(), grid.view(-1))
    assert m.sampling_offsets.bias.requires_grad
    for t in (m.sampling_offsets.weight, m.attention_weights.weight, m.attention_weights.bias, m.value_proj.bias,
              m.output_proj.bias):
        assert not t.detach().any()
    assert m.value_proj.weight.detach().abs().max() > 0 and m.output_proj.weight.detach().abs().max() > 0


def test_frozen_batchnorm_affine_cache_tracks_its_buffers():
    """the cached (scale, shift) pair equals a fresh computation and is dropped when a buffer is written or copied"""
    import copy
    from trackformer_b200.backbone import FrozenBatchNorm2d
    bn = FrozenBatchNorm2d(6)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(6, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(6, generator=g))
        bn.running_mean.copy_(torch.randn(6, generator=g))
        bn.running_var.copy_(torch.rand(6, generator=g) + 0.1)
    x = torch.randn(2, 6, 5, 7, generator=g, requires_grad=True)

    def fresh(m, t):
        w, b, rm, rv = (v.view(1, -1, 1, 1) for v in (m.weight, m.bias, m.running_mean, m.running_var))
        scale = w * (rv + 1e-5).rsqrt()
        return t * scale + (b - rm * scale)
    y1 = bn(x)
    torch.testing.assert_close(y1, fresh(bn, x), rtol=1e-6, atol=1e-6)
    s1 = bn.affine()[0]
    assert bn.affine()[0] is s1                                   # reused
    y1.sum().backward()                                           # the cached pair works inside autograd
    assert x.grad is not None
    with torch.no_grad():
        bn.running_var.mul_(2.0)                                  # in-place write -> recomputed
    assert bn.affine()[0] is not s1
    torch.testing.assert_close(bn(x), fresh(bn, x), rtol=1e-6, atol=1e-6)
    clone = copy.deepcopy(bn)
    with torch.no_grad():
        clone.bias.add_(1.0)
    torch.testing.assert_close(clone(x), fresh(clone, x), rtol=1e-6, atol=1e-6)
    bn.load_state_dict({k: torch.ones(6) for k in ("weight", "bias", "running_mean", "running_var")})
    torch.testing.assert_close(bn(x), fresh(bn, x), rtol=1e-6, atol=1e-6)
    bn2 = FrozenBatchNorm2d(6)
    with torch.inference_mode():
        bn2(torch.zeros(1, 6, 2, 2))
    assert getattr(bn2, "_affine_cache", None) is None           # nothing cached under inference mode
    bn2(x).sum().backward()
    with torch.inference_mode():
        bn3 = FrozenBatchNorm2d(6)                               # buffers that ARE inference tensors
    assert bn3(torch.zeros(1, 6, 2, 2)).shape == (1, 6, 2, 2) and getattr(bn3, "_affine_cache", None) is None
