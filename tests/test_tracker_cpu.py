"""Tracker bookkeeping parity on CPU: trackformer_b200.tracker.Tracker against fixtures recorded from the reference
Tracker (tests/golden/make_golden_tracker.py) on the scripted scene of tests/tracker_fixtures.py.

Ids, frames, obj_ind, ReID counts and list orders must be identical; scores and boxes are the detector's fp32 values
pushed through the same arithmetic, compared to 1e-6 relative."""
import os

import numpy as np
import pytest
import torch

import tracker_fixtures as tf
from trackformer_b200.deformable_detr import DeformablePostProcess
from trackformer_b200.tracker import Tracker, nms_keep, pairwise_iou

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def check_against_gold(out, case, rtol=1e-6):
    gold = np.load(os.path.join(GOLD, f"tracker_{case}.npz"))
    for key in ("num_reids", "track_num", "frame_index", "active_ids", "inactive_ids", "inactive_counts", "prev_log",
                "per_frame_ids"):
        np.testing.assert_array_equal(out[key], gold[key], err_msg=f"{case}: {key}")
    assert out["rows"].shape == gold["rows"].shape, case
    np.testing.assert_array_equal(out["rows"][:, :3], gold["rows"][:, :3], err_msg=f"{case}: (id, frame, obj_ind)")
    np.testing.assert_allclose(out["rows"][:, 3:], gold["rows"][:, 3:], rtol=rtol, atol=1e-4 * rtol / 1e-6,
                               err_msg=f"{case}: score / bbox")


@pytest.mark.parametrize("case", list(tf.CASES))
def test_tracker_matches_reference(case):
    out = tf.run_case(Tracker, DeformablePostProcess(), case)
    check_against_gold(out, case)


def test_fixtures_exercise_every_branch():
    """the scene must actually trigger ReID (both kinds), both NMS passes and the public-detection gate"""
    g = {c: np.load(os.path.join(GOLD, f"tracker_{c}.npz")) for c in tf.CASES}
    assert g["reid_embedding"]["num_reids"] > 0 and g["reid_greedy"]["num_reids"] > 0
    assert g["default"]["num_reids"] == 0 and len(g["default"]["inactive_ids"]) == 0
    assert g["public_center"]["track_num"] < g["default"]["track_num"]
    # prev_features handed to the detector: the frame before (prev_frame_dist 1) / two frames before (2)
    assert g["default"]["prev_log"].tolist()[:3] == [-1, 0, 1] and g["public_iou"]["prev_log"].tolist()[:4] == [-1, -1, 0, 1]
    logs = []
    scene = tf.Scene()
    det = tf.ScriptedDetector(scene)
    tr = Tracker(det, {"bbox": DeformablePostProcess()}, tf.tracker_cfg("reid_embedding"), False, logger=logs.append)
    for blob in tf.blobs(scene):
        tr.step(blob)
    text = "\n".join(logs)
    for needle in ("REMOVE TRACK IDS (detection_nms_thresh", "REID: track.id", "NEW INACTIVE TRACK IDS", "INIT TRACK IDS"):
        assert needle in text, needle


def test_nms_matches_torchvision():
    from torchvision.ops import nms
    rng = np.random.RandomState(0)
    for n in (1, 2, 17, 120):
        xy = rng.uniform(0, 100, (n, 2)).astype(np.float32)
        wh = rng.uniform(5, 60, (n, 2)).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = rng.uniform(0, 1, n).astype(np.float32)
        scores[rng.rand(n) < 0.3] = np.inf                  # established tracks (tracker.py:503)
        for thr in (0.3, 0.6, 0.9):
            ref = np.zeros(n, bool)
            ref[nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()] = True
            np.testing.assert_array_equal(nms_keep(boxes, scores, thr), ref)
    assert nms_keep(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.5).shape == (0,)


def test_pairwise_iou_matches_torchvision():
    from torchvision.ops import box_iou
    rng = np.random.RandomState(1)
    a = np.sort(rng.uniform(0, 50, (7, 2, 2)).astype(np.float32), axis=1).reshape(7, 4)
    b = np.sort(rng.uniform(0, 50, (5, 2, 2)).astype(np.float32), axis=1).reshape(5, 4)
    np.testing.assert_array_equal(pairwise_iou(a, b), box_iou(torch.from_numpy(a), torch.from_numpy(b)).numpy())


def test_reset_and_track_views():
    scene = tf.Scene()
    det = tf.ScriptedDetector(scene)
    tr = Tracker(det, {"bbox": DeformablePostProcess()}, tf.tracker_cfg("reid_embedding"), False)
    frames = list(tf.blobs(scene, 6))
    for blob in frames:
        tr.step(blob)
    assert tr.frame_index == 6 and tr.tracks and tr.get_results() is tr.results
    t = tr.tracks[0]
    assert t.pos.shape == (4,) and t.hs_embed[-1].shape == (scene.c,) and t.has_positive_area()
    n_results = len(tr.results)
    tr.reset(hard=False)                                     # keeps ids / results, drops the live tracks
    assert tr.tracks == [] and tr.inactive_tracks == [] and tr.frame_index == 6 and len(tr.results) == n_results
    tr.reset()
    assert tr.frame_index == 0 and tr.results == {} and tr.track_num == 0
    with pytest.raises(ValueError):
        Tracker(det, {"bbox": DeformablePostProcess()}, tf.tracker_cfg("default"), True)


def test_many_tracks_grow_the_bank():
    """more simultaneous targets than the initial capacity (64 slots)"""
    class Many(torch.nn.Module):
        num_queries, overflow_boxes = 150, True

        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        def forward(self, img, targets=None, prev_features=None):
            k = 0 if targets is None else len(targets[0]["track_query_boxes"])
            g = torch.Generator().manual_seed(3)
            centres = torch.rand(150, 2, generator=g) * 0.8 + 0.1
            boxes = torch.cat([centres, torch.full((150, 2), 0.01)], 1)
            logits = torch.full((k + 150, 4), -5.0)
            logits[:, 0] = 2.0
            embeds = torch.arange(k + 150, dtype=torch.float32)[:, None].expand(-1, 8).contiguous()
            if k:
                boxes = torch.cat([targets[0]["track_query_boxes"], boxes], 0)
                logits[k:, 0] = -2.0
                embeds[:k] = targets[0]["track_query_hs_embeds"]
            return {"pred_logits": logits[None], "pred_boxes": boxes[None], "hs_embed": embeds[None]}, None, None, None, None

    tr = Tracker(Many(), {"bbox": DeformablePostProcess()}, tf.tracker_cfg("default"), False)
    blob = {"img": torch.zeros(1, 3, 8, 8), "orig_size": torch.tensor([[1000, 1000]]), "dets": torch.zeros(1, 0, 4)}
    for _ in range(3):
        tr.step(blob)
    assert len(tr.tracks) == 150 and tr.track_num == 150
    assert sorted(t.id for t in tr.tracks) == list(range(150))
    # embeddings followed their tracks through the growth of the bank
    for t in tr.tracks:
        assert float(t.hs_embed[-1][0]) == float(t.obj_ind.item())


@pytest.mark.parametrize("name,multi_frame", [("model_sequence", False), ("model_sequence_multi_frame", True)])
def test_tracker_over_the_real_detector_matches_reference(name, multi_frame, monkeypatch):
    """Reference Tracker + reference tracking model vs ours + ours on a 5-frame random 'video' (canonical weights, CPU,
    the CUDA op replaced by the oracle's torch restatement).  Thresholds sit in gaps of the score distribution; the
    fixture records the smallest |score - threshold| of the run, which must dwarf the model-level fp32 noise."""
    from oracle.torch_ref import msda_core_torch
    import trackformer_b200.msda_module as mm
    from test_model_parity_cpu import build

    class _OracleFn:
        @staticmethod
        def apply(value, shapes, loc, attn, step):
            return msda_core_torch(value, shapes, loc, attn)
    monkeypatch.setattr(mm, "MSDeformAttnFunction", _OracleFn)
    gold = np.load(os.path.join(GOLD, f"tracker_{name}.npz"))
    assert gold["margin"] > 1e-5
    cfg = {k: float(v) for k, v in zip(gold["cfg_keys"], gold["cfg_values"])}
    cfg["prev_frame_dist"] = int(cfg["prev_frame_dist"])
    out = tf.run_model_sequence(build, Tracker, DeformablePostProcess(), cfg, multi_frame=multi_frame)
    for key in ("num_reids", "track_num", "frame_index", "active_ids", "inactive_ids", "inactive_counts"):
        np.testing.assert_array_equal(out[key], gold[key], err_msg=key)
    assert out["rows"].shape == gold["rows"].shape
    np.testing.assert_array_equal(out["rows"][:, :3], gold["rows"][:, :3])
    np.testing.assert_allclose(out["rows"][:, 3:], gold["rows"][:, 3:], rtol=1e-3, atol=1e-3)


class _Silent(torch.nn.Module):
    """detector that never scores above any threshold"""
    num_queries, overflow_boxes, hidden_dim = 10, False, 8

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.calls = []

    def forward(self, img, targets=None, prev_features=None):
        self.calls.append(None if targets is None else len(targets[0]["track_query_boxes"]))
        q = self.num_queries + (self.calls[-1] or 0)
        out = {"pred_logits": torch.full((1, q, 3), -6.0), "pred_boxes": torch.full((1, q, 4), 0.5),
               "hs_embed": torch.zeros(1, q, self.hidden_dim)}
        return out, None, None, None, None


def test_no_detections_means_no_tracks_and_no_track_queries():
    det = _Silent()
    tr = Tracker(det, {"bbox": DeformablePostProcess()}, tf.tracker_cfg("reid_embedding"), False)
    for _ in range(3):
        tr.step({"img": torch.zeros(1, 3, 8, 8), "orig_size": torch.tensor([[100, 200]])})      # no 'dets' key either
    assert tr.results == {} and tr.tracks == [] and tr.inactive_tracks == [] and tr.track_num == 0
    assert det.calls == [None, None, None] and tr.frame_index == 3


def test_boxes_are_clipped_to_the_image_unless_overflow_is_allowed():
    class Edge(_Silent):
        def forward(self, img, targets=None, prev_features=None):
            out, *rest = super().forward(img, targets, prev_features)
            k = out["pred_logits"].shape[1] - self.num_queries
            out["pred_logits"][0, k, 0] = 3.0                       # one confident person ...
            out["pred_boxes"][0, k] = torch.tensor([0.98, 0.5, 0.2, 0.4])      # ... sticking out on the right
            return (out, *rest)
    for overflow, x1 in ((False, 200.0), (True, 216.0)):
        det = Edge()
        det.overflow_boxes = overflow
        tr = Tracker(det, {"bbox": DeformablePostProcess()}, tf.tracker_cfg("default"), False)
        tr.step({"img": torch.zeros(1, 3, 8, 8), "orig_size": torch.tensor([[100, 200]]), "dets": torch.zeros(1, 0, 4)})
        (res,) = tr.results[0].values()
        np.testing.assert_allclose(res["bbox"], [176.0, 30.0, x1, 70.0], rtol=1e-6)
        assert res["obj_ind"] == 0 and res["score"].shape == () and res["score"].dtype == np.float32


def test_generic_postprocessor_route_equals_packed_route():
    """a reference-shaped post-processor (no `packed`) gives the tracker the same rows"""
    class Plain:
        def __init__(self):
            self.inner = DeformablePostProcess()

        def __call__(self, outputs, sizes):
            return self.inner(outputs, sizes)
    a = tf.run_case(Tracker, DeformablePostProcess(), "reid_greedy")
    b = tf.run_case(Tracker, Plain(), "reid_greedy")
    for key in a:
        np.testing.assert_array_equal(a[key], b[key], err_msg=key)
