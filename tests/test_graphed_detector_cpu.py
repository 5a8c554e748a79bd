"""Bucketed track-query padding of GraphedDetector, checked on CPU with graphs off: filler queries must not change the
real rows (beyond fp32 summation noise of the masked softmax), for single- and multi-frame models, and a tracker
driven through the padded detector must make the reference's decisions."""
import os

import numpy as np
import pytest
import torch

import model_fixtures as mf
import tracker_fixtures as tf
from trackformer_b200.deformable_detr import DeformablePostProcess
from trackformer_b200.graphed_detector import GraphedDetector
from trackformer_b200.tracker import Tracker
from test_model_parity_cpu import build, oracle_op  # noqa: F401  (fixture)

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("multi_frame", [False, True])
def test_filler_queries_do_not_change_real_rows(oracle_op, multi_frame):
    model, _ = build(True, multi_frame)
    mf.canonical_weights_(model, 0)
    model.tracking()
    padded = GraphedDetector(model, bucket=8, use_graphs=False)
    assert padded.num_queries == model.num_queries and padded.overflow_boxes == model.overflow_boxes
    f1, f2 = tf.model_frames((96, 128), 2)
    with torch.no_grad():
        o1, _, feat1, _, _ = model(f1[None], None, None)
        g1, _, gfeat1, _, _ = padded(f1[None], None, None)
        torch.testing.assert_close(g1["pred_logits"], o1["pred_logits"], rtol=0, atol=0)
        for k in (3, 8, 13):                     # 8 is a whole bucket: no filler, bit-identical
            tgt = [{"track_query_boxes": o1["pred_boxes"][0, :k], "track_query_hs_embeds": o1["hs_embed"][0, :k],
                    "image_id": torch.tensor([1])}]
            o2, _, _, mem2, hs2 = model(f2[None], tgt, feat1)
            g2, _, _, gmem2, ghs2 = padded(f2[None], tgt, gfeat1)
            assert g2["pred_logits"].shape == (1, k + model.num_queries, o2["pred_logits"].shape[-1])
            tol = 0 if k % 8 == 0 else 2e-5
            for name in ("pred_logits", "pred_boxes", "hs_embed"):
                torch.testing.assert_close(g2[name], o2[name], rtol=tol, atol=tol)
            torch.testing.assert_close(ghs2, hs2, rtol=tol, atol=tol)
            torch.testing.assert_close(gmem2[0], mem2[0], rtol=0, atol=0)


def test_tracker_through_padded_detector_matches_reference(oracle_op):
    gold = np.load(os.path.join(GOLD, "tracker_model_sequence.npz"))
    cfg = {k: float(v) for k, v in zip(gold["cfg_keys"], gold["cfg_values"])}
    cfg["prev_frame_dist"] = int(cfg["prev_frame_dist"])

    # run_model_sequence builds + weights the model itself; wrap it when the tracker is constructed
    class PaddedTracker(Tracker):
        def __init__(self, model, post, cfg_, attn):
            super().__init__(GraphedDetector(model, bucket=16, use_graphs=False), post, cfg_, attn)
    out = tf.run_model_sequence(build, PaddedTracker, DeformablePostProcess(), cfg)
    for key in ("num_reids", "track_num", "active_ids", "inactive_ids"):
        np.testing.assert_array_equal(out[key], gold[key], err_msg=key)
    np.testing.assert_array_equal(out["rows"][:, :3], gold["rows"][:, :3])


def test_graph_mode_refuses_cpu_tensors():
    class Tiny(torch.nn.Module):
        num_queries, hidden_dim, overflow_boxes = 4, 8, True
    det = GraphedDetector(Tiny(), use_graphs=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        det(torch.zeros(1, 3, 8, 8))
    with pytest.raises(ValueError):
        GraphedDetector(Tiny(), bucket=0)
