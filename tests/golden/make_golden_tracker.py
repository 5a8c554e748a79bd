#!/usr/bin/env python
"""Tracker golden fixtures recorded FROM THE REFERENCE (build container only).

Runs the unmodified reference `Tracker` (src/trackformer/models/tracker.py) with the reference's own deformable
`PostProcess` on CPU, driven by the scripted detector and scene of tests/tracker_fixtures.py, once per tracker
configuration in tracker_fixtures.CASES, and stores everything observable (result rows, ids per frame, ReID count, ...)
as tests/golden/tracker_<case>.npz.  tests/test_tracker_cpu.py / test_tracker_gpu.py drive the product with the same
scene and compare.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
sys.path.insert(0, HERE)


def main():
    from make_golden_model import import_reference
    import_reference()
    from trackformer.models.tracker import Tracker
    from trackformer.models.deformable_detr import DeformablePostProcess
    import tracker_fixtures as tf
    if "--scripted-only" not in sys.argv:
        model_sequences(tf, Tracker, DeformablePostProcess)
    for case in tf.CASES:
        out = tf.run_case(Tracker, DeformablePostProcess(), case)
        path = os.path.join(HERE, f"tracker_{case}.npz")
        np.savez_compressed(path, **out)
        ids = sorted(set(out["rows"][:, 0].astype(int)))
        print(f"{case}: {len(out['rows'])} result rows, {len(ids)} ids, reids={int(out['num_reids'])}, "
              f"track_num={int(out['track_num'])}, final active={out['active_ids'].tolist()} "
              f"inactive={out['inactive_ids'].tolist()}")


def model_sequences(tf, Tracker, PostProcess):
    """The reference Tracker over the reference tracking model (CPU, canonical weights) on a short random 'video'.
    Random weights give scores in a narrow band, so the thresholds are placed inside the widest gaps of the first
    frame's score distribution and the smallest |score - threshold| over the whole run is recorded as `margin`."""
    import torch
    import model_fixtures as mf
    from make_golden_model import import_reference
    import yaml
    build_model, to_ns = import_reference()
    REF = "/root/reference"

    def build(tracking, multi_frame, **overrides):
        cfg = yaml.safe_load(open(os.path.join(REF, "cfgs", "train.yaml")))
        for name in ("train_deformable.yaml", "train_tracking.yaml") + (("train_multi_frame.yaml",) if multi_frame else ()):
            cfg.update(yaml.safe_load(open(os.path.join(REF, "cfgs", name))))
        cfg["dataset"] = "mot"
        cfg["device"] = "cpu"
        cfg.update(overrides)
        torch.manual_seed(0)
        model, criterion, _ = build_model(to_ns(cfg))
        return model, criterion

    for name, multi_frame in (("model_sequence", False), ("model_sequence_multi_frame", True)):
        # probe run with thresholds nobody passes, to see the first frame's scores
        probe = tf.run_model_sequence(build, Tracker, PostProcess(), dict(detection_obj_score_thresh=2.0), n_frames=1,
                                      multi_frame=multi_frame, log_scores=True)
        s0 = np.sort(probe["_scores"][0][probe["_labels"][0] == 0])[::-1]
        gaps = s0[:-1] - s0[1:]
        lo, hi = 4, min(14, len(gaps))
        k = lo + int(np.argmax(gaps[lo:hi]))
        thr = float((s0[k] + s0[k + 1]) / 2)
        best = None
        for f_track, f_reid in ((0.97, 0.99), (0.96, 0.985), (0.975, 1.0), (0.95, 0.98), (0.98, 0.995), (0.965, 1.01)):
            cfg = dict(detection_obj_score_thresh=thr, track_obj_score_thresh=thr * f_track,
                       reid_score_thresh=thr * f_reid, inactive_patience=3, reid_sim_threshold=2.0,
                       detection_nms_thresh=0.7, track_nms_thresh=0.7, prev_frame_dist=1)
            out = tf.run_model_sequence(build, Tracker, PostProcess(), cfg, multi_frame=multi_frame, log_scores=True)
            margin = min(float(np.abs(np.concatenate(out["_scores"]) - t).min())
                         for t in (cfg["detection_obj_score_thresh"], cfg["track_obj_score_thresh"], cfg["reid_score_thresh"]))
            if best is None or margin > best[0]:
                best = (margin, cfg, out)
        margin, cfg, out = best
        out = {k: v for k, v in out.items() if not k.startswith("_")}
        out["cfg_keys"] = np.asarray(sorted(cfg))
        out["cfg_values"] = np.asarray([cfg[k] for k in sorted(cfg)], np.float64)
        out["margin"] = np.float64(margin)
        np.savez_compressed(os.path.join(HERE, f"tracker_{name}.npz"), **out)
        print(f"{name}: thr={thr:.6f} (gap {gaps[k]:.2e} at rank {k}), margin={margin:.2e}, rows={len(out['rows'])}, "
              f"ids={sorted(set(out['rows'][:, 0].astype(int)))}, reids={int(out['num_reids'])}, "
              f"active={out['active_ids'].tolist()} inactive={out['inactive_ids'].tolist()}")


if __name__ == "__main__":
    main()
