#!/usr/bin/env python
"""Model-level golden fixtures recorded FROM THE REFERENCE (build container only).

Runs the unmodified reference classes from /root/reference/src (DeformableDETR, DeformableDETRTracking,
SetCriterion, HungarianMatcher, ...) on CPU, with
  * `visdom` and the compiled extension stubbed (neither is importable here),
  * the extension call replaced by the reference's own pure-PyTorch ms_deform_attn_core_pytorch,
  * backbone weights never downloaded,
driven by the shared case runners in tests/model_fixtures.py, and stores the outputs as
tests/golden/model_<case>.npz.  The product is later driven by the very same runners.
"""
import os
import sys
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def import_reference():
    vis = types.ModuleType("visdom")
    vis.Visdom = type("Visdom", (), {})
    sys.modules["visdom"] = vis
    sys.modules["MultiScaleDeformableAttention"] = types.ModuleType("MultiScaleDeformableAttention")
    sys.path.insert(0, os.path.join(REF, "src"))
    import trackformer.models.backbone as bb
    bb.is_main_process = lambda: False
    import trackformer.models.ops.modules.ms_deform_attn as mod
    from trackformer.models.ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch

    class _Shim:
        @staticmethod
        def apply(value, shapes, loc, attn, step):
            return ms_deform_attn_core_pytorch(value, shapes, loc, attn)
    mod.MSDeformAttnFunction = _Shim
    from trackformer.models import build_model
    from trackformer.util.misc import nested_dict_to_namespace
    return build_model, nested_dict_to_namespace


def main():
    import model_fixtures as mf
    build_model, to_ns = import_reference()

    def load(name):
        return yaml.safe_load(open(os.path.join(REF, "cfgs", name)))

    def build(tracking, multi_frame, **overrides):
        cfg = load("train.yaml")
        cfg.update(load("train_deformable.yaml"))
        if tracking:
            cfg.update(load("train_tracking.yaml"))
            cfg["dataset"] = "mot"
        if multi_frame:
            cfg.update(load("train_multi_frame.yaml"))
        cfg["device"] = "cpu"
        cfg.update(overrides)
        torch.manual_seed(0)
        model, criterion, _ = build_model(to_ns(cfg))
        return model, criterion

    cases = {
        "det_mini": lambda: mf.run_detection(build, [(160, 224)]),
        "det_c1_480x640": lambda: mf.run_detection(build, [(480, 640)]),
        "det_padded_batch": lambda: mf.run_detection(build, [(160, 224), (128, 192)]),
        "track_two_frames": lambda: mf.run_two_frame_tracking(build, (160, 224), 12),
        "track_multi_frame": lambda: mf.run_two_frame_tracking(build, (128, 160), 9, multi_frame=True),
        "train_step_det": lambda: mf.run_train_step(build, [(160, 224), (160, 224)], 6),
        "train_step_tracking": lambda: mf.run_train_step(build, [(128, 160)], 7, tracking=True),
        # BASELINE.json configs[1], [2], [4] at their full sizes (outputs only; a few hundred KB each)
        "det_c2_800x1333": lambda: mf.run_detection(build, [(800, 1333)]),
        "train_step_c2": lambda: mf.run_train_step(build, [(800, 1333)], 20, optimizer=True),
        "track_c3_800x1333": lambda: mf.run_two_frame_tracking(build, (800, 1333), 100),
        "track_c5_1080x1920": lambda: mf.run_two_frame_tracking(build, (1080, 1920), 300, multi_frame=True),
    }
    only = sys.argv[1:]
    for name, fn in cases.items():
        if only and name not in only:
            continue
        res = fn()
        path = os.path.join(HERE, f"model_{name}.npz")
        np.savez(path, **{k: np.asarray(v) for k, v in res.items()})
        print(f"{name:22s} -> {os.path.relpath(path)} ({os.path.getsize(path)/1024:.0f} KiB)", flush=True)

    if not only or "bookkeeping" in only:
        model, _ = build(True, False)
        res = mf.run_bookkeeping(model, 40)
        path = os.path.join(HERE, "model_bookkeeping.npz")
        np.savez(path, **res)
        print(f"bookkeeping            -> {os.path.relpath(path)} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
