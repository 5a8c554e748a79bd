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
    for case in tf.CASES:
        out = tf.run_case(Tracker, DeformablePostProcess(), case)
        path = os.path.join(HERE, f"tracker_{case}.npz")
        np.savez_compressed(path, **out)
        ids = sorted(set(out["rows"][:, 0].astype(int)))
        print(f"{case}: {len(out['rows'])} result rows, {len(ids)} ids, reids={int(out['num_reids'])}, "
              f"track_num={int(out['track_num'])}, final active={out['active_ids'].tolist()} "
              f"inactive={out['inactive_ids'].tolist()}")


if __name__ == "__main__":
    main()
